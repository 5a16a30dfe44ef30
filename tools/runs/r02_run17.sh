#!/bin/bash
# run 17: residual-by-TMA epilogue (every command under a short timeout: a hung kernel must not eat the budget)
O=gpurun_out/run17; mkdir -p $O
echo "== lin_bn res_tma on"
UWU_BENCH_GRAPH=1 timeout 90 python tools/bench_kernels.py lin_bn 2>&1 | grep "block_n=0\|addmm" | tee $O/lin_res_tma.log
echo "== pytest gemm/conv"
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv" 2>&1 | tail -4
echo "== unet tests"
timeout 240 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== bench weak"
timeout 300 python bench.py --scaling weak --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo DONE
