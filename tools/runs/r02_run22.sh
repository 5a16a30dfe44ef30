#!/bin/bash
# run 22: compact (32-column loop) GEMM epilogue
O=gpurun_out/run22; mkdir -p $O
echo "== pytest gemm/conv"; timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv or geglu" 2>&1 | tail -3
echo "== lin_bn"; UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py lin_bn 2>&1 | grep "block_n=0\|addmm" | tee $O/lin.log
echo "== perf12"; timeout 200 python tools/diag_gemm.py perf12 > $O/perf12.log 2>&1; tail -20 $O/perf12.log
echo "== unet tests"; timeout 240 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -2
echo "== bench weak"; timeout 300 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo DONE
