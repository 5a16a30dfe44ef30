#!/bin/bash
# run 16: one-launch GroupNorm, tile-width sweep
O=gpurun_out/run16; mkdir -p $O
echo "== pytest gn"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "groupnorm" 2>&1 | tail -4
echo "== gn fused (graph-timed)"
UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py gn 2>&1 | grep "^gn" | tee $O/gn_fused.log
echo "== gn three-kernel (graph-timed)"
UWU_GN_FUSED=0 UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py gn 2>&1 | grep "^gn" | tee $O/gn_three.log
echo "== lin_bn"
UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py lin_bn 2>&1 | grep "lin+\|addmm\|block_n" | tee $O/lin_bn.log
echo "== unet tests"
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== bench weak"
timeout 900 python bench.py --scaling weak --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo DONE
