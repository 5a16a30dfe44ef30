#!/bin/bash
O=gpurun_out/run37; mkdir -p $O
echo "== pytest gemm/lokr"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or lokr or conv" 2>&1 | tail -2
echo "== lokr_mirror (graph-timed)"; UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py lokr_mirror lokr_fact 2>&1 | grep "total\|gemm" | tee $O/lokr_mirror.log
echo "== unet"; timeout 240 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in 1 0 1; do echo "== bench weak MIRROR=$v"; UWU_LOKR_MIRROR=$v timeout 300 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak_$v.json 2> $O/bench_weak.err; cut -c1-170 $O/bench_weak_$v.json; done
echo DONE
