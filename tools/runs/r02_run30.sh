#!/bin/bash
O=gpurun_out/run30; mkdir -p $O
echo "== pytest gn"; timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "groupnorm" 2>&1 | tail -3
echo "== gn (graph-timed)"; UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py gn 2>&1 | grep "^gn" | tee $O/gn_batched.log
echo "== bench weak"; timeout 300 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo DONE
