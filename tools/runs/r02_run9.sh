#!/bin/bash
# run 9: fused LoKr kernel v2 (4-stage ring, early producer)
O=gpurun_out/run9; mkdir -p $O
echo "== pytest lokr fused"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "lokr_fused" 2>&1 | tail -5
echo "== bench lokr_fused (graph-timed)"
UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py lokr_fused 2>&1 | tail -9 | tee $O/lokr_fused_graph.log
echo "== bench weak graph"
timeout 900 python bench.py --scaling weak --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-260 $O/bench_weak.json
echo DONE
