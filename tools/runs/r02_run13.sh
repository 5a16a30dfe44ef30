#!/bin/bash
O=gpurun_out/run13; mkdir -p $O
echo "== pytest lokr fused"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "lokr_fused" 2>&1 | tail -5
for st in 4 3 2; do
echo "== stages $st"
UWU_LF_STAGES=$st UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py lokr_fused 2>&1 | grep "^lokr_fused M\|^G route" | tee -a $O/lokr_fused_graph.log
done
echo "== bench weak graph"
timeout 900 python bench.py --scaling weak --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-260 $O/bench_weak.json
echo DONE
