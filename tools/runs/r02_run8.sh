#!/bin/bash
# run 8: fused LoKr kernel — true device time (graph-timed), ncu stall picture, breakdown
O=gpurun_out/run8; mkdir -p $O
echo "== graph test"
timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -k graph 2>&1 | grep -E "^E|passed|failed" | cut -c1-300 | head
echo "== pytest lokr fused"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "lokr_fused" 2>&1 | tail -3
echo "== bench lokr_fused (graph-timed)"
UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py lokr_fused ln 2>&1 | tail -20 | tee $O/lokr_fused_graph.log
echo "== breakdown"; timeout 600 python tools/step_breakdown.py > $O/breakdown.log 2>&1; head -60 $O/breakdown.log
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lokr_fused --launch-skip 1 --launch-count 1 -o $O/lokr_fused -f python tools/profile_one.py lokr_fused > $O/ncu_lokr_fused.log 2>&1; tail -2 $O/ncu_lokr_fused.log
echo DONE
