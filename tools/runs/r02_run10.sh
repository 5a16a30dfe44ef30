#!/bin/bash
O=gpurun_out/run10; mkdir -p $O
for d in 0 1 2 3 7; do for st in 4 2; do
echo "== dbg $d stages $st"
UWU_LF_DBG=$d UWU_LF_STAGES=$st UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py lokr_fused 2>&1 | grep "^lokr_fused M" 
done; done
echo DONE
