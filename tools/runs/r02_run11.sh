#!/bin/bash
for d in 0 4 8 16 32 64 24 96 120; do
echo "== dbg $d"
UWU_LF_DBG=$d UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py lokr_fused 2>&1 | grep "^lokr_fused M16384\|^lokr_fused M4096" 
done
echo DONE
