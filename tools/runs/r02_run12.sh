#!/bin/bash
for d in 0 128 256 512 896 4; do
echo "== dbg $d"
UWU_LF_DBG=$d UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py lokr_fused 2>&1 | grep "^lokr_fused M16384\|^lokr_fused M4096" 
done
echo DONE
