#!/bin/bash
# run 25: same-box A/B of the persistent attention backward
O=gpurun_out/run25; mkdir -p $O
for v in 1 0 1 0; do
echo "== PERSIST=$v"
UWU_ATTN_BWD_PERSIST=$v timeout 300 python bench.py --scaling weak --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench.err; cut -c1-160 $O/bench_$v.json
done
UWU_ATTN_BWD_PERSIST=0 UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py attn 2>&1 | grep "attn_bwd"
UWU_ATTN_BWD_PERSIST=1 UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py attn 2>&1 | grep "attn_bwd"
echo DONE
