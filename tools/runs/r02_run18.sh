#!/bin/bash
# run 18: same-box A/B of the step: residual by TMA on/off, one-launch GroupNorm on/off
O=gpurun_out/run18; mkdir -p $O
for v in "1 1" "0 1" "1 0" "1 1"; do set -- $v
echo "== RES_TMA=$1 GN_FUSED=$2"
UWU_GEMM_RES_TMA=$1 UWU_GN_FUSED=$2 timeout 300 python bench.py --scaling weak --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$1_$2.json 2> $O/bench.err; cut -c1-160 $O/bench_$1_$2.json
done
echo "== breakdown"; timeout 300 python tools/step_breakdown.py > $O/breakdown.log 2>&1; head -64 $O/breakdown.log | tail -62
echo DONE
