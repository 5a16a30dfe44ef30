#!/bin/bash
# run 35: final validation of the last commit
O=gpurun_out/run35; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (all)"; timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log; cp gpurun_out/sdxl_parity.json $O/ 2>/dev/null
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
echo "== bench (default)"; timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; cut -c1-200 $O/bench_final.json; tail -1 $O/bench_final.err
echo "== reference arm"; timeout 200 python bench.py --impl reference --steps 2 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-200 $O/bench_reference.json
echo "== breakdown"; timeout 300 python tools/step_breakdown.py > $O/breakdown.log 2>&1; head -14 $O/breakdown.log | tail -12
echo DONE
