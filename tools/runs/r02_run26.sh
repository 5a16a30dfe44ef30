#!/bin/bash
# run 26: evidence refresh on the final build (pair kernel for the factored-route GEMMs, persistent attention backward)
O=gpurun_out/run26; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (all)"; timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log; cp gpurun_out/sdxl_parity.json $O/ 2>/dev/null
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench (default)"; timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; cut -c1-220 $O/bench_final.json; tail -1 $O/bench_final.err
echo "== bench weak"; timeout 400 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-220 $O/bench_weak.json
echo "== breakdown"; timeout 300 python tools/step_breakdown.py > $O/breakdown.log 2>&1; head -30 $O/breakdown.log | tail -28
echo "== perf12"; timeout 240 python tools/diag_gemm.py perf12 > $O/perf12.log 2>&1; tail -20 $O/perf12.log
echo "== kernels (graph-timed)"; UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py ln gn geglu attn attn_cross lokr_fused lokr_fact > $O/kernels_graph_timed.log 2>&1; grep -c "us" $O/kernels_graph_timed.log
echo "== ncu launch list (one eager step)"; UWU_PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --scaling weak --graph off --no-cpu-baseline > $O/ncu_list.log 2>&1; wc -l $O/launches.csv
echo DONE
