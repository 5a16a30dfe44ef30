#!/bin/bash
# run 19: final evidence of round 2 (every command under its own timeout)
O=gpurun_out/run19; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt
echo "== pytest -m gpu (all)"; timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log; cp gpurun_out/sdxl_parity.json $O/ 2>/dev/null
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
echo "== bench (default: strong, graph on)"; timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; cut -c1-220 $O/bench_final.json; tail -2 $O/bench_final.err
echo "== bench weak"; timeout 400 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-220 $O/bench_weak.json
echo "== reference arm"; timeout 200 python bench.py --impl reference --steps 2 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
echo "== other configs"; for c in c2 c4 c1 latent; do timeout 300 python bench.py --config $c --steps 5 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err; cut -c1-200 $O/bench_$c.json; done
echo "== kernels (graph-timed)"; UWU_BENCH_GRAPH=1 timeout 300 python tools/bench_kernels.py ln gn geglu attn attn_cross lokr_fused noise > $O/kernels_graph_timed.log 2>&1; tail -45 $O/kernels_graph_timed.log
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 8400 --launch-count 3000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --scaling weak --graph off --no-cpu-baseline > $O/ncu_list.log 2>&1; tail -2 $O/ncu_list.log | cut -c1-200; wc -l $O/launches.csv
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 4 --launch-count 1 -o $O/gemm_lin -f python tools/profile_one.py gemm_lin > $O/ncu_gemm.log 2>&1; tail -1 $O/ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ln_bwd_stream|gn_fwd_fused|gn_bwd_fused|ln_fwd" --launch-count 8 -o $O/norms -f python tools/profile_one.py bw > $O/ncu_norms.log 2>&1; tail -1 $O/ncu_norms.log
echo DONE
