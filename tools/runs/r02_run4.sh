#!/bin/bash
# round-2 GPU run 4: single-TMEM-read attention softmax, graph-capture fix, VAE / CLIP / sampling tests, profiles
cd "$(dirname "$0")/../.."
O=gpurun_out/run4; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_sdxl_parity_gpu.py > $O/pytest.log 2>&1; tail -12 $O/pytest.log
echo "== sdxl parity"; timeout 1200 python -m pytest tests/test_sdxl_parity_gpu.py -q -s > $O/sdxl_parity.log 2>&1; tail -8 $O/sdxl_parity.log; cp gpurun_out/sdxl_parity.json $O/ 2>/dev/null
echo "== attention kernels"; timeout 300 python tools/bench_kernels.py attn attn_cross > $O/attn.log 2>&1; tail -12 $O/attn.log
echo "== bench weak graph on"; timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline --graph on > $O/bench_weak_graph.json 2> $O/bench_weak_graph.err; cut -c1-200 $O/bench_weak_graph.json; tail -4 $O/bench_weak_graph.err
echo "== breakdown"; timeout 600 python tools/step_breakdown.py > $O/breakdown.log 2>&1; head -48 $O/breakdown.log
echo "== ncu attn"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel --launch-skip 1 --launch-count 1 -o $O/attn_fwd -f python tools/profile_one.py attn > $O/ncu_attn_fwd.log 2>&1; tail -2 $O/ncu_attn_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel --launch-skip 1 --launch-count 1 -o $O/attn_bwd -f python tools/profile_one.py attn > $O/ncu_attn_bwd.log 2>&1; tail -2 $O/ncu_attn_bwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 1 --launch-count 1 -o $O/gemm_res -f python tools/profile_one.py gemm_res > $O/ncu_gemm_res.log 2>&1; tail -2 $O/ncu_gemm_res.log
ls -la $O/*.ncu-rep
echo DONE
