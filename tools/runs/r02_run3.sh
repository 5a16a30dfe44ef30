#!/bin/bash
# round-2 GPU run 3: elect-one MMA / TMA issue loops (GEMM + attention), wide conv tiles, fused GEGLU, CUDA graph, parity yardstick
cd "$(dirname "$0")/../.."
O=gpurun_out/run3; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_sdxl_parity_gpu.py > $O/pytest.log 2>&1; tail -12 $O/pytest.log
echo "== sdxl parity"; timeout 1200 python -m pytest tests/test_sdxl_parity_gpu.py -q -s > $O/sdxl_parity.log 2>&1; tail -8 $O/sdxl_parity.log; cp gpurun_out/sdxl_parity.json $O/ 2>/dev/null
echo "== perf12"; timeout 300 python tools/diag_gemm.py perf12 > $O/perf12.log 2>&1; cat $O/perf12.log
echo "== perf12 no wide"; UWU_GEMM_WIDE=0 timeout 300 python tools/diag_gemm.py perf12 2>&1 | grep conv
echo "== perfw"; timeout 300 python tools/diag_gemm.py perfw > $O/perfw.log 2>&1; cat $O/perfw.log
echo "== attention kernels"; timeout 300 python tools/bench_kernels.py attn > $O/attn.log 2>&1; tail -20 $O/attn.log
echo "== bench weak graph off"; timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline --graph off > $O/bench_weak_nograph.json 2> $O/bench_weak_nograph.err; cut -c1-200 $O/bench_weak_nograph.json; tail -3 $O/bench_weak_nograph.err
echo "== bench weak graph on"; timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline --graph on > $O/bench_weak_graph.json 2> $O/bench_weak_graph.err; cut -c1-200 $O/bench_weak_graph.json; tail -4 $O/bench_weak_graph.err
echo "== breakdown"; timeout 600 python tools/step_breakdown.py > $O/breakdown.log 2>&1; head -60 $O/breakdown.log
echo DONE
