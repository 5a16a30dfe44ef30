#!/bin/bash
# round-2 GPU run 2: full test suite (pair GEMM default, CUDA graph, accumulation, parity with yardstick), bench graph on/off, ncu
cd "$(dirname "$0")/../.."
O=gpurun_out/run2; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_sdxl_parity_gpu.py > $O/pytest.log 2>&1; tail -15 $O/pytest.log
echo "== sdxl parity"; timeout 1200 python -m pytest tests/test_sdxl_parity_gpu.py -q -s > $O/sdxl_parity.log 2>&1; tail -12 $O/sdxl_parity.log; cp gpurun_out/sdxl_parity.json $O/ 2>/dev/null
echo "== perfw pair=1 / 2"
UWU_GEMM_PAIR=1 timeout 300 python tools/diag_gemm.py perfw > $O/perfw_pair1.log 2>&1; cat $O/perfw_pair1.log
UWU_GEMM_PAIR=2 timeout 300 python tools/diag_gemm.py perfw > $O/perfw_pair2.log 2>&1; cat $O/perfw_pair2.log
echo "== bench weak graph off, unfused GEGLU"; UWU_GEGLU_FUSE=0 timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline --graph off > $O/bench_weak_nograph_nofuse.json 2> $O/bench_weak_nograph_nofuse.err; cut -c1-200 $O/bench_weak_nograph_nofuse.json; tail -3 $O/bench_weak_nograph_nofuse.err
echo "== bench weak graph off"; timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline --graph off > $O/bench_weak_nograph.json 2> $O/bench_weak_nograph.err; cut -c1-300 $O/bench_weak_nograph.json; tail -3 $O/bench_weak_nograph.err
echo "== bench weak graph on"; timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline --graph on > $O/bench_weak_graph.json 2> $O/bench_weak_graph.err; cut -c1-300 $O/bench_weak_graph.json; tail -5 $O/bench_weak_graph.err
echo "== breakdown"; timeout 600 python tools/step_breakdown.py > $O/breakdown.log 2>&1; head -24 $O/breakdown.log
echo "== ncu conv320"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 2 --launch-count 1 -o $O/conv320 -f python tools/profile_one.py conv320 > $O/ncu_conv320.log 2>&1; tail -2 $O/ncu_conv320.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 5 --launch-count 1 -o $O/gemm_geglu -f python tools/profile_one.py gemm_lin > $O/ncu_gemm.log 2>&1; tail -2 $O/ncu_gemm.log
ls -la $O/*.ncu-rep
echo DONE
