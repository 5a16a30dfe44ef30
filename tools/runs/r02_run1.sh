#!/bin/bash
# round-2 GPU run 1: validate phase A on the 1-CTA GEMM, baseline profile, then try the CTA-pair GEMM
cd "$(dirname "$0")/../.."
O=gpurun_out/run1; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
echo "== pytest (pair=0)"; UWU_GEMM_PAIR=0 timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_sdxl_parity_gpu.py > $O/pytest_pair0.log 2>&1; tail -5 $O/pytest_pair0.log
echo "== sdxl parity (pair=0)"; UWU_GEMM_PAIR=0 timeout 1200 python -m pytest tests/test_sdxl_parity_gpu.py -q -s > $O/sdxl_parity.log 2>&1; tail -15 $O/sdxl_parity.log
echo "== smoke"; UWU_GEMM_PAIR=0 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== breakdown (pair=0)"; UWU_GEMM_PAIR=0 timeout 600 python tools/step_breakdown.py > $O/breakdown_pair0.log 2>&1; head -30 $O/breakdown_pair0.log
echo "== bench weak (pair=0)"; UWU_GEMM_PAIR=0 timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline > $O/bench_weak_pair0.json 2> $O/bench_weak_pair0.err; cat $O/bench_weak_pair0.json | cut -c1-400
echo "== pair parity"
for c in kmajor bkn acol conv; do UWU_GEMM_PAIR=1 timeout 120 python tools/diag_gemm.py $c > $O/diag_pair1_$c.log 2>&1; echo "rc=$? $c"; tail -4 $O/diag_pair1_$c.log; done
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
echo "== perf12"
UWU_GEMM_PAIR=0 timeout 300 python tools/diag_gemm.py perf12 > $O/perf12_pair0.log 2>&1; cat $O/perf12_pair0.log
UWU_GEMM_PAIR=1 timeout 300 python tools/diag_gemm.py perf12 > $O/perf12_pair1.log 2>&1; cat $O/perf12_pair1.log
echo "== pytest kernels (pair=1)"; UWU_GEMM_PAIR=1 timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_sdxl_parity_gpu.py > $O/pytest_pair1.log 2>&1; tail -5 $O/pytest_pair1.log
echo "== bench weak (pair=1)"; UWU_GEMM_PAIR=1 timeout 900 python bench.py --steps 5 --warmup 3 --scaling weak --no-cpu-baseline > $O/bench_weak_pair1.json 2> $O/bench_weak_pair1.err; cat $O/bench_weak_pair1.json | cut -c1-400
echo "== eager comparator"; timeout 900 python tools/bench_eager_cuda.py --steps 3 --warmup 2 > $O/eager.json 2> $O/eager.err; cat $O/eager.json; tail -3 $O/eager.err
echo "== bench strong default (pair=0, with cpu baseline)"; UWU_GEMM_PAIR=0 timeout 1200 python bench.py --steps 2 --warmup 3 > $O/bench_strong.json 2> $O/bench_strong.err; cat $O/bench_strong.json | cut -c1-600
echo DONE
