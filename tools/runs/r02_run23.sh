#!/bin/bash
# run 23: CTA-pair kernel for the grouped-N / segmented-K GEMMs of the factored LoKr route
O=gpurun_out/run23; mkdir -p $O
echo "== pytest"; timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or lokr or conv" 2>&1 | tail -3
echo "== lokr_fact (pair)"; UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py lokr_fact 2>&1 | grep -v Warn | tee $O/lokr_fact_pair.log
echo "== unet"; timeout 240 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -2
echo "== bench weak"; timeout 300 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo DONE
