#!/bin/bash
# run 7: one-pass LoKr gradient kernel — parity, isolated perf, step effect
mkdir -p gpurun_out/run7
echo "== pytest lokr fused"
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "lokr_fused" 2>&1 | tail -15 | tee gpurun_out/run7/pytest_fused.log
echo "== bench lokr_fused"
timeout 300 python tools/bench_kernels.py lokr_fused 2>&1 | tail -14 | tee gpurun_out/run7/lokr_fused.log
echo "== unet tests"
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/run7/pytest.log
echo "== bench weak graph"
timeout 900 python bench.py --scaling weak --steps 4 --warmup 3 > gpurun_out/run7/bench_weak.json 2> gpurun_out/run7/bench_weak.err; tail -c 1500 gpurun_out/run7/bench_weak.json
echo "== bench weak, fused off"
UWU_LOKR_FUSED=0 timeout 900 python bench.py --scaling weak --steps 4 --warmup 3 > gpurun_out/run7/bench_weak_nofused.json 2> gpurun_out/run7/bench_weak_nofused.err; tail -c 600 gpurun_out/run7/bench_weak_nofused.json
echo DONE
