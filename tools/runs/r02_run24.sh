#!/bin/bash
# run 24: persistent attention backward
O=gpurun_out/run24; mkdir -p $O
echo "== pytest attention"; timeout 180 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -4
echo "== attn perf (graph-timed)"; UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py attn attn_cross 2>&1 | grep "attn_" | tee $O/attn.log
echo "== unet"; timeout 240 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -2
echo "== bench weak"; timeout 300 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo DONE
