#!/bin/bash
# run 36: final bench lines of the last commit (mirrored LoKr route off = default, and on for the record)
O=gpurun_out/run36; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== bench (default)"; timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; cut -c1-200 $O/bench_final.json; tail -1 $O/bench_final.err
echo "== bench weak"; timeout 400 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo "== bench weak, mirrored route on"; UWU_LOKR_MIRROR=1 timeout 400 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak_mirror.json 2> $O/bench_weak_mirror.err; cut -c1-200 $O/bench_weak_mirror.json
echo "== unet + lokr tests"; timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -2
echo DONE
