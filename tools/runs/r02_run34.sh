#!/bin/bash
O=gpurun_out/run34; mkdir -p $O
echo "== unet tests"; timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== bench strong (fold once per accumulation window)"; timeout 500 python bench.py --steps 3 --no-cpu-baseline > $O/bench_strong.json 2> $O/bench_strong.err; cut -c1-200 $O/bench_strong.json; tail -1 $O/bench_strong.err
echo "== bench strong graph off"; timeout 500 python bench.py --steps 2 --graph off --no-cpu-baseline > $O/bench_strong_eager.json 2> $O/bench_strong_eager.err; cut -c1-200 $O/bench_strong_eager.json; tail -1 $O/bench_strong_eager.err
echo "== bench weak"; timeout 300 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo DONE
