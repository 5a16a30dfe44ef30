#!/bin/bash
# run 20: GroupNorm image groups; clean one-step launch list (cudaProfilerStart/Stop around one eager step)
O=gpurun_out/run20; mkdir -p $O
echo "== pytest gn + unet"; timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q -k "groupnorm or unet or step or lycoris" 2>&1 | tail -3
echo "== gn (graph-timed)"; UWU_BENCH_GRAPH=1 timeout 120 python tools/bench_kernels.py gn 2>&1 | grep "^gn" | tee $O/gn_groups.log
echo "== bench weak"; timeout 300 python bench.py --scaling weak --no-cpu-baseline > $O/bench_weak.json 2> $O/bench_weak.err; cut -c1-200 $O/bench_weak.json
echo "== ncu launch list (one eager step)"; UWU_PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --scaling weak --graph off --no-cpu-baseline > $O/ncu_list.log 2>&1; tail -1 $O/ncu_list.log | cut -c1-160; wc -l $O/launches.csv
echo DONE
