#!/bin/bash
# run 21: residual-TMA test over many tiles; DiT 1-GPU batch 32 (denominator of the 8-GPU efficiency); ncu of the short-K GEMM
O=gpurun_out/run21; mkdir -p $O
echo "== pytest"; timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "residual_prefetched or epilogues" 2>&1 | tail -3
echo "== c4 B32"; timeout 200 python tools/bench_configs.py c4 --batch 32 --steps 10 --warmup 3 > $O/bench_c4_b32.json 2>$O/c4.err; cut -c1-200 $O/bench_c4_b32.json
echo "== ncu gemm_res"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 1 --launch-count 1 -o $O/gemm_res -f python tools/profile_one.py gemm_res > $O/ncu.log 2>&1; tail -1 $O/ncu.log
echo DONE
