#!/bin/bash
O=gpurun_out/run33; mkdir -p $O
echo "== fold tests"; timeout 240 python -m pytest tests/test_unet_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "fold or lycoris or loha or lora" 2>&1 | tail -2
echo "== fold timing"; timeout 200 python tools/bench_fold.py 2>&1 | grep fold_all | tee $O/fold.log
echo DONE
