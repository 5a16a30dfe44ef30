#!/bin/bash
O=gpurun_out/run28; mkdir -p $O
echo "== vae + sampling"; timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_sampling_gpu.py -m gpu -x -q -s 2>&1 | grep -E "vae\]|passed|failed|^E " | head -20
echo DONE
