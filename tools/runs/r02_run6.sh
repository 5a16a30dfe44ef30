#!/bin/bash
# run 6: LN backward streaming kernel — parity + bandwidth; graph test re-check
mkdir -p gpurun_out/run6
echo "== pytest norm/graph"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q -k "norm or layernorm or graph or accumulation or dit or adaln" 2>&1 | tail -8 | tee gpurun_out/run6/pytest.log
echo "== dbg graph"
timeout 300 python tools/dbg/graph_bits.py 2>&1 | grep -v Warn | grep "graph\]" | cut -c1-200
echo "== bench ln (stream)"
timeout 300 python tools/bench_kernels.py ln 2>&1 | tail -12 | tee gpurun_out/run6/ln_stream.log
echo "== bench ln (old)"
UWU_LN_STREAM=0 timeout 300 python tools/bench_kernels.py ln 2>&1 | tail -12 | tee gpurun_out/run6/ln_old.log
echo DONE
