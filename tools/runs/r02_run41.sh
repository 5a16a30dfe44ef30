#!/bin/bash
# run 41 (8 GPUs): final build, default bench (strong scaling: one micro-batch of 16 per GPU and optimizer step)
O=gpurun_out/run41; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_8gpu_strong.json 2> $O/bench_8gpu_strong.err; cut -c1-300 $O/bench_8gpu_strong.json; tail -2 $O/bench_8gpu_strong.err
echo DONE
