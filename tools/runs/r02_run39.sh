#!/bin/bash
# run 39 (2 GPUs): last commit under data parallelism, strong scaling (4 micro-batches per GPU: twin graphs, one exchange per window)
O=gpurun_out/run39; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== 2 GPUs strong graph on"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_2gpu_strong.json 2> $O/bench_2gpu_strong.err; cut -c1-200 $O/bench_2gpu_strong.json; tail -1 $O/bench_2gpu_strong.err
echo "== 2 GPUs strong graph off"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --graph off --no-cpu-baseline > $O/bench_2gpu_strong_eager.json 2> $O/bench_2gpu_strong_eager.err; cut -c1-200 $O/bench_2gpu_strong_eager.json; tail -1 $O/bench_2gpu_strong_eager.err
echo DONE
