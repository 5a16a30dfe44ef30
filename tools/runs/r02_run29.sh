#!/bin/bash
# run 29 (2 GPUs): final build under data parallelism, CUDA graphs on (default) and off (bucketed all-reduce overlapping the
# backward pass, i.e. NCCL kernels running next to the one-launch GroupNorm / persistent attention kernels)
O=gpurun_out/run29; mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== 2 GPUs strong graph on"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_2gpu_strong.json 2> $O/bench_2gpu_strong.err; cut -c1-200 $O/bench_2gpu_strong.json; tail -1 $O/bench_2gpu_strong.err
echo "== 2 GPUs weak graph off"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --scaling weak --graph off --no-cpu-baseline > $O/bench_2gpu_weak_eager.json 2> $O/bench_2gpu_weak_eager.err; cut -c1-200 $O/bench_2gpu_weak_eager.json; tail -1 $O/bench_2gpu_weak_eager.err
echo "== ddp_check c2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/ddp_check.py c2 > $O/ddp_check_c2.log 2>&1; tail -1 $O/ddp_check_c2.log | cut -c1-300
echo DONE
