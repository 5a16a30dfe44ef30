#!/bin/bash
# run 15 (8 GPUs): data-parallel bench in both scaling modes
O=gpurun_out/run15; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
echo "== bench 8 GPUs strong"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_8gpu_strong.json 2> $O/bench_8gpu_strong.err; cut -c1-300 $O/bench_8gpu_strong.json; tail -2 $O/bench_8gpu_strong.err
echo "== bench 8 GPUs weak"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --scaling weak > $O/bench_8gpu_weak.json 2> $O/bench_8gpu_weak.err; cut -c1-300 $O/bench_8gpu_weak.json; tail -2 $O/bench_8gpu_weak.err
echo "== bench 4 GPUs strong"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 > $O/bench_4gpu_strong.json 2> $O/bench_4gpu_strong.err; cut -c1-300 $O/bench_4gpu_strong.json; tail -2 $O/bench_4gpu_strong.err
echo "== ddp_check c4 8 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/ddp_check.py c4 > $O/ddp_check_c4_8gpu.log 2>&1; tail -1 $O/ddp_check_c4_8gpu.log | cut -c1-400
echo DONE
