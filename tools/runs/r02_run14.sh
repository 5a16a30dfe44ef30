#!/bin/bash
# run 14 (2 GPUs): data-parallel bench in both scaling modes, DiT ddp check
O=gpurun_out/run14; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L
echo "== bench 2 GPUs strong"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu_strong.json 2> $O/bench_2gpu_strong.err; cut -c1-400 $O/bench_2gpu_strong.json; tail -3 $O/bench_2gpu_strong.err
echo "== bench 2 GPUs weak"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --scaling weak > $O/bench_2gpu_weak.json 2> $O/bench_2gpu_weak.err; cut -c1-400 $O/bench_2gpu_weak.json; tail -3 $O/bench_2gpu_weak.err
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/bench_ref_2.json 2> $O/bench_ref_2.err; cut -c1-300 $O/bench_ref_2.json; tail -2 $O/bench_ref_2.err
echo "== ddp_check c4"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/ddp_check.py c4 > $O/ddp_check_c4.log 2>&1; tail -2 $O/ddp_check_c4.log | cut -c1-400
echo DONE
