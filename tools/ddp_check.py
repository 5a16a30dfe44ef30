"""Multi-GPU sanity check of the data-parallel step for the full-fine-tuning configurations (NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py c4

Each rank runs `DMTrainer.fit_step` on its own synthetic batch (seed + rank); after every step the parameters of all ranks
must be bit-identical (same reduced gradients, same optimizer), the loss finite, and every gradient element must have been
exchanged exactly once.  Prints one JSON line on rank 0 (samples/s over all ranks, max |param difference| across ranks).
"""
import json
import os

os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")  # synthetic text-encoder outputs (no weights offline)
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist

import bench_configs
from uwudiff_b200 import config as ucfg


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "c4"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = {"c1": 4, "c2": 8, "c4": 32, "latent": 4}[kind]
    conf, _, shape, label = bench_configs.make_conf(kind, B)
    torch.manual_seed(0)  # identical replicas
    trainer = ucfg.instantiate_any(conf["trainer"])
    fit = trainer.setup_fit(gradient_clip_val=1.0, seed=1215)
    assert fit["buckets"] is not None
    torch.manual_seed(100 + rank)
    x = torch.randn((B, *shape), device=dev)
    cond = ({"class_labels": torch.randint(0, 1000, (B,), device=dev)} if kind == "c4"
            else {"time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B, device=dev)})
    batch = (x, ["DUMMY TEST"] * B, [], cond, {})
    losses, worst = [], 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(6):
        if i == 3:
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
        out = trainer.fit_step(batch, i)
        losses.append(float(out["loss"].detach()))
        assert fit["buckets"].reduced_elems == fit["buckets"].flat.numel()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    # replicas identical?
    for p in trainer.unet.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, 0)
        worst = max(worst, float((ref - p.detach()).abs().max()))
    t = torch.tensor([ms, worst, float(all(map(lambda v: v == v and abs(v) < 1e4, losses)))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": label, "world": world, "per_gpu_batch": B, "ms_per_step": t[0].item(),
                          "samples_per_s": B * world / (t[0].item() * 1e-3), "max_param_diff_across_ranks": t[1].item(),
                          "losses_rank0": losses, "finite": bool(t[2].item())}))
    assert t[1].item() == 0.0, "replicas diverged"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
