"""Worker of tests/test_ddp_cpu.py: one rank of a world-size-2 `gloo` job on CPU.

Runs the PRODUCT host logic of the data-parallel step (DMTrainer.fit_step -> GradientBuckets: per-rank seeding, bucket
readiness from the hand-scheduled backward, averaged all-reduce, clip + optimizer on the reduced gradients) with the kernels
replaced by the test-only torch emulation (tests/fake_ops.py), and writes what the test asserts on to a file.

    python tests/ddp_worker.py <rank> <world> <port> <out.pt> <mode>      mode: lycoris | full | dit
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def main():
    rank, world, port, out_path, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_ops

    fake_ops.install(_Patch())
    from conftest import LYCORIS_CFG, LYCORIS_PRESET
    from oracle import unet_oracle as U
    from uwudiff_b200.data import DummyDataset
    from uwudiff_b200 import config as ucfg
    from uwudiff_b200.trainer import DMTrainer

    ucfg.use_synthetic_conditioning(True)  # no pretrained text-encoder weights offline (explicit opt-in)

    torch.manual_seed(40 + rank)  # DIFFERENT initial weights per rank: GradientBuckets must broadcast rank 0's replica
    cfgd = U.tiny_config()
    tr = DMTrainer(
        model_config={"unet": {"_target_": "duwu.modules.unet_patch.UNet2DFromScratch.from_config", "config": dict(cfgd)},
                      "te": {"_target_": "duwu.modules.text_encoders.ConcatTextEncoders", "hidden_dim": 128, "pooled_dim": 64},
                      "vae": None},
        lycoris_config={"preset": LYCORIS_PRESET, "config": LYCORIS_CFG} if mode in ("lycoris", "accum") else None,
        lr=1e-3, optimizer="torch.optim.SGD", opt_config={}, use_warm_up=False, lr_scheduler=None,
        loss_config={"_target_": "duwu.loss.DiffusionLoss",
                     "scheduler": {"_target_": "diffusers.EulerDiscreteScheduler.from_pretrained",
                                   "pretrained_model_name_or_path": "stabilityai/stable-diffusion-xl-base-1.0",
                                   "subfolder": "scheduler"}},
        device="cpu")
    if mode in ("lycoris", "accum"):  # non-trivial adapter state so every gradient is non-zero (per-rank, overwritten by the broadcast)
        g = torch.Generator().manual_seed(1 + rank)
        tr.lycoris_model.flat_params.copy_(torch.randn(tr.lycoris_model.flat_params.shape, generator=g) * 0.05)
    if mode == "accum":
        return accum_mode(tr, rank, world, out_path)
    fit = tr.setup_fit(gradient_clip_val=None, seed=1215, n_buckets=3)
    buckets = fit["buckets"]
    assert buckets is not None and buckets.world == world
    assert tr.loss.seed == 1215 + rank  # pl.seed_everything(seed + global_rank), test_scripts/test_train.py:68-69
    params = list(tr.lycoris_model.parameters()) if mode in ("lycoris", "accum") else [p for p in tr.unet.parameters() if p.requires_grad]
    before = torch.cat([p.detach().reshape(-1).clone() for p in params])
    frozen = torch.cat([p.detach().reshape(-1)[:64].clone() for p in tr.unet.parameters()])  # frozen base synced as well
    torch.manual_seed(100 + rank)  # different data per rank
    ds = DummyDataset(sample_size=[4, 16, 16], n_samples=2)
    batch = ds.collate([ds[0], ds[1]])

    # local (un-reduced) gradient of this rank for the same draws: run the step once with the exchange disabled
    seen = {}
    orig_reduce = buckets._reduce

    def spy(view):
        seen.setdefault("local", []).append((view.data_ptr(), view.detach().clone()))
        orig_reduce(view)
        seen.setdefault("reduced", []).append(view.detach().clone())

    buckets._reduce = spy
    out = tr.fit_step(batch, 0)
    after = torch.cat([p.detach().reshape(-1).clone() for p in params])
    ds2 = DummyDataset(sample_size=[4, 16, 16], n_samples=2)
    for i in range(2):  # two more steps on different data per rank: replicas must stay identical
        buckets._reduce = orig_reduce
        tr.fit_step(ds2.collate([ds2[0], ds2[1]]), i + 1)
    after3 = torch.cat([p.detach().reshape(-1).clone() for p in params])
    flat0 = buckets.flat.data_ptr()
    local = torch.zeros_like(buckets.flat)
    reduced = torch.zeros_like(buckets.flat)
    for (ptr, loc), red in zip(seen["local"], seen["reduced"]):
        off = (ptr - flat0) // 4
        local[off:off + loc.numel()] = loc
        reduced[off:off + red.numel()] = red
    torch.save({"loss": float(out["loss"]), "before": before, "after": after, "local": local, "reduced": reduced,
                "n_calls": len(seen["local"]), "reduced_elems": buckets.reduced_elems, "n": buckets.flat.numel(),
                "t": out["aux_output"].timesteps.clone(), "frozen": frozen, "after3": after3}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def accum_mode(tr, rank, world, out_path):
    """configs[4] on fewer GPUs (bench.py --scaling strong): accumulate_grad_batches = 2 under data parallelism.  The first
    micro-batch only accumulates locally (no exchange, parameters untouched, adapter fold skipped on the second forward); the
    exchange runs once, during the last micro-batch's backward, on the MEAN of the micro-batch gradients."""
    from uwudiff_b200.data import DummyDataset

    fit = tr.setup_fit(gradient_clip_val=None, seed=1215, n_buckets=3, accumulate_grad_batches=2)
    buckets = fit["buckets"]
    params = list(tr.lycoris_model.parameters())
    before = torch.cat([p.detach().reshape(-1).clone() for p in params])
    torch.manual_seed(100 + rank)
    ds = DummyDataset(sample_size=[4, 16, 16], n_samples=4)
    mb = [ds.collate([ds[0], ds[1]]), ds.collate([ds[2], ds[3]])]
    seen = {"local": [], "reduced": []}
    orig_reduce = buckets._reduce

    def spy(view):
        seen["local"].append((view.data_ptr(), view.detach().clone()))
        orig_reduce(view)
        seen["reduced"].append(view.detach().clone())

    buckets._reduce = spy
    tr.fit_step(mb[0], 0)
    calls_mb1 = len(seen["local"])
    mid = torch.cat([p.detach().reshape(-1).clone() for p in params])
    g_mb1 = tr.lycoris_model.flat_grads.detach().clone()  # (g0 / 2), local
    tr.fit_step(mb[1], 1)
    after = torch.cat([p.detach().reshape(-1).clone() for p in params])
    flat0 = buckets.flat.data_ptr()
    local, reduced = torch.zeros_like(buckets.flat), torch.zeros_like(buckets.flat)
    for (ptr, loc), red in zip(seen["local"], seen["reduced"]):
        off = (ptr - flat0) // 4
        local[off:off + loc.numel()] = loc
        reduced[off:off + red.numel()] = red
    torch.save({"calls_mb1": calls_mb1, "n_calls": len(seen["local"]), "before": before, "mid": mid, "after": after, "local": local,
                "reduced": reduced, "g_mb1": g_mb1, "reduced_elems": buckets.reduced_elems, "n": buckets.flat.numel(),
                "global_step": tr.global_step}, out_path)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
