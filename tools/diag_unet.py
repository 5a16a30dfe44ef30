"""GPU diagnostic: kernel-backed UNet (+ LyCORIS) vs the fp32 CPU oracle on a tiny SDXL-shaped config.

    python tools/diag_unet.py [fwd|lyco|sdxl_mem]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import lycoris_oracle as LY
from oracle import unet_oracle as U
from uwudiff_b200 import lycoris as PL
from uwudiff_b200 import unet as P

PRESET = dict(enable_conv=False, target_module=["Transformer2DModel"], target_name=[],
              module_algo_map={"Attention": dict(algo="lokr", factor=64, full_matrix=True),
                               "FeedForward": dict(algo="lokr", factor=6, full_matrix=True)})
LYCFG = dict(linear_dim=4, linear_alpha=1, conv_dim=4, conv_alpha=1, algo="lora", use_tucker=True, train_norm=True)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def build(seed=0, B=2, HW=16, zero_init=False):
    torch.manual_seed(seed)
    cfg = U.tiny_config()
    o = U.UNet2DConditionModel(**cfg)
    if zero_init:
        o.init_weight()
    p = P.UNet2DFromScratch.from_config(cfg)
    p.load_state_dict(o.state_dict())
    p = p.cuda()
    x = torch.randn(B, 4, HW, HW)
    t = torch.randint(0, 1000, (B,))
    ctx = torch.randn(B, 77, cfg["cross_attention_dim"])
    ac = dict(text_embeds=torch.randn(B, 64), time_ids=torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B))
    return cfg, o, p, x, t, ctx, ac


def case_fwd():
    cfg, o, p, x, t, ctx, ac = build()
    with torch.no_grad():
        yo = o(x, t, encoder_hidden_states=ctx, added_cond_kwargs=ac)[0]
        yp = p(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda(), added_cond_kwargs={k: v.cuda() for k, v in ac.items()})[0]
    torch.cuda.synchronize()
    print(f"unet fwd tiny: rel={rel(yp, yo):.3e} ref_max={yo.abs().max():.3e}", flush=True)
    print("DONE fwd")


def case_lyco():
    cfg, o, p, x, t, ctx, ac = build()
    LY.LycorisNetwork.apply_preset(PRESET)
    PL.LycorisNetwork.apply_preset(PRESET)
    no = LY.create_lycoris(o, **LYCFG)
    # non-trivial adapter state so every delta matters
    g = torch.Generator().manual_seed(1)
    for prm in no.parameters():
        prm.data = torch.randn(prm.shape, generator=g) * 0.05
    npd = PL.create_lycoris(p, **LYCFG)
    npd.load_state_dict(no.state_dict())
    no.apply_to()
    npd.apply_to()
    o.requires_grad_(False)
    p.requires_grad_(False)
    gout = torch.randn(x.shape, generator=g)
    yo = o(x, t, encoder_hidden_states=ctx, added_cond_kwargs=ac)[0]
    yo.backward(gout)
    yp = p(x.cuda(), t.cuda(), encoder_hidden_states=ctx.cuda(), added_cond_kwargs={k: v.cuda() for k, v in ac.items()})[0]
    yp.backward(gout.cuda())
    torch.cuda.synchronize()
    print(f"unet+lycoris fwd: rel={rel(yp, yo):.3e}", flush=True)
    worst = {}
    po = dict(no.named_parameters())
    for name, prm in npd.named_parameters():
        r = rel(prm.grad, po[name].grad)
        kind = name.rsplit(".", 1)[-1] if "lora" not in name else name.split(".")[-2]
        if r > worst.get(kind, (0, ""))[0]:
            worst[kind] = (r, name)
    for k, (r, n) in sorted(worst.items()):
        print(f"worst grad rel [{k}] = {r:.3e} ({n}) ref_max={po[n].grad.abs().max():.3e}", flush=True)
    tot_o = torch.sqrt(sum((q.grad.float() ** 2).sum() for q in no.parameters()))
    tot_p = torch.sqrt(sum((q.grad.float() ** 2).sum() for q in npd.parameters()))
    print(f"grad norm oracle {tot_o:.6e} product {tot_p.item():.6e}", flush=True)
    print("DONE lyco")


def case_sdxl_mem():
    """One SDXL-size LyCORIS step at 128x128 latents: time and peak memory per batch size."""
    torch.manual_seed(0)
    with torch.device("cuda"):
        p = P.UNet2DFromScratch.from_config("stabilityai/stable-diffusion-xl-base-1.0", subfolder="unet")
    PL.LycorisNetwork.apply_preset(PRESET)
    net = PL.create_lycoris(p, **LYCFG)
    net.apply_to()
    p.requires_grad_(False)
    print(f"model on device: {torch.cuda.memory_allocated()/2**30:.1f} GiB", flush=True)
    for B in (1, 2, 4, 8, 16):
        x = torch.randn(B, 4, 128, 128, device="cuda")
        t = torch.randint(0, 1000, (B,), device="cuda")
        ctx = torch.randn(B, 77, 2048, device="cuda")
        ac = dict(text_embeds=torch.randn(B, 1280, device="cuda"),
                  time_ids=torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B, device="cuda"))
        for it in range(3):
            torch.cuda.reset_peak_memory_stats()
            torch.cuda.synchronize()
            t0 = time.time()
            y = p(x, t, encoder_hidden_states=ctx, added_cond_kwargs=ac)[0]
            torch.cuda.synchronize()
            t1 = time.time()
            y.backward(torch.randn_like(y))
            torch.cuda.synchronize()
            t2 = time.time()
        print(f"B={B}: fwd {1e3*(t1-t0):.1f} ms bwd {1e3*(t2-t1):.1f} ms peak {torch.cuda.max_memory_allocated()/2**30:.1f} GiB "
              f"finite={bool(torch.isfinite(y).all())} gradnorm={net.flat_grads.norm().item():.3e}", flush=True)
        net.zero_grad()
    print("DONE sdxl_mem")


if __name__ == "__main__":
    for c in sys.argv[1:]:
        try:
            globals()["case_" + c]()
        except Exception as e:
            import traceback

            traceback.print_exc()
            print(f"CASE {c} FAILED: {e}", flush=True)
