"""Algorithmic FLOP counts of the denoiser (GEMM = 2 M N K, attention = 4 L Lk d per head; norms / elementwise excluded),
used by bench.py for the roofline figures (SDXL @128x128 latents = 6.761 TFLOP forward per sample)."""
from __future__ import annotations


def unet_forward_flops(cfg: dict, H: int, W: int, ctx_len: int = 77) -> dict:
    """Per-sample forward FLOPs by class for a UNet2DConditionModel config at H x W latents."""
    boc = tuple(cfg["block_out_channels"])
    n = len(boc)
    tl = cfg["transformer_layers_per_block"]
    tl = (tl,) * n if isinstance(tl, int) else tuple(tl)
    hd = cfg["attention_head_dim"]
    hd = (hd,) * n if isinstance(hd, int) else tuple(hd)
    cross = cfg["cross_attention_dim"]
    temb = boc[0] * 4
    f = dict(linear=0.0, conv3x3=0.0, conv1x1=0.0, attn=0.0)

    def conv3(hw, cin, cout):
        f["conv3x3"] += 2.0 * hw * cout * 9 * cin

    def resnet(hw, cin, cout):
        conv3(hw, cin, cout)
        conv3(hw, cout, cout)
        f["linear"] += 2.0 * temb * cout
        if cin != cout:
            f["conv1x1"] += 2.0 * hw * cin * cout

    def t2d(hw, c, heads, depth):
        f["linear"] += 2 * 2.0 * hw * c * c  # proj_in, proj_out
        for _ in range(depth):
            f["linear"] += 4 * 2.0 * hw * c * c  # q, k, v, out (self)
            f["attn"] += 4.0 * hw * hw * c
            f["linear"] += 2 * 2.0 * hw * c * c + 2 * 2.0 * ctx_len * cross * c  # q, out / k, v (cross)
            f["attn"] += 4.0 * hw * ctx_len * c
            f["linear"] += 2.0 * hw * c * 8 * c + 2.0 * hw * 4 * c * c  # GEGLU proj, ff out

    hw = H * W
    conv3(hw, cfg["in_channels"], boc[0])
    f["linear"] += 2.0 * boc[0] * temb + 2.0 * temb * temb
    if cfg.get("addition_embed_type") == "text_time":
        f["linear"] += 2.0 * cfg["projection_class_embeddings_input_dim"] * temb + 2.0 * temb * temb
    out = boc[0]
    res = [hw]
    for i, t in enumerate(cfg["down_block_types"]):
        cin, out = out, boc[i]
        for j in range(cfg["layers_per_block"]):
            resnet(hw, cin if j == 0 else out, out)
            if t.startswith("CrossAttn"):
                t2d(hw, out, hd[i], tl[i])
        if i != n - 1:
            hw //= 4
            conv3(hw, out, out)
    resnet(hw, boc[-1], boc[-1])
    t2d(hw, boc[-1], hd[-1], tl[-1])
    resnet(hw, boc[-1], boc[-1])
    rb, rh, rt = boc[::-1], hd[::-1], tl[::-1]
    out = rb[0]
    for i, t in enumerate(cfg["up_block_types"]):
        prev, out = out, rb[i]
        cin = rb[min(i + 1, n - 1)]
        L = cfg["layers_per_block"] + 1
        for j in range(L):
            skip = cin if j == L - 1 else out
            rin = prev if j == 0 else out
            resnet(hw, rin + skip, out)
            if t.startswith("CrossAttn"):
                t2d(hw, out, rh[i], rt[i])
        if i != n - 1:
            hw *= 4
            conv3(hw, out, out)
    conv3(hw, boc[0], cfg["out_channels"])
    f["total"] = sum(f.values())
    return f
