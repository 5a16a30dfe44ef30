"""Frozen conditioning stack of the reference: `ConcatTextEncoders` over CLIP text towers
(src/duwu/modules/text_encoders.py:38-264), executed by the sm_100a kernels of libuwu_b200.so.

The reference wraps `transformers.CLIPTextModel` objects (configs/demo_training_lycoris.yaml:91-110: SDXL's `text_encoder`
= CLIP ViT-L/14 text tower, `text_encoder_2` = OpenCLIP bigG text tower) and calls
`text_model(input_ids, attention_mask=..., output_hidden_states=True, return_dict=False)`.  `CLIPTextModel` here is a drop-in
for that call: same constructor config keys, same parameter names (a transformers state dict loads unchanged), same return
tuple `(last_hidden_state, pooler_output, hidden_states)`; the arithmetic runs on the tcgen05 GEMM (fused q|k|v projection,
bias / residual epilogues), the LayerNorm kernel and a causal + padding-masked attention kernel.  Forward only: the towers
are frozen (`to_freeze: true`).  There is no CPU path.

Pretrained weights live on the HF hub (unreachable offline): `from_pretrained` loads a local directory when given one, and
otherwise builds the architecture from the embedded public config with RANDOM weights and says so loudly.
"""
from __future__ import annotations

import json
import os
import types
import warnings
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .data import BaseTextEncoder
from .unet import LayerNorm, Linear

BF16 = torch.bfloat16

# public config.json constants of the SDXL text encoders (stabilityai/stable-diffusion-xl-base-1.0)
CLIP_L_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                     max_position_embeddings=77, hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2, pad_token_id=1,
                     bos_token_id=0)
CLIP_BIGG_CONFIG = dict(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                        max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5, eos_token_id=2, pad_token_id=1,
                        bos_token_id=0)
KNOWN_TEXT_CONFIGS = {("stabilityai/stable-diffusion-xl-base-1.0", "text_encoder"): CLIP_L_CONFIG,
                      ("stabilityai/stable-diffusion-xl-base-1.0", "text_encoder_2"): CLIP_BIGG_CONFIG,
                      ("openai/clip-vit-large-patch14", None): CLIP_L_CONFIG,
                      ("laion/CLIP-ViT-bigG-14-laion2B-39B-b160k", None): CLIP_BIGG_CONFIG}
_ACT = {"quick_gelu": ops.EW_QUICK_GELU, "gelu": ops.EW_GELU_ERF, "gelu_pytorch_tanh": ops.EW_GELU_TANH}


class _KernelModule(nn.Module):
    """fp32 master parameters, bf16 compute: floating dtype casts (`_load_config_.precision: torch.float16`, `.half()`) only
    record the dtype the outputs are returned in; the parameters stay fp32 on whatever device they are moved to."""

    out_dtype = torch.float32

    def _apply(self, fn, *a, **k):
        probe = fn(torch.zeros((), dtype=torch.float32))
        if probe.dtype != torch.float32:  # a dtype cast: keep masters, remember the requested output precision
            self.out_dtype = probe.dtype

            def fn_dev(t, _fn=fn):
                r = _fn(t)
                return r.to(t.dtype) if r.is_floating_point() and t.is_floating_point() else r

            r = super()._apply(fn_dev, *a, **k)
        else:
            r = super()._apply(fn, *a, **k)
        for m in self.modules():
            if hasattr(m, "drop_cache"):
                m.drop_cache()
        return r

    @property
    def dtype(self):
        return self.out_dtype

    @property
    def device(self):
        return next(self.parameters()).device


class CLIPAttention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads, self.dim_head, self.dim = heads, dim // heads, dim
        if self.dim_head > 64 or self.dim_head % 8:
            raise NotImplementedError(f"uwudiff_b200: CLIP head_dim {self.dim_head} unsupported (<= 64, multiple of 8)")
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (Linear(dim, dim) for _ in range(4))
        self._fused = None

    def drop_cache(self):
        self._fused = None

    def _qkv(self):
        if self._fused is None:
            D = self.dim
            w = torch.empty((3 * D, D), device=self.q_proj.weight.device, dtype=BF16)
            for i, m in enumerate((self.q_proj, self.k_proj, self.v_proj)):
                m.w16(dst=w[i * D:(i + 1) * D])
            b = torch.cat([self.q_proj.bias, self.k_proj.bias, self.v_proj.bias]).float().contiguous()
            self._fused = (w, b)
        return self._fused

    def fwd(self, n, x_res, B, L, key_mask):
        D, M = self.dim, B * L
        w, b = self._qkv()
        qkv = ops.gemm(n, w, M, 3 * D, D, bias=b)
        o = ops.attn_fwd_masked(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, self.heads, L, causal=True, key_mask=key_mask,
                                head_dim=self.dim_head)
        return self.out_proj.fwd(o, M, residual=x_res)


class CLIPMLP(nn.Module):
    def __init__(self, dim: int, inner: int, act: str):
        super().__init__()
        if act not in _ACT:
            raise NotImplementedError(f"uwudiff_b200: CLIP hidden_act '{act}' is not built ({sorted(_ACT)})")
        self.fc1, self.fc2, self.act = Linear(dim, inner), Linear(inner, dim), _ACT[act]

    def fwd(self, n, x_res, M):
        h = self.fc1.fwd(n, M)
        return self.fc2.fwd(ops.elementwise(h, None, self.act), M, residual=x_res)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, dim: int, heads: int, inner: int, act: str, eps: float):
        super().__init__()
        self.self_attn = CLIPAttention(dim, heads)
        self.layer_norm1 = LayerNorm(dim, eps=eps)
        self.mlp = CLIPMLP(dim, inner, act)
        self.layer_norm2 = LayerNorm(dim, eps=eps)

    def fwd(self, x, B, L, key_mask):
        n1, _ = ops.layernorm_fwd(x, self.layer_norm1.weight, self.layer_norm1.bias, self.layer_norm1.eps, want_stats=False)
        x = self.self_attn.fwd(n1, x, B, L, key_mask)
        n2, _ = ops.layernorm_fwd(x, self.layer_norm2.weight, self.layer_norm2.bias, self.layer_norm2.eps, want_stats=False)
        return self.mlp.fwd(n2, x, B * L)


class _Embeddings(nn.Module):
    def __init__(self, vocab: int, dim: int, max_pos: int):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, dim)
        self.position_embedding = nn.Embedding(max_pos, dim)


class _Encoder(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)


class CLIPTextTransformer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.embeddings = _Embeddings(c.vocab_size, c.hidden_size, c.max_position_embeddings)
        self.encoder = _Encoder([CLIPEncoderLayer(c.hidden_size, c.num_attention_heads, c.intermediate_size, c.hidden_act,
                                                  c.layer_norm_eps) for _ in range(c.num_hidden_layers)])
        self.final_layer_norm = LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.eos_token_id = c.eos_token_id
        self._pos_cache = {}

    def final_layer_norm_tokens(self, x2d: torch.Tensor) -> torch.Tensor:
        f = self.final_layer_norm
        return ops.layernorm_fwd(x2d, f.weight, f.bias, f.eps, want_stats=False)[0]


class CLIPTextModel(_KernelModule):
    """Drop-in for `transformers.CLIPTextModel` as `ConcatTextEncoders.forward` calls it (text_encoders.py:167-173)."""

    def __init__(self, config=None, **kw):
        super().__init__()
        c = dict(CLIP_L_CONFIG)
        if config is not None:
            c.update(config if isinstance(config, dict) else {k: getattr(config, k) for k in c if hasattr(config, k)})
        c.update(kw)
        self.config = types.SimpleNamespace(**c)
        self.text_model = CLIPTextTransformer(self.config)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **_):
        root = os.path.join(pretrained_model_name_or_path, subfolder or "")
        cfg_file = os.path.join(root, "config.json")
        if os.path.exists(cfg_file):
            with open(cfg_file) as f:
                raw = json.load(f)
            model = cls({k: raw[k] for k in CLIP_L_CONFIG if k in raw})
            for fn in ("model.safetensors", "pytorch_model.bin"):
                path = os.path.join(root, fn)
                if os.path.exists(path):
                    if fn.endswith(".safetensors"):
                        from safetensors.torch import load_file

                        sd = load_file(path)
                    else:
                        sd = torch.load(path, map_location="cpu")
                    sd = {k: v for k, v in sd.items() if "position_ids" not in k}
                    model.load_state_dict(sd)
                    return model
            warnings.warn(f"uwudiff_b200: no weight file under {root}: CLIP text tower initialised RANDOMLY")
            return model
        key = (pretrained_model_name_or_path, subfolder)
        if key in KNOWN_TEXT_CONFIGS:
            warnings.warn(f"uwudiff_b200: '{pretrained_model_name_or_path}' ({subfolder}) is not a local directory and the HF hub "
                          "is unreachable: CLIP text tower built from the embedded config with RANDOM weights")
            return cls(KNOWN_TEXT_CONFIGS[key])
        raise OSError(f"CLIP text model '{pretrained_model_name_or_path}' ({subfolder}): neither a local directory nor an "
                      f"embedded config (known: {sorted(str(k) for k in KNOWN_TEXT_CONFIGS)})")

    def load_state_dict(self, sd, *a, **k):
        sd = {kk: v for kk, v in sd.items() if not kk.endswith("position_ids")}
        r = super().load_state_dict(sd, *a, **k)
        for m in self.modules():
            if hasattr(m, "drop_cache"):
                m.drop_cache()
        return r

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, output_hidden_states: bool = False, return_dict: bool = False, **_):
        tm = self.text_model
        dev = tm.embeddings.token_embedding.weight.device
        ops._req_cuda(tm.embeddings.token_embedding.weight)
        input_ids = input_ids.to(dev)
        B, L = input_ids.shape
        D = self.config.hidden_size
        ids = input_ids.reshape(-1).to(torch.int64).contiguous()
        x = ops.embed_gather(tm.embeddings.token_embedding.weight.detach(), ids)              # [B*L, D] bf16
        pos = tm._pos_cache.get((B, L))
        if pos is None:
            p = ops.embed_gather(tm.embeddings.position_embedding.weight.detach(), torch.arange(L, device=dev).repeat(B))
            tm._pos_cache = {(B, L): p}
            pos = p
        x = ops.elementwise(x, pos, ops.EW_ADD)
        key_mask = None if attention_mask is None else attention_mask.to(dev)
        hidden = [x]
        for layer in tm.encoder.layers:
            x = layer.fwd(x, B, L, key_mask)
            hidden.append(x)
        last = tm.final_layer_norm_tokens(x)
        if self.config.eos_token_id == 2:  # legacy configs: the EOT token is the highest id of the sequence
            eos = input_ids.to(torch.int).argmax(dim=-1)
        else:
            eos = (input_ids.to(torch.int) == self.config.eos_token_id).int().argmax(dim=-1)
        dt = self.out_dtype
        last3 = last.view(B, L, D)
        pooled = last3[torch.arange(B, device=dev), eos].to(dt)
        out = (last3.to(dt), pooled)
        if output_hidden_states:
            out = out + (tuple(h.view(B, L, D).to(dt) for h in hidden),)
        if return_dict:
            return types.SimpleNamespace(last_hidden_state=out[0], pooler_output=out[1], hidden_states=out[2] if output_hidden_states else None)
        return out


def remove_none(xs):
    return [x for x in xs if x is not None]


@dataclass
class TextModelExtraConfig:
    concat_bucket: int = 0
    use_pooled: bool = False
    layer_idx: int = -1
    need_mask: bool = False
    disable_autocast: bool = False


class ConcatTextEncoders(BaseTextEncoder):
    """src/duwu/modules/text_encoders.py:38-264: several tokenizer / text-model pairs, hidden states of layer `layer_idx`
    concatenated along channels inside a bucket and along tokens across buckets, pooled outputs concatenated, optional zeroing
    of padding positions and mask output.  `transformers.CLIPTextModel.from_pretrained` targets resolve to the kernel-backed
    `CLIPTextModel` above through uwudiff_b200.config."""

    def __init__(self, tokenizers: List = [], text_model_and_configs: List = [], zero_for_padding: bool = True, max_length: int = 256,
                 use_normed_ctx: bool = False):
        super().__init__()
        from .config import load_any

        self.tokenizers = []
        for tok in tokenizers:
            if isinstance(tok, str):
                from transformers import AutoTokenizer

                tok = AutoTokenizer.from_pretrained(tok)  # needs the tokenizer files locally (the hub is unreachable offline)
            if not tok.pad_token:
                tok.pad_token = tok.eos_token
            if tok.model_max_length > max_length:
                tok.model_max_length = max_length
            self.tokenizers.append(tok)
        models, self.configs, self.max_bucket = [], [], 0
        self.use_normed_ctx = use_normed_ctx
        for text_model, extra in text_model_and_configs:
            models.append(load_any(text_model))
            if not isinstance(extra, TextModelExtraConfig):
                extra = TextModelExtraConfig(**extra)
            self.configs.append(extra)
            self.max_bucket = max(self.max_bucket, extra.concat_bucket)
        self.text_models = nn.ModuleList(models)
        self.zero_for_padding = zero_for_padding
        self.out_dtype = torch.float32

    def _apply(self, fn, *a, **k):
        probe = fn(torch.zeros((), dtype=torch.float32))
        if probe.dtype != torch.float32:
            self.out_dtype = probe.dtype
        return super()._apply(fn, *a, **k)

    @property
    def dtype(self):
        return self.out_dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def tokenize(self, text, **kwargs):
        return [tok(text, **kwargs, return_tensors="pt") for tok in self.tokenizers]

    def encode(self, text, **kwargs):
        return self.forward(self.tokenize(text, **kwargs))

    @torch.no_grad()
    def forward(self, tokenizers_outputs, batch_size: Optional[int] = None):
        nb = self.max_bucket + 1
        attn_masks = [None] * nb
        embs, normed, pooled = [[] for _ in range(nb)], [[] for _ in range(nb)], [[] for _ in range(nb)]
        dev = self.device
        for tokens, text_model, cfg in zip(tokenizers_outputs, self.text_models, self.configs):
            b = cfg.concat_bucket
            input_ids = tokens["input_ids"].to(dev)
            attn_mask = tokens["attention_mask"].to(dev)
            if attn_masks[b] is None and cfg.need_mask:
                attn_masks[b] = attn_mask
            normed_e, pooled_e, *rest = text_model(input_ids, attention_mask=attn_mask, output_hidden_states=True, return_dict=False)
            emb = rest[-1][cfg.layer_idx]
            if isinstance(text_model, CLIPTextModel):  # "SD1/SD2 need this" (:181-182): final LayerNorm of the chosen layer
                Bq, L, D = emb.shape
                normed_e = text_model.text_model.final_layer_norm_tokens(emb.reshape(Bq * L, D).to(BF16).contiguous()).view(Bq, L, D)
            emb, normed_e, pooled_e = emb.to(self.dtype), normed_e.to(self.dtype), pooled_e.to(self.dtype)
            if self.zero_for_padding:
                m = attn_mask.unsqueeze(-1)
                emb, normed_e = emb * m, normed_e * m
            embs[b].append(emb)
            normed[b].append(normed_e)
            if cfg.use_pooled and pooled_e is not None:
                pooled[b].append(pooled_e)
        for i in range(nb):
            if not embs[i]:
                embs[i] = normed[i] = pooled[i] = None
                continue
            embs[i] = torch.cat(embs[i], dim=-1)
            normed[i] = torch.cat(normed[i], dim=-1)
            pooled[i] = torch.cat(pooled[i], dim=-1) if pooled[i] else None
        max_dim = max(e.size(-1) for e in embs if e is not None)
        for lst in (embs, normed):
            for i, e in enumerate(lst):
                if e is not None and e.size(-1) < max_dim:
                    lst[i] = torch.nn.functional.pad(e, (0, max_dim - e.size(-1)))
        if any(m is not None for m in attn_masks):
            for i, e in enumerate(embs):
                if e is not None and attn_masks[i] is None:
                    attn_masks[i] = torch.ones(e.size(0), e.size(1), device=e.device).long()
            masks = torch.cat(remove_none(attn_masks), dim=1)
        else:
            masks = None
        pooled_out = torch.cat(remove_none(pooled), dim=-1) if any(p is not None for p in pooled) else None
        return torch.cat(remove_none(embs), dim=1), torch.cat(remove_none(normed), dim=1), pooled_out, masks
