"""The reference's batch contract, kept AS IS: `UwUBaseDataset.collate`, `DummyDataset` and `TrainDataModule` below are a
near-verbatim mirror of src/duwu/data/base.py:9-95 (about 45 lines: the 5-tuple layout, the caption / `add_time_ids`
constants and the method names are the drop-in API itself and leave nothing to redesign; they are off the hot path).
`DummyDataset` (pre-generated `torch.randn(sample_size)`
samples, fixed caption, fixed `add_time_ids`), `collate` -> the 5-tuple `(samples, captions, tokenizer_outputs,
{"time_ids": ...}, {})` that `DMTrainer.get_latent_and_conditioning` unpacks (src/duwu/trainer/trainer.py:233-261),
and `TrainDataModule`.  Lightning is not installable here, so `TrainDataModule` is a plain object with the same methods.

`SyntheticTextEncoders` / `SyntheticVAE` stand in for the frozen conditioning stack (`ConcatTextEncoders` over two CLIP
text models, `AutoencoderKL`): their weights live on the HF hub, which is unreachable, and they run under `no_grad`
outside the denoiser (≈1.5 % of the step).  They return deterministic synthetic
tensors of the real shapes/dtypes so the shipped YAMLs instantiate and step end to end.
"""
from __future__ import annotations

import types
from typing import List, Sequence

import torch
import torch.utils.data as Data

from .config import load_any


class UwUBaseDataset(Data.Dataset):
    @staticmethod
    def collate(batch):
        samples = torch.stack([x["sample"] for x in batch])
        caption = [x["caption"] for x in batch]
        tokenizer_outs = [x["tokenizer_out"] for x in batch]
        add_time_ids = torch.stack([x["add_time_ids"] for x in batch]).float()
        tokenizer_outputs = []
        for tokenizer_out in zip(*tokenizer_outs):
            input_ids = torch.concat([x["input_ids"] for x in tokenizer_out])
            attention_mask = torch.concat([x["attention_mask"] for x in tokenizer_out])
            tokenizer_outputs.append({"input_ids": input_ids, "attention_mask": attention_mask})
        return (samples, caption, tokenizer_outputs, {"time_ids": add_time_ids}, {})


class DummyDataset(UwUBaseDataset):
    def __init__(self, sample_size: Sequence[int] = (3, 1024, 1024), n_samples: int = 100, tokenizers: List = [], **kwargs):
        sample_size = tuple(sample_size)
        self.samples = [torch.randn(sample_size) for _ in range(n_samples)]
        self.tokenizers = tokenizers if isinstance(tokenizers, list) else [tokenizers]

    def set_tokenizers(self, tokenizers):
        self.tokenizers = tokenizers

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, index):
        caption = "DUMMY TEST"
        return {
            "sample": self.samples[index],
            "caption": caption,
            "tokenizer_out": [tok(caption, padding="max_length", truncation=True, return_tensors="pt") for tok in self.tokenizers],
            # org_h, org_w, crop_top, crop_left, target_h, target_w
            "add_time_ids": torch.tensor([1024, 1024, 0, 0, 1024, 1024]),
        }


class TrainDataModule:
    def __init__(self, dataset_config, dataloader_config):
        self.dataset_config = dataset_config
        self.dataloader_config = dataloader_config

    def setup(self, stage: str = "fit"):
        self.dataset = load_any(self.dataset_config)
        if hasattr(self, "tokenizers"):
            self.dataset.set_tokenizers(self.tokenizers)

    def train_dataloader(self):
        return Data.DataLoader(self.dataset, collate_fn=self.dataset.collate, **dict(self.dataloader_config))

    def set_tokenizers(self, tokenizers):
        self.tokenizers = tokenizers


class BaseTextEncoder(torch.nn.Module):
    """Marker base class (src/duwu/modules/text_encoders.py): `te(tokenizer_outputs)` -> (emb, normed_emb, pooled, mask)."""


class SyntheticTextEncoders(BaseTextEncoder):
    """Shape-faithful stand-in for `ConcatTextEncoders` (SDXL: CLIP-L 768 + CLIP-bigG 1280 hidden -> 2048, pooled 1280)."""

    def __init__(self, tokenizers=None, text_model_and_configs=None, zero_for_padding: bool = False, hidden_dim: int = 2048,
                 pooled_dim: int = 1280, seq_len: int = 77, **kwargs):
        super().__init__()
        self.tokenizers: list = []
        self.hidden_dim, self.pooled_dim, self.seq_len = hidden_dim, pooled_dim, seq_len
        self.register_buffer("_dev", torch.zeros(()), persistent=False)
        self._cache = {}

    @torch.no_grad()
    def forward(self, tokenizer_outputs, batch_size: int = None):
        B = batch_size
        if B is None:
            B = tokenizer_outputs[0]["input_ids"].shape[0] if tokenizer_outputs else 1
        key = (B, str(self._dev.device))
        if key not in self._cache:
            g = torch.Generator().manual_seed(77)  # caption is constant ("DUMMY TEST") -> constant embedding
            emb = torch.randn((1, self.seq_len, self.hidden_dim), generator=g).expand(B, -1, -1).contiguous()
            pooled = torch.randn((1, self.pooled_dim), generator=g).expand(B, -1).contiguous()
            normed = torch.nn.functional.layer_norm(emb, (self.hidden_dim,))
            self._cache[key] = tuple(t.to(self._dev.device) for t in (emb, normed, pooled))
        emb, normed, pooled = self._cache[key]
        return emb, normed, pooled, None


class SyntheticVAE(torch.nn.Module):
    """Shape-faithful stand-in for the frozen `AutoencoderKL` encoder: [B,3,H,W] pixels -> N(0,1)/scaling latents [B,4,H/8,W/8]."""

    def __init__(self, scaling_factor: float = 0.13025, latent_channels: int = 4):
        super().__init__()
        self.config = types.SimpleNamespace(scaling_factor=scaling_factor, latent_channels=latent_channels)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, **kwargs):
        return cls()

    @torch.no_grad()
    def encode(self, x):
        B, _, H, W = x.shape
        c = self.config.latent_channels
        # deterministic function of the pixels: 8x8 average pooling of the 3 channels (+ their mean), unit-variance scaled
        p = torch.nn.functional.avg_pool2d(x.float(), 8)
        lat = torch.cat([p, p.mean(1, keepdim=True)], dim=1)[:, :c] * (8.0 / self.config.scaling_factor)

        class _Dist:
            def sample(self_inner):
                return lat

        return types.SimpleNamespace(latent_dist=_Dist())
