"""Time of the one-launch adapter fold of the SDXL + LyCORIS trainer (uwu_fold_batch over all 721 adapters): CUDA events over 20 calls."""
import os
import sys

os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from uwudiff_b200 import config as ucfg


def main():
    conf = bench.trainer_config(128, 16)
    trainer = ucfg.instantiate_any(conf["trainer"])
    ly = trainer.lycoris_model
    for _ in range(3):
        ly.fold_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ly.fold_all()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    n = sum(e[0].numel() for e in [(w,) for w in ly._fold["keep"][::4] if w is not None])
    print(f"fold_all: {ms:.3f} ms per call, {n / 1e9:.2f} G weight elements, {n * 6 / ms / 1e6:.0f} GB/s (4 B read + 2 B written per element)")


if __name__ == "__main__":
    main()
