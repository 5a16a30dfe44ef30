"""Generate tests/golden/sampling_golden.npz from the REFERENCE's `src/duwu/sampling/k_diffusion_wrapper.py` run verbatim
(it imports only torch; container only).  `cfg.py` / `k_diffusion_euler.py` import the absent k-diffusion package, so their
arithmetic is pinned indirectly: the golden trajectory below is produced by the reference wrapper class driven by a plain
restatement of the Euler-ancestral loop with the k-diffusion formulas.  TEST INFRASTRUCTURE.

    python -m oracle.make_sampling_golden
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

from . import diffusers_shim

REF = "/root/reference/src/duwu/sampling/k_diffusion_wrapper.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sampling_golden.npz")


def main():
    spec = importlib.util.spec_from_file_location("_ref_kdw", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    sch = diffusers_shim.EulerDiscreteScheduler.from_pretrained("x")
    acp = sch.alphas_cumprod
    den = ref.DiscreteEpsDDPMDenoiser(lambda x, t, **k: 0.3 * x + 0.01 * t.view(-1, 1, 1, 1), acp, False)
    out = {}
    sig = torch.tensor([14.6146, 9.0, 3.0, 1.0, 0.5, 0.1, 0.03, 0.0292])
    out["sigma_in"] = sig.numpy()
    out["sigma_to_t"] = den.sigma_to_t(sig).numpy()
    out["sigma_to_t_quant"] = den.sigma_to_t(sig, quantize=True).numpy()
    tt = torch.tensor([0.0, 0.25, 3.5, 500.0, 998.2, 999.0])
    out["t_in"] = tt.numpy()
    out["t_to_sigma"] = den.t_to_sigma(tt).numpy()
    out["get_sigmas_7"] = den.get_sigmas(7).numpy()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 4, 8, 8, generator=g)
    s4 = torch.tensor([14.0, 3.0, 0.7, 0.05])
    out["x"] = x.numpy()
    out["s4"] = s4.numpy()
    out["denoised"] = den(x, s4).numpy()
    out["denoised_cond"] = den(x, s4, sigma_cond=s4 * 0.5).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
