"""Load the REFERENCE's `src/duwu/loss/diffusion.py` verbatim (importlib, by path) with the diffusers shim.

Works only where /root/reference exists (the build container).  Used by oracle/make_golden.py to pin the
restatement in oracle/loss_oracle.py; nothing at GPU run time depends on it.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_FILE = "/root/reference/src/duwu/loss/diffusion.py"


def available() -> bool:
    return os.path.exists(REF_FILE)


def load_reference_loss_module():
    from . import diffusers_shim

    shim = types.ModuleType("diffusers")
    shim.EulerDiscreteScheduler = diffusers_shim.EulerDiscreteScheduler
    prev = sys.modules.get("diffusers")
    sys.modules["diffusers"] = shim
    try:
        spec = importlib.util.spec_from_file_location("_duwu_ref_loss_diffusion", REF_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if prev is None:
            sys.modules.pop("diffusers", None)
        else:
            sys.modules["diffusers"] = prev
    return mod


RF_FILE = "/root/reference/src/duwu/loss/rectified_flow.py"


def load_reference_rf_module():
    """Load the reference's `src/duwu/loss/rectified_flow.py` verbatim; its `from duwu.loss.diffusion import ...` is satisfied
    by registering the verbatim-loaded diffusion module under that name (stub parent packages, removed afterwards)."""
    diff = load_reference_loss_module()
    saved = {k: sys.modules.get(k) for k in ("duwu", "duwu.loss", "duwu.loss.diffusion")}
    pkg, sub = types.ModuleType("duwu"), types.ModuleType("duwu.loss")
    pkg.__path__, sub.__path__ = [], []
    sys.modules.update({"duwu": pkg, "duwu.loss": sub, "duwu.loss.diffusion": diff})
    try:
        spec = importlib.util.spec_from_file_location("_duwu_ref_loss_rf", RF_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
