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
