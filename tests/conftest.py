import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def _synthetic_conditioning():
    """No pretrained CLIP / VAE weights exist offline: trainer configs in the tests resolve `ConcatTextEncoders` /
    `AutoencoderKL` to the synthetic stand-ins (explicit opt-in, uwudiff_b200/config.py); the kernel-backed text towers and
    VAE encoder are tested directly in tests/test_conditioning_gpu.py."""
    from uwudiff_b200 import config

    prev = config.use_synthetic_conditioning(True)
    yield
    config.use_synthetic_conditioning(prev)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "loss_golden.npz"))


@pytest.fixture()
def fake_ops(monkeypatch):
    """Route uwudiff_b200.ops to the test-only torch emulation (tests/fake_ops.py) so host logic runs on CPU."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_ops as F

    F.install(monkeypatch)
    return F


LYCORIS_PRESET = dict(enable_conv=False, target_module=["Transformer2DModel"], target_name=[],
                      module_algo_map={"Attention": dict(algo="lokr", factor=64, full_matrix=True),
                                       "FeedForward": dict(algo="lokr", factor=6, full_matrix=True)})
LYCORIS_CFG = dict(linear_dim=4, linear_alpha=1, conv_dim=4, conv_alpha=1, algo="lora", use_tucker=True, train_norm=True)
