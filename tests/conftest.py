import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests fail loudly (never skip silently) when selected without a GPU."""
    return


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def recipe_state_dict(module, seed: int):
    """Deterministic weights that depend only on the state-dict KEYS and shapes (not on construction order or
    torch's init code), so two module trees with equal keys -- the reference's and ours -- get equal parameters
    without a 40 M-parameter fixture.  Used by make_golden.py (reference side) and the tests (our side)."""
    import zlib
    import torch
    out = {}
    for k, v in module.state_dict().items():
        g = torch.Generator().manual_seed(seed + zlib.crc32(k.encode()))
        if k.endswith("num_batches_tracked"):
            out[k] = v.clone()
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif k.endswith("running_mean") or k.endswith("bias"):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1:
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)           # norm scales
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (1.5 / fan_in ** 0.5)
    return out
