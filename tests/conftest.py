import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

WINDOWS = [
    (0, 0, np.array([1.0])),
    (1, 1, np.array([-0.5, 0.0, 0.5])),
    (1, 1, np.array([1.0, -2.0, 1.0])),
]

TTS_HP = dict(stream_sizes=[180, 3, 1, 3], has_dynamic_features=[True, True, False, True],
              adversarial_streams=[True, False, False, False], mask_nth_mgc_for_adv_loss=2,
              num_windows=3, discriminator_linguistic_condition=False)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_ops():
    return np.load(os.path.join(GOLDEN, "ops.npz"))


@pytest.fixture(scope="session")
def golden_models():
    return np.load(os.path.join(GOLDEN, "models.npz"))


@pytest.fixture(scope="session")
def golden_step():
    return np.load(os.path.join(GOLDEN, "step.npz"))


@pytest.fixture(scope="session")
def golden_step_models():
    return np.load(os.path.join(GOLDEN, "step_models.npz"))


def rel_err(a, b):
    """max |a-b| / max(|b|) -- the 'relative to output scale' error used for float parity."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
