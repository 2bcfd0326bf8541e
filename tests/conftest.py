import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def synthetic_sd():
    """The deterministic random-init weights every golden vector was produced with (seed 0)."""
    from lavie_b200.synthetic import synthetic_state_dict
    return synthetic_state_dict(seed=0)


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), map_location="cpu")


def rel_l2(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
