import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_model():
    """Small-V SMPL-shaped model so the CPU suite stays fast (same layouts as the full model)."""
    from human_3d_reconstruction_b200 import synthetic
    return synthetic.make_model(3, num_verts=300)


@pytest.fixture(scope="session")
def full_model():
    from human_3d_reconstruction_b200 import synthetic
    return synthetic.make_model(0)
