import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device: on a box without one they are skipped, not failed.

    On a box WITH a GPU nothing is skipped: a missing libsmpl_b200.so must fail loudly there
    (capi.lib() raises), never turn into a silent skip or a CPU fallback.
    """
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def small_model():
    """Small-V SMPL-shaped model so the CPU suite stays fast (same layouts as the full model)."""
    from human_3d_reconstruction_b200 import synthetic
    return synthetic.make_model(3, num_verts=300)


@pytest.fixture(scope="session")
def full_model():
    from human_3d_reconstruction_b200 import synthetic
    return synthetic.make_model(0)
