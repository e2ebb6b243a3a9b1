"""-m gpu: model-file ingest end to end and packed-model cache hygiene.

An official-layout ``.pkl`` (chumpy arrays, scipy-sparse J_regressor, uint32 kintree root) is written
by tests/_official_pkl.py, loaded through ``SMPL(path)`` with neither chumpy nor the writer's fake
modules importable, packed onto the device and run through the kernels; results must match the CPU
oracle evaluated on the ORIGINAL eager-layout tensors.
"""
import copy
import sys

import numpy as np
import pytest
import torch

from human_3d_reconstruction_b200 import SMPL, synthetic
from oracle.smpl_ref import smpl_forward
from _official_pkl import write_official_pickle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _dev(dev, *arrs):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in arrs)


@pytest.mark.parametrize("sparse", ["fake-old-path", "scipy"])
@pytest.mark.parametrize("n,kw", [(5, dict(precision="fp32", lbs="fma")), (160, dict(precision="auto", lbs="auto")),
                                  (512, dict(precision="bf16x3", lbs="tc", joints="regressed"))])
def test_official_pickle_through_the_kernels(tmp_path, dev, full_model, sparse, n, kw):
    path = str(tmp_path / "basicModel_f_lbs_10_207_0_v1.0.0.pkl")
    write_official_pickle(path, full_model, sparse=sparse)
    assert "chumpy" not in sys.modules and "chumpy.ch" not in sys.modules
    layer = SMPL(path, **kw).to(dev)
    betas, pose, cam = synthetic.make_inputs(n, 91)
    with torch.no_grad():
        v, j, k = layer(*_dev(dev, betas, pose, cam))
    rv, rj, rk = smpl_forward(full_model, betas, pose, cam, dtype=torch.float32,
                              joints_from=kw.get("joints", "kinematic"))
    atol_v = 1e-6 if kw["precision"] == "fp32" else 1e-5
    assert torch.allclose(v.cpu(), rv, rtol=1e-5, atol=atol_v), (v.cpu() - rv).abs().max()
    atol_j = 1e-6 if kw.get("joints", "kinematic") == "kinematic" else 1e-5
    assert torch.allclose(j.cpu(), rj, rtol=1e-5, atol=atol_j), (j.cpu() - rj).abs().max()
    assert torch.allclose(k.cpu(), rk, rtol=1e-5, atol=2 * atol_j), (k.cpu() - rk).abs().max()


def test_load_state_dict_repacks_the_device_model(dev):
    """the packed model was cached forever; a load_state_dict after the first forward was ignored."""
    a, b = synthetic.make_model(0), synthetic.make_model(5)
    layer = SMPL(a, precision="fp32", lbs="fma").to(dev)
    betas, pose, cam = synthetic.make_inputs(6, 3)
    args = _dev(dev, betas, pose, cam)
    with torch.no_grad():
        v_a = layer(*args)[0].clone()
        layer.load_state_dict(SMPL(b).state_dict())
        v_b = layer(*args)[0]
    ref_b = smpl_forward(b, betas, pose, cam, dtype=torch.float32)[0]
    assert not torch.equal(v_a, v_b)
    assert torch.allclose(v_b.cpu(), ref_b, rtol=1e-5, atol=1e-6)
    # in-place edits are picked up after invalidate()
    with torch.no_grad():
        layer.v_template.add_(0.25)
        layer.invalidate()
        v_c = layer(*args)[0]
    assert torch.allclose(v_c.cpu(), ref_b + 0.25, rtol=1e-5, atol=2e-6)


def test_deepcopy_and_torch_save_of_a_used_module(dev, tmp_path):
    layer = SMPL.synthetic(0, precision="fp32", lbs="fma").to(dev)
    betas, pose, cam = synthetic.make_inputs(4, 3)
    args = _dev(dev, betas, pose, cam)
    with torch.no_grad():
        ref = layer(*args)
        ema = copy.deepcopy(layer)                      # raised "cannot pickle '_thread.lock'" in round 1
        torch.save(layer, tmp_path / "layer.pt")
        loaded = torch.load(tmp_path / "layer.pt", weights_only=False)
        for other in (ema, loaded):
            out = other(*args)
            assert all(torch.equal(x, y) for x, y in zip(out, ref))


def test_small_model_after_large_model_keeps_the_backward_working(dev):
    """the backward skinning kernel's shared-memory opt-in was set per model (from its V); a
    second, smaller model lowered the limit and the larger model's next backward failed to launch."""
    big = SMPL.synthetic(0).to(dev)
    small = SMPL(synthetic.make_model(3, num_verts=300)).to(dev)

    def loss_grad(layer, n):
        b, p, c = (torch.from_numpy(x).to(dev).requires_grad_() for x in synthetic.make_inputs(n, 11))
        v, j, k = layer(b, p, c)
        (v.square().mean() + j.square().mean() + k.abs().mean()).backward()
        torch.cuda.synchronize()
        return p.grad.clone()

    g0 = loss_grad(big, 8)
    loss_grad(small, 8)
    g1 = loss_grad(big, 8)
    assert torch.equal(g0, g1)
