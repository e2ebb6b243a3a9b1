"""The hand-derived backward (oracle/smpl_backward_np.py, the algorithm the CUDA kernels follow)
against torch autograd of the forward oracle, float64.  CPU only."""
import numpy as np
import pytest
import torch

from human_3d_reconstruction_b200 import synthetic
from oracle.smpl_backward_np import rodrigues_backward, smpl_backward_np
from oracle.smpl_ref import smpl_forward


def autograd_grads(model, betas, pose, cam, gV, gJ, gK, rotate_base=False, joints="kinematic"):
    tb, tp, tc = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (betas, pose, cam))
    v, j, k = smpl_forward(model, tb, tp, tc, dtype=torch.float64, rotate_base=rotate_base, joints_from=joints)
    loss = 0.0
    for out, g in ((v, gV), (j, gJ), (k, gK)):
        if g is not None:
            loss = loss + (out * torch.as_tensor(g, dtype=torch.float64)).sum()
    loss.backward()
    return tb.grad.numpy(), tp.grad.numpy(), tc.grad.numpy() if tc.grad is not None else np.zeros_like(cam)


@pytest.mark.parametrize("rotate_base", [False, True])
@pytest.mark.parametrize("which", ["all", "verts", "joints", "kp2d"])
def test_derivation_matches_autograd(rotate_base, which):
    m = synthetic.make_model(3, num_verts=200)
    b, p, c = synthetic.make_inputs(4, 7)
    rng = np.random.default_rng(1)
    gV = rng.normal(size=(4, 200, 3)) if which in ("all", "verts") else None
    gJ = rng.normal(size=(4, 24, 3)) if which in ("all", "joints") else None
    gK = rng.normal(size=(4, 24, 2)) if which in ("all", "kp2d") else None
    ref = autograd_grads(m, b, p, c, gV, gJ, gK, rotate_base)
    got = smpl_backward_np(m, b, p, c, gV, gJ, gK, rotate_base=rotate_base)
    for name, a, r in zip(("betas", "pose", "cam"), got, ref):
        scale = max(np.abs(r).max(), 1e-12)
        assert np.abs(a - r).max() <= 1e-7 * scale + 1e-12, name


def test_derivation_matches_autograd_regressed_joints():
    m = synthetic.make_model(5, num_verts=200)
    b, p, c = synthetic.make_inputs(3, 9)
    rng = np.random.default_rng(2)
    gV, gJ, gK = rng.normal(size=(3, 200, 3)), rng.normal(size=(3, 24, 3)), rng.normal(size=(3, 24, 2))
    for gv in (gV, None):
        ref = autograd_grads(m, b, p, c, gv, gJ, gK, True, joints="regressed")
        got = smpl_backward_np(m, b, p, c, gv, gJ, gK, rotate_base=True, joints_from="regressed")
        for name, a, r in zip(("betas", "pose", "cam"), got, ref):
            assert np.abs(a - r).max() <= 1e-7 * max(np.abs(r).max(), 1e-12) + 1e-12, name


def test_rodrigues_backward_small_and_zero_angles():
    rng = np.random.default_rng(0)
    g = rng.normal(size=(3, 3))
    for theta in (np.zeros(3), np.array([1e-4, -2e-4, 5e-5]), np.array([3.0, 0.1, -0.2])):
        got = rodrigues_backward(theta, g)
        assert np.all(np.isfinite(got))
        # central finite differences of the true rotation map (skipped at exactly 0: same limit)
        def R(t):
            a = np.linalg.norm(t)
            if a < 1e-300:
                return np.eye(3)
            k = t / a
            K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K
        h = 1e-6
        fd = np.array([((R(theta + h * e) - R(theta - h * e)) * g).sum() / (2 * h) for e in np.eye(3)])
        assert np.allclose(got, fd, rtol=1e-4, atol=1e-6)
