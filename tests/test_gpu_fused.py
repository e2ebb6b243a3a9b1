"""-m gpu: the fused blendshapes + skinning kernel (precision 'f16', csrc/k_fused_tc.cuh) vs the CPU oracle.

Stated bound for the vertices: 5e-5 m absolute (fp16 operands on the 207 pose rows; measured 1.9e-5 against
float64 -- an order of magnitude inside the TF32 error BASELINE.json configs[2] allows); shape rows, template
and the skinning blend are split exactly, so with a zero pose the result meets the fp32 tolerance.  Joints and
kp2d come from k2 and always meet the fp32 tolerance.  PARITY UNPINNED like every SMPL test (no reference code).
"""
import numpy as np
import pytest
import torch

from human_3d_reconstruction_b200 import SMPL, GraphedSMPL, capi, synthetic
from human_3d_reconstruction_b200 import smpl as ops
from human_3d_reconstruction_b200.smpl import HostRunner
from oracle.smpl_ref import smpl_forward, smpl_forward_chunked

pytestmark = pytest.mark.gpu
F16_ATOL = 5e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def models():
    return {w: synthetic.make_model(0, weights=w) for w in ("sparse", "dense")}


def to_dev(dev, *arrs):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in arrs)


def close(got, ref, rtol, atol, what):
    got, ref = got.detach().cpu(), torch.as_tensor(ref)
    err = (got.double() - ref.double()).abs().max().item() if got.numel() else 0.0
    assert torch.allclose(got, ref.to(got.dtype), rtol=rtol, atol=atol), f"{what}: max abs err {err:.3e}"
    return err


@pytest.mark.parametrize("weights", ["sparse", "dense"])
@pytest.mark.parametrize("n", [1, 3, 4, 63, 64, 65, 200])
def test_blend_skin_entry_point(dev, models, weights, n):
    """smplb200_blend_skin on k2's own coef / A: sub-block (4) and block (64) boundaries, ragged tails."""
    model = models[weights]
    betas, pose, cam = synthetic.make_inputs(n, 70 + n)
    ref_v = smpl_forward(model, betas, pose)[0]
    layer = SMPL(model).to(dev)
    coef, A, _ = ops.pose_chain(layer, *to_dev(dev, betas, pose))
    v = ops.blend_skin(layer, coef, A)
    close(v, ref_v, 0.0, F16_ATOL, f"vertices [{weights}] n={n}")


@pytest.mark.parametrize("n", [1, 2, 31, 64, 257, 1000])
@pytest.mark.parametrize("joints", ["kinematic", "regressed"])
@pytest.mark.parametrize("rotate_base", [False, True])
def test_forward_f16_vs_oracle(dev, models, n, joints, rotate_base):
    model = models["sparse"]
    betas, pose, cam = synthetic.make_inputs(n, 120 + n)
    ref_v, ref_j, ref_k = smpl_forward_chunked(model, betas, pose, cam, chunk=256, joints_from=joints,
                                               rotate_base=rotate_base)
    layer = SMPL(model, precision="f16", joints=joints, rotate_base=rotate_base).to(dev)
    v, j, k = layer(*to_dev(dev, betas, pose, cam))
    err = close(v, ref_v, 0.0, F16_ATOL, f"vertices n={n}")
    jt = (1e-5, 1e-6) if joints == "kinematic" else (0.0, F16_ATOL)      # regressed joints inherit the vertex error
    close(j, ref_j, *jt, "joints")
    close(k, ref_k, jt[0], 2 * jt[1] if joints == "kinematic" else 2 * F16_ATOL, "kp2d")
    v2, j2 = layer(*to_dev(dev, betas, pose))
    assert torch.equal(v2, v) and torch.equal(j2, j)
    print(f"n={n} {joints}: max vertex err {err:.2e}")


def test_zero_pose_meets_fp32_tolerance(dev, models):
    """Only the pose rows are single-MMA fp16: with pose = 0 (pose_feature = 0 up to the 1e-8 guard) the
    shape blend, template and skinning are all exact-split -> fp32-class result."""
    n = 130
    betas, pose, cam = synthetic.make_inputs(n, 5)
    pose[:] = 0.0
    ref_v = smpl_forward(models["dense"], betas, pose)[0]
    layer = SMPL(models["dense"], precision="f16").to(dev)
    v, _ = layer(*to_dev(dev, betas, pose))
    close(v, ref_v, 1e-5, 2e-6, "vertices at zero pose")


def test_f16_shard_equivalence_bitwise(dev, models):
    """Bodies are columns of the MMAs: a body's vertices do not depend on where in a batch it sits."""
    betas, pose, cam = synthetic.make_inputs(333, 33)
    tb, tp, tc = to_dev(dev, betas, pose, cam)
    layer = SMPL(models["sparse"], precision="f16").to(dev)
    full = layer(tb, tp, tc)
    parts = [layer(tb[a:b], tp[a:b], tc[a:b]) for a, b in ((0, 7), (7, 150), (150, 214), (214, 333))]
    for i in range(3):
        assert torch.equal(full[i], torch.cat([p[i] for p in parts])), i
    again = layer(tb, tp, tc)
    assert all(torch.equal(a, b) for a, b in zip(full, again)), "run-to-run reproducible"


@pytest.mark.parametrize("num_verts,num_betas", [(300, 10), (1000, 8), (129, 13), (128, 1)])
def test_f16_other_model_shapes(dev, num_verts, num_betas):
    model = synthetic.make_model(5, num_verts=num_verts, num_betas=num_betas)
    n = 150
    betas, pose, cam = synthetic.make_inputs(n, 61, num_betas=num_betas)
    ref_v, ref_j, ref_k = smpl_forward(model, betas, pose, cam)
    layer = SMPL(model, precision="f16").to(dev)
    v, j, k = layer(*to_dev(dev, betas, pose, cam))
    close(v, ref_v, 0.0, F16_ATOL, f"V={num_verts} NB={num_betas}")
    close(j, ref_j, 1e-5, 1e-6, "joints")
    close(k, ref_k, 1e-5, 2e-6, "kp2d")


def test_f16_rejected_when_the_model_has_too_many_betas(dev):
    model = synthetic.make_model(5, num_verts=200, num_betas=14)          # 14 + 3 template rows > one K step
    layer = SMPL(model, precision="f16").to(dev)
    betas, pose, cam = synthetic.make_inputs(4, 1, num_betas=14)
    with pytest.raises(RuntimeError):
        layer(*to_dev(dev, betas, pose, cam))
    assert capi.lib().smplb200_blend_skin_workspace_bytes(layer.handle(dev).ptr, 4) == 0


def test_f16_full_batch_4096(dev, models):
    """BASELINE.json configs[2] at its full size: every vertex of 4096 bodies against the CPU oracle."""
    n = 4096
    betas, pose, cam = synthetic.make_inputs(n, 1)
    ref_v, ref_j, ref_k = smpl_forward_chunked(models["sparse"], betas, pose, cam, chunk=512)
    layer = SMPL(models["sparse"], precision="f16").to(dev)
    v, j, k = layer(*to_dev(dev, betas, pose, cam))
    err = close(v, ref_v, 0.0, F16_ATOL, "vertices")
    close(j, ref_j, 1e-5, 1e-6, "joints")
    close(k, ref_k, 1e-5, 2e-6, "kp2d")
    print(f"4096 bodies: max vertex err {err:.2e} m")


def test_f16_host_entry_graph_and_backward(dev, models):
    layer = SMPL(models["sparse"], precision="f16").to(dev)
    n = 96
    betas, pose, cam = synthetic.make_inputs(n, 17)
    args = to_dev(dev, betas, pose, cam)
    with torch.no_grad():
        v, j, k = layer(*args)
    # host entry point
    r = HostRunner(layer, n, dev, with_vertices=True)
    r.betas.copy_(torch.from_numpy(betas)); r.pose.copy_(torch.from_numpy(pose)); r.cam.copy_(torch.from_numpy(cam))
    r.run()
    torch.cuda.synchronize()
    assert torch.equal(r.vertices, v.cpu()) and torch.equal(r.joints, j.cpu()) and torch.equal(r.kp2d, k.cpu())
    # CUDA graph replay
    g = GraphedSMPL(layer, n, dev)
    g.betas.copy_(args[0]); g.pose.copy_(args[1]); g.cam.copy_(args[2])
    gv, gj, gk = g.replay()
    torch.cuda.synchronize()
    assert torch.equal(gv, v) and torch.equal(gj, j) and torch.equal(gk, k)
    # backward through the autograd node: the fused forward keeps no vposed, the backward recomputes it
    gb, gp, gc = (a.clone().requires_grad_() for a in args)
    ov, oj, ok = layer(gb, gp, gc)
    (ov.square().mean() + oj.square().mean() + ok.abs().mean()).backward()
    rb, rp, rc = (torch.from_numpy(x).double().requires_grad_() for x in (betas, pose, cam))
    rv, rj, rk = smpl_forward(models["sparse"], rb, rp, rc, dtype=torch.float64)
    (rv.square().mean() + rj.square().mean() + rk.abs().mean()).backward()
    for name, got, ref in (("g_betas", gb.grad, rb.grad), ("g_pose", gp.grad, rp.grad), ("g_cam", gc.grad, rc.grad)):
        err, scale = (got.cpu().double() - ref).abs().max().item(), ref.abs().max().item()
        assert err <= 1e-4 * scale + 1e-7, f"{name}: {err:.3e} vs scale {scale:.3e}"
