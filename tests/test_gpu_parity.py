"""GPU parity tests proper (-m gpu): the CUDA path through the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star):
  fp32 path (FMA blendshapes; FMA or 3xTF32-tcgen05 skinning)   rtol 1e-5, atol 1e-6 vs fp32 oracle
  tensor-core blendshape operands, stated looser bounds on vertices (metres, abs):
      f16x3  (split fp16, ~22 mantissa bits)  4e-6   (measured 1.9e-6; what 'auto' resolves to from 32 bodies)
      bf16x3 (split bf16, ~16 mantissa bits)  1e-5
      tf32                                      5e-4
      bf16                                      4e-3
  joints / kp2d never depend on the blendshape precision and always meet the fp32 tolerance.
PARITY UNPINNED: the oracle restates the published formulation (reference has no SMPL code).
"""
import os

import numpy as np
import pytest
import torch

from human_3d_reconstruction_b200 import SMPL, capi, sharding, synthetic
from human_3d_reconstruction_b200 import smpl as ops
from human_3d_reconstruction_b200.smpl import HostRunner
from oracle.smpl_ref import smpl_forward, smpl_forward_chunked

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6
VERT_ATOL = {"fp32": None, "f16x3": 4e-6, "bf16x3": 1e-5, "tf32": 5e-4, "bf16": 4e-3}
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smpl_golden_v1.npz")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def models():
    return {w: synthetic.make_model(0, weights=w) for w in ("sparse", "dense")}


def to_dev(dev, *arrs):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in arrs)


def assert_close(got, ref, rtol=RTOL, atol=ATOL, what=""):
    got, ref = got.detach().cpu(), torch.as_tensor(ref)
    err = (got.double() - ref.double()).abs().max().item() if got.numel() else 0.0
    assert torch.allclose(got, ref.to(got.dtype), rtol=rtol, atol=atol), f"{what}: max abs err {err:.3e}"


def check_verts(got, ref, precision, what=""):
    a = VERT_ATOL[precision]
    if a is None:
        assert_close(got, ref, what=what)
    else:
        assert_close(got, ref, rtol=0.0, atol=a, what=what)


# ------------------------------------------------------------------------------------------------
# per-kernel parity (smplb200_pose_chain / _blendshapes / _lbs / _regress_joints)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rotate_base", [False, True])
def test_k2_pose_chain(dev, models, rotate_base):
    n = 70
    betas, pose, cam = synthetic.make_inputs(n, 11)
    ref = smpl_forward(models["sparse"], betas, pose, rotate_base=rotate_base, return_intermediates=True)
    inter = ref[-1]
    layer = SMPL(models["sparse"], rotate_base=rotate_base).to(dev)
    tb, tp = to_dev(dev, betas, pose)
    coef, A, joints = ops.pose_chain(layer, tb, tp)
    NB = layer.num_betas
    assert torch.equal(coef[:, :NB].cpu(), torch.from_numpy(betas))
    assert_close(coef[:, NB:NB + 207], inter["pose_feature"], what="pose_feature")
    assert torch.all(coef[:, NB + 207:NB + 210] == 1.0) and torch.all(coef[:, NB + 210:] == 0.0)
    assert_close(A.view(n, 24, 3, 4), inter["A"][:, :, :3, :], what="A")
    assert_close(joints, ref[1], what="J_posed")


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "bf16x3", "tf32", "bf16"])
@pytest.mark.parametrize("n", [1, 33, 70])
def test_k1_blendshapes(dev, models, precision, n):
    betas, pose, _ = synthetic.make_inputs(n, 12)
    inter = smpl_forward(models["sparse"], betas, pose, return_intermediates=True)[-1]
    layer = SMPL(models["sparse"]).to(dev)
    tb, tp = to_dev(dev, betas, pose)
    coef, _, _ = ops.pose_chain(layer, tb, tp)
    vp = ops.blendshapes(layer, coef, flags=capi.make_flags(precision=precision))
    V = layer.num_verts
    assert vp.shape == (n, 3, layer.handle(dev).padded_verts)
    check_verts(vp[:, :, :V].permute(0, 2, 1), inter["v_posed"], precision, f"v_posed[{precision}]")
    assert torch.all(vp[:, :, V:] == 0), "padded planar columns must be zero"


@pytest.mark.parametrize("weights", ["sparse", "dense"])
@pytest.mark.parametrize("path", ["fma", "dense", "tc"])
@pytest.mark.parametrize("n", [1, 17, 70])
def test_k3_lbs(dev, models, weights, path, n):
    model = models[weights]
    betas, pose, cam = synthetic.make_inputs(n, 13)
    ref_v, ref_j, ref_k, inter = smpl_forward(model, betas, pose, cam, return_intermediates=True)
    layer = SMPL(model).to(dev)
    V, VP = layer.num_verts, layer.handle(dev).padded_verts
    vp = torch.zeros((n, 3, VP), device=dev)
    vp[:, :, :V] = inter["v_posed"].permute(0, 2, 1).to(dev)
    A = inter["A"][:, :, :3, :].reshape(n, 24, 12).contiguous().to(dev)
    (tc,) = to_dev(dev, cam)
    verts, kp = ops.lbs(layer, vp, A, joints=ref_j.to(dev).contiguous(), cam=tc,
                        flags=capi.make_flags(lbs=path))
    assert_close(verts, ref_v, what=f"verts[{weights},{path}]")
    assert_close(kp, ref_k, what="kp2d")


def test_regressed_joints_kernel(dev, models):
    n = 9
    betas, pose, cam = synthetic.make_inputs(n, 14)
    for reg in ("sparse", "dense"):
        model = synthetic.make_model(0, regressor=reg)
        ref_v, ref_j, ref_k = smpl_forward(model, betas, pose, cam, joints_from="regressed")
        layer = SMPL(model).to(dev)
        (tc,) = to_dev(dev, cam)
        j, k = ops.regress_joints(layer, ref_v.to(dev).contiguous(), tc)
        assert_close(j, ref_j, what=f"regressed joints [{reg}]")
        assert_close(k, ref_k, atol=2e-6, what="kp2d")


# ------------------------------------------------------------------------------------------------
# the forward pass through smplb200_forward (nn.Module)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 31, 64, 257])
@pytest.mark.parametrize("precision,lbs", [("fp32", "fma"), ("fp32", "tc"), ("bf16x3", "tc"), ("f16x3", "tc"),
                                           ("tf32", "fma"), ("bf16", "tc"), ("auto", "auto")])
def test_forward_vs_oracle(dev, models, n, precision, lbs):
    model = models["sparse"]
    betas, pose, cam = synthetic.make_inputs(n, 20 + n)
    ref_v, ref_j, ref_k = smpl_forward(model, betas, pose, cam)
    layer = SMPL(model, precision=precision, lbs=lbs).to(dev)
    v, j, k = layer(*to_dev(dev, betas, pose, cam))
    eff = precision if precision != "auto" else ("f16x3" if n >= capi.TC_MIN_BATCH else "fp32")
    check_verts(v, ref_v, eff, f"vertices n={n} {precision}/{lbs}")
    assert_close(j, ref_j, what="joints")
    assert_close(k, ref_k, atol=2e-6, what="kp2d")
    v2, j2 = layer(*to_dev(dev, betas, pose))  # no camera -> 2 outputs, same numbers
    assert torch.equal(v2, v) and torch.equal(j2, j)


@pytest.mark.parametrize("weights", ["sparse", "dense"])
@pytest.mark.parametrize("rotate_base", [False, True])
@pytest.mark.parametrize("joints", ["kinematic", "regressed"])
def test_golden_vectors(dev, weights, rotate_base, joints):
    g = np.load(GOLDEN)
    model = synthetic.make_model(int(g["model_seed"]), weights=weights)
    tag = f"{weights}_rb{int(rotate_base)}_{joints}"
    idx = torch.from_numpy(g["vert_idx"])
    for lbs in ("fma", "tc"):
        layer = SMPL(model, precision="fp32", lbs=lbs, rotate_base=rotate_base, joints=joints).to(dev)
        v, j, k = layer(*to_dev(dev, g["betas"], g["pose"], g["cam"]))
        # golden vectors are float64; the fp32 kernels must be within fp32 rounding of them
        assert_close(v.cpu()[:, idx], g[f"verts_{tag}"].astype(np.float32), atol=2e-6, what="golden verts")
        assert_close(j, g[f"joints_{tag}"].astype(np.float32), atol=2e-6, what="golden joints")
        assert_close(k, g[f"kp2d_{tag}"].astype(np.float32), atol=3e-6, what="golden kp2d")


def test_error_vs_fp64_not_worse_than_cpu_fp32(dev, models):
    """arbiter check -- err(kernel, fp64) <= c * err(fp32 CPU oracle, fp64)."""
    model = models["sparse"]
    betas, pose, cam = synthetic.make_inputs(128, 31)
    r64 = smpl_forward(model, betas, pose, cam, dtype=torch.float64)
    r32 = smpl_forward(model, betas, pose, cam, dtype=torch.float32)
    layer = SMPL(model, precision="fp32", lbs="fma").to(dev)
    out = layer(*to_dev(dev, betas, pose, cam))
    for got, a32, a64 in zip(out, r32, r64):
        e_gpu = (got.cpu().double() - a64).abs().max().item()
        e_cpu = (a32.double() - a64).abs().max().item()
        assert e_gpu <= 3.0 * e_cpu + 1e-7, (e_gpu, e_cpu)


def test_empty_batch(dev, models):
    layer = SMPL(models["sparse"]).to(dev)
    z = lambda w: torch.zeros((0, w), device=dev)
    v, j, k = layer(z(10), z(72), z(3))
    assert v.shape == (0, 6890, 3) and j.shape == (0, 24, 3) and k.shape == (0, 24, 2)


def test_shard_equivalence_bitwise(dev, models):
    """A.9(viii): no cross-body term => forward(N) == concat(forward(shards)) bit for bit per path."""
    betas, pose, cam = synthetic.make_inputs(300, 33)
    tb, tp, tc = to_dev(dev, betas, pose, cam)
    for precision, lbs in (("fp32", "fma"), ("bf16x3", "tc"), ("f16x3", "tc"), ("tf32", "tc")):
        layer = SMPL(models["sparse"], precision=precision, lbs=lbs).to(dev)
        full = layer(tb, tp, tc)
        parts = [layer(tb[a:b], tp[a:b], tc[a:b]) for a, b in ((0, 7), (7, 150), (150, 300))]
        for i in range(3):
            assert torch.equal(full[i], torch.cat([p[i] for p in parts])), (precision, lbs, i)


def test_sparse_and_dense_lbs_paths_agree(dev, models):
    """A.9(vi): the ELL path skips exact zeros, so it equals the dense path on a sparse model."""
    betas, pose, cam = synthetic.make_inputs(40, 34)
    args = to_dev(dev, betas, pose, cam)
    a = SMPL(models["sparse"], precision="fp32", lbs="fma").to(dev)(*args)
    b = SMPL(models["sparse"], precision="fp32", lbs="dense").to(dev)(*args)
    assert (a[0] - b[0]).abs().max().item() <= 1e-7  # same terms, different association only


def test_host_entry_matches_device_entry(dev, models):
    n = 100
    layer = SMPL(models["sparse"], precision="fp32", lbs="fma").to(dev)
    betas, pose, cam = synthetic.make_inputs(n, 35)
    v, j, k = layer(*to_dev(dev, betas, pose, cam))
    runner = HostRunner(layer, n, dev, with_vertices=True, with_cam=True)
    runner.betas.copy_(torch.from_numpy(betas)); runner.pose.copy_(torch.from_numpy(pose))
    runner.cam.copy_(torch.from_numpy(cam))
    runner.run()
    torch.cuda.synchronize()
    assert torch.equal(runner.vertices, v.cpu()) and torch.equal(runner.joints, j.cpu())
    assert torch.equal(runner.kp2d, k.cpu())
    assert runner.h2d_bytes == n * (10 + 72 + 3) * 4 and runner.d2h_bytes == n * (6890 * 3 + 72 + 48) * 4


def test_non_default_stream_and_reentrancy(dev, models):
    layer = SMPL(models["sparse"], precision="fp32", lbs="fma").to(dev)
    betas, pose, cam = synthetic.make_inputs(50, 36)
    args = to_dev(dev, betas, pose, cam)
    ref = layer(*args)
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        out = [layer(*args) for _ in range(3)]
    s.synchronize()
    for o in out:
        assert all(torch.equal(x, y) for x, y in zip(o, ref))


@pytest.mark.parametrize("num_verts,num_betas", [(300, 10), (1000, 8), (129, 14)])
def test_other_model_shapes(dev, num_verts, num_betas):
    """Odd tile counts (dummy second tile of a CTA pair / tile pair), V not a multiple of 128, NB != 10."""
    model = synthetic.make_model(5, num_verts=num_verts, num_betas=num_betas)
    n = 200
    betas, pose, cam = synthetic.make_inputs(n, 61, num_betas=num_betas)
    ref_v, ref_j, ref_k = smpl_forward(model, betas, pose, cam)
    for precision, lbs in (("fp32", "fma"), ("fp32", "tc"), ("bf16x3", "tc"), ("f16x3", "tc"), ("tf32", "tc")):
        layer = SMPL(model, precision=precision, lbs=lbs).to(dev)
        v, j, k = layer(*to_dev(dev, betas, pose, cam))
        check_verts(v, ref_v, precision, f"V={num_verts} NB={num_betas} {precision}/{lbs}")
        assert_close(j, ref_j, what="joints")
        assert_close(k, ref_k, atol=2e-6, what="kp2d")


def test_cuda_graph_replay_matches_eager(dev, models):
    from human_3d_reconstruction_b200 import GraphedSMPL
    for n, kw in ((64, dict(precision="fp32", lbs="fma")), (300, dict(precision="bf16x3", lbs="tc"))):
        layer = SMPL(models["sparse"], **kw).to(dev)
        g = GraphedSMPL(layer, n, dev, with_cam=True)
        for seed in (51, 52):  # replay twice with different parameters
            betas, pose, cam = synthetic.make_inputs(n, seed)
            tb, tp, tc = to_dev(dev, betas, pose, cam)
            g.betas.copy_(tb); g.pose.copy_(tp); g.cam.copy_(tc)
            v, j, k = g.replay()
            torch.cuda.synchronize()
            ref = layer(tb, tp, tc)
            assert torch.equal(v, ref[0]) and torch.equal(j, ref[1]) and torch.equal(k, ref[2])


def test_return_kp2d_flag(dev, models):
    layer = SMPL(models["sparse"]).to(dev)
    b, p, c = to_dev(dev, *synthetic.make_inputs(3, 2))
    with torch.no_grad():
        assert len(layer(b, p, c)) == 3 and len(layer(b, p, c, return_kp2d=True)) == 3
        assert len(layer(b, p, c, return_kp2d=False)) == 2 and len(layer(b, p)) == 2
        with pytest.raises(ValueError):
            layer(b, p, return_kp2d=True)


def test_errors(dev, models):
    layer = SMPL(models["sparse"]).to(dev)
    b, p = torch.zeros(2, 10, device=dev), torch.zeros(2, 72, device=dev)
    with pytest.raises(TypeError):
        layer(b.double(), p)
    with pytest.raises(ValueError):
        layer(b, torch.zeros(2, 71, device=dev))
    with torch.no_grad():
        layer(b.clone().requires_grad_(), p)  # no graph under no_grad
    h = layer.handle(dev)
    ws = torch.empty(16, dtype=torch.uint8, device=dev)
    st = capi.lib().smplb200_forward(h.ptr, b.data_ptr(), p.data_ptr(), None, 2,
                                     torch.empty(2, 6890, 3, device=dev).data_ptr(), None, None,
                                     ws.data_ptr(), 16, 0, None)
    assert st == 3  # SMPLB200_ERR_WORKSPACE
    assert capi.lib().smplb200_workspace_bytes(h.ptr, 2, 0xFFFF0000) == 0  # unknown flag bits


def test_data_parallel_wrapper(dev, models):
    """The reference's only multi-GPU mechanism is nn.DataParallel (trainer.py:176); the layer must
    survive replicate() -- buffers are per-device, the packed handle is created per device."""
    layer = SMPL(models["sparse"], precision="fp32", lbs="fma").to(dev)
    dp = torch.nn.DataParallel(layer, device_ids=[0])
    betas, pose, cam = synthetic.make_inputs(12, 37)
    args = to_dev(dev, betas, pose, cam)
    with torch.no_grad():
        out = dp(*args)
        ref = layer(*args)
    assert all(torch.equal(x, y) for x, y in zip(out, ref))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two visible GPUs")
def test_data_parallel_two_devices_forward_and_backward(models):
    """nn.DataParallel across two GPUs, the way the reference trainer wraps its model
    (trainer.py:176): per-device handles created from replica threads, scatter/gather autograd around
    the layer's own autograd node.  Pinned FMA paths => bitwise equal to the single-device result."""
    d0 = torch.device("cuda:0")
    layer = SMPL(models["sparse"], precision="fp32", lbs="fma").to(d0)
    dp = torch.nn.DataParallel(layer, device_ids=[0, 1])
    n = 40
    betas, pose, cam = synthetic.make_inputs(n, 43)
    g = torch.Generator().manual_seed(7)
    uv, uj, uk = torch.randn(n, 6890, 3, generator=g), torch.randn(n, 24, 3, generator=g), torch.randn(n, 24, 2, generator=g)

    def run(module):
        args = [torch.from_numpy(x).to(d0).requires_grad_() for x in (betas, pose, cam)]
        v, j, k = module(*args)
        ((v * uv.to(d0)).sum() + (j * uj.to(d0)).sum() + (k * uk.to(d0)).sum()).backward()
        torch.cuda.synchronize()
        return (v.detach(), j.detach(), k.detach()) + tuple(a.grad for a in args)

    got, ref = run(dp), run(layer)
    assert set(layer._handles) == {0, 1}, "one packed model handle per device"
    for name, a, b in zip(("verts", "joints", "kp2d", "g_betas", "g_pose", "g_cam"), got, ref):
        assert torch.equal(a, b), name


# ------------------------------------------------------------------------------------------------
# BASELINE.json full size (N = 4096): oracle parity (chunked) + size-independent properties
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def big(dev, models):
    n = 4096
    betas, pose, cam = synthetic.make_inputs(n, 41)
    return n, betas, pose, cam, to_dev(dev, betas, pose, cam)


def test_full_size_vs_oracle(dev, models, big):
    n, betas, pose, cam, args = big
    ref_v, ref_j, ref_k = smpl_forward_chunked(models["sparse"], betas, pose, cam, chunk=512)
    for precision, lbs in (("bf16x3", "tc"), ("fp32", "tc")):
        layer = SMPL(models["sparse"], precision=precision, lbs=lbs).to(dev)
        v, j, k = layer(*args)
        check_verts(v, ref_v, precision, f"N=4096 vertices {precision}")
        assert_close(j, ref_j, what="joints")
        assert_close(k, ref_k, atol=2e-6, what="kp2d")


def test_largest_config_is_shard_invariant(dev, models):
    """BASELINE.json configs[3]: 65,536 bodies in one call (11 GB of outputs + scratch on one GPU) equal,
    bit for bit, the eight 8,192-body shards an 8-GPU job would compute (SURVEY A.9 viii), and every
    sampled body matches the oracle."""
    n, world = 65536, 8
    betas, pose, cam = synthetic.make_inputs(n, 65)
    layer = SMPL(models["sparse"], precision="bf16x3", lbs="tc").to(dev)
    tb, tp, tc = to_dev(dev, betas, pose, cam)
    with torch.no_grad():
        v, j, k = layer(tb, tp, tc)
        for r in (0, 3, 7):
            lo, hi = sharding.shard_bounds(n, world, r)
            sv, sj, sk = layer(tb[lo:hi], tp[lo:hi], tc[lo:hi])
            assert torch.equal(sv, v[lo:hi]) and torch.equal(sj, j[lo:hi]) and torch.equal(sk, k[lo:hi])
            del sv, sj, sk
    pick = np.array([0, 1, 8191, 8192, 40000, 65535])
    rv, rj, rk = smpl_forward(models["sparse"], betas[pick], pose[pick], cam[pick])
    check_verts(v[pick], rv, "bf16x3", "N=65536 sampled vertices")
    assert_close(j[pick], rj, what="joints")
    assert_close(k[pick], rk, atol=2e-6, what="kp2d")


def test_full_size_properties(dev, models, big):
    n, betas, pose, cam, (tb, tp, tc) = big
    layer = SMPL(models["sparse"], precision="auto", lbs="auto").to(dev)
    # (i) zero pose => vertices == v_template + betas . shapedirs   (A.9 i)
    v0, j0 = layer(tb, torch.zeros_like(tp))
    vs = (tb @ layer.shapedirs).view(n, -1, 3) + layer.v_template
    assert (v0 - vs).abs().max().item() < 1e-5
    # (iii) root-only rotation about z is a rigid motion about the root joint   (A.9 iii)
    ang = 0.9
    pz = torch.zeros_like(tp); pz[:, 2] = ang
    v1, j1 = layer(tb, pz)
    c, s = np.cos(ang), np.sin(ang)
    Rz = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0, 0, 1.0]], device=dev, dtype=torch.float32)
    J0 = j0[:, :1]
    assert (v1 - ((v0 - J0) @ Rz.T + J0)).abs().max().item() < 1e-5
    # (vii) identity camera => kp2d == joints_xy
    ident = torch.tensor([[1.0, 0.0, 0.0]], device=dev).expand(n, 3).contiguous()
    v2, j2, k2 = layer(tb, tp, ident)
    assert torch.equal(k2, j2[:, :, :2])
    # (viii) shard equivalence at full size, bitwise
    halves = [layer(tb[a:b], tp[a:b], ident[a:b]) for a, b in ((0, 1000), (1000, 4096))]
    assert torch.equal(v2, torch.cat([h[0] for h in halves]))
    # idempotence / determinism
    v3, _, _ = layer(tb, tp, ident)
    assert torch.equal(v2, v3)
