"""GPU parity of the backward pass (-m gpu): smplb200_backward through the autograd node vs torch
autograd of the CPU oracle in float64.

Tolerance: gradients are fp32 sums over up to 6890 vertices; every component must agree with the
float64 autograd value to 1e-4 of the largest component of that gradient tensor (GRAD_RTOL), plus
1e-6 absolute.  With tensor-core blendshape operands in the recompute (bf16x3, batch >= 256) the
same bound holds.  PARITY UNPINNED, like the forward (the reference has no SMPL code).
"""
import numpy as np
import pytest
import torch

from human_3d_reconstruction_b200 import SMPL, capi, synthetic
from oracle.smpl_ref import smpl_forward

pytestmark = pytest.mark.gpu
GRAD_RTOL, GRAD_ATOL = 1e-4, 1e-6


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def models():
    return {w: synthetic.make_model(0, weights=w) for w in ("sparse", "dense")}


def oracle_grads(model, betas, pose, cam, ups, rotate_base=False, joints="kinematic"):
    """ups: list of upstream gradients for (vertices, joints, kp2d); None = output unused."""
    tb, tp = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (betas, pose))
    tc = None if cam is None else torch.tensor(cam, dtype=torch.float64, requires_grad=True)
    outs = smpl_forward(model, tb, tp, tc, dtype=torch.float64, rotate_base=rotate_base, joints_from=joints)
    loss = sum((o * torch.as_tensor(g, dtype=torch.float64)).sum() for o, g in zip(outs, ups) if g is not None)
    loss.backward()
    return tb.grad, tp.grad, None if tc is None else tc.grad


def gpu_grads(layer, dev, betas, pose, cam, ups):
    tb, tp = (torch.from_numpy(x).to(dev).requires_grad_() for x in (betas, pose))
    tc = None if cam is None else torch.from_numpy(cam).to(dev).requires_grad_()
    outs = layer(tb, tp, tc)
    loss = sum((o * torch.from_numpy(g.astype(np.float32)).to(dev)).sum() for o, g in zip(outs, ups) if g is not None)
    loss.backward()
    torch.cuda.synchronize()
    return tb.grad, tp.grad, None if tc is None else tc.grad


def assert_grads(got, ref, what):
    for name, g, r in zip(("g_betas", "g_pose", "g_cam"), got, ref):
        if r is None:
            assert g is None, f"{what}: {name} should be None"
            continue
        g, r = g.detach().cpu().double(), r.double()
        assert torch.isfinite(g).all(), f"{what}: {name} not finite"
        scale = max(r.abs().max().item(), 1e-12)
        err = (g - r).abs().max().item()
        assert err <= GRAD_RTOL * scale + GRAD_ATOL, f"{what}: {name} max err {err:.3e} vs scale {scale:.3e}"


def upstream(n, V, seed, which=("v", "j", "k")):
    rng = np.random.default_rng(seed)
    gv, gj, gk = rng.normal(size=(n, V, 3)), rng.normal(size=(n, 24, 3)), rng.normal(size=(n, 24, 2))
    return [gv if "v" in which else None, gj if "j" in which else None, gk if "k" in which else None]


@pytest.mark.parametrize("weights", ["sparse", "dense"])
@pytest.mark.parametrize("rotate_base", [False, True])
@pytest.mark.parametrize("joints", ["kinematic", "regressed"])
def test_backward_all_outputs(dev, models, weights, rotate_base, joints):
    m = models[weights]
    n = 5
    b, p, c = synthetic.make_inputs(n, 11)
    ups = upstream(n, 6890, 3)
    layer = SMPL(m, precision="fp32", lbs="fma", joints=joints, rotate_base=rotate_base).to(dev)
    got = gpu_grads(layer, dev, b, p, c, ups)
    ref = oracle_grads(m, b, p, c, ups, rotate_base, joints)
    assert_grads(got, ref, f"{weights}/{joints}/rb={rotate_base}")


@pytest.mark.parametrize("which", [("v",), ("j",), ("k",), ("j", "k"), ("v", "k")])
@pytest.mark.parametrize("joints", ["kinematic", "regressed"])
def test_backward_output_subsets(dev, models, which, joints):
    """Outputs that do not reach the loss arrive as None (no materialised zero gradients)."""
    m = models["sparse"]
    n = 3
    b, p, c = synthetic.make_inputs(n, 5)
    ups = upstream(n, 6890, 4, which)
    layer = SMPL(m, precision="fp32", joints=joints).to(dev)
    got = gpu_grads(layer, dev, b, p, c, ups)
    ref = oracle_grads(m, b, p, c, ups, False, joints)
    if "k" not in which:      # cam does not reach the loss: autograd leaves .grad unset on both sides
        assert got[2] is None or got[2].abs().max().item() == 0.0
        got, ref = (got[0], got[1], None), (ref[0], ref[1], None)
    assert_grads(got, ref, f"{which}/{joints}")


def test_backward_without_cam(dev, models):
    m = models["sparse"]
    b, p, _ = synthetic.make_inputs(4, 8)
    ups = upstream(4, 6890, 6, ("v", "j"))[:2]
    layer = SMPL(m, precision="fp32").to(dev)
    got = gpu_grads(layer, dev, b, p, None, ups)
    ref = oracle_grads(m, b, p, None, ups)
    assert_grads(got, ref, "no cam")


@pytest.mark.parametrize("n", [1, 17, 300])
def test_backward_batch_sizes_and_auto_precision(dev, n):
    """n = 300 takes the tensor-core recompute (bf16x3 blendshapes) and several column slices."""
    m = synthetic.make_model(2, num_verts=1000)
    b, p, c = synthetic.make_inputs(n, 21)
    ups = upstream(n, 1000, 7)
    layer = SMPL(m).to(dev)
    got = gpu_grads(layer, dev, b, p, c, ups)
    ref = oracle_grads(m, b, p, c, ups)
    assert_grads(got, ref, f"n={n}")


@pytest.mark.parametrize("precision,rtol", [("bf16x3", GRAD_RTOL), ("tf32", 3e-3), ("bf16", 1e-2)])
@pytest.mark.parametrize("n", [5, 130])
def test_backward_tensor_core_blendshape_gradient(dev, models, precision, rtol, n):
    """tcgen05 blendshape backward: 3xTF32 for bf16x3 (fp32-class bound), 1xTF32 for tf32 / bf16
    (stated looser bounds: the recomputed vposed and the g_coef operands are reduced precision)."""
    m = models["sparse"] if n == 5 else synthetic.make_model(6, num_verts=2000)
    V = 6890 if n == 5 else 2000
    b, p, c = synthetic.make_inputs(n, 17)
    ups = upstream(n, V, 12)
    layer = SMPL(m, precision=precision).to(dev)
    got = gpu_grads(layer, dev, b, p, c, ups)
    ref = oracle_grads(m, b, p, c, ups)
    for name, g, r in zip(("g_betas", "g_pose", "g_cam"), got, ref):
        g, r = g.cpu().double(), r.double()
        err, scale = (g - r).abs().max().item(), r.abs().max().item()
        assert err <= rtol * scale + GRAD_ATOL, f"{precision} n={n}: {name} err {err:.3e} scale {scale:.3e}"


def test_backward_explicit_fp32_large_batch_uses_3xtf32(dev):
    """precision='fp32' at >= 256 bodies: FMA recompute, 3xTF32 tcgen05 blendshape gradient."""
    m = synthetic.make_model(9, num_verts=1200)
    n = 260
    b, p, c = synthetic.make_inputs(n, 33)
    ups = upstream(n, 1200, 18)
    layer = SMPL(m, precision="fp32").to(dev)
    got = gpu_grads(layer, dev, b, p, c, ups)
    ref = oracle_grads(m, b, p, c, ups)
    assert_grads(got, ref, "fp32 n=260")


def test_backward_large_batch_default_path(dev):
    """n = 520 under AUTO: tcgen05 recompute and tcgen05 blendshape backward, 4 persistent bodies per CTA."""
    m = synthetic.make_model(8, num_verts=1500)
    n = 520
    b, p, c = synthetic.make_inputs(n, 29)
    ups = upstream(n, 1500, 16)
    layer = SMPL(m).to(dev)
    got = gpu_grads(layer, dev, b, p, c, ups)
    ref = oracle_grads(m, b, p, c, ups)
    assert_grads(got, ref, "n=520 auto")


def test_backward_unstaged_large_mesh(dev):
    """V = 10000: one body's g_v + vposed exceed shared memory -> the non-staged skinning kernel."""
    m = synthetic.make_model(4, num_verts=10000)
    b, p, c = synthetic.make_inputs(2, 3)
    ups = upstream(2, 10000, 8)
    for joints in ("kinematic", "regressed"):
        layer = SMPL(m, precision="fp32", joints=joints).to(dev)
        got = gpu_grads(layer, dev, b, p, c, ups)
        ref = oracle_grads(m, b, p, c, ups, False, joints)
        assert_grads(got, ref, f"V=10000/{joints}")


def test_backward_zero_pose(dev, models):
    m = models["sparse"]
    b, p, c = synthetic.make_inputs(2, 1)
    p[:] = 0.0
    ups = upstream(2, 6890, 9)
    layer = SMPL(m, precision="fp32").to(dev)
    got = gpu_grads(layer, dev, b, p, c, ups)
    ref = oracle_grads(m, b, p, c, ups)
    assert_grads(got, ref, "zero pose")


def test_backward_is_bitwise_reproducible(dev, models):
    m = models["sparse"]
    b, p, c = synthetic.make_inputs(33, 2)
    ups = upstream(33, 6890, 10)
    for precision in ("bf16x3", "fp32"):      # pinned kernel paths (AUTO switches path with the batch size)
        layer = SMPL(m, precision=precision, lbs="fma").to(dev)
        g1 = gpu_grads(layer, dev, b, p, c, ups)
        g2 = gpu_grads(layer, dev, b, p, c, ups)
        assert all(torch.equal(a, b_) for a, b_ in zip(g1, g2))
        # sharding the batch does not change any body's gradient: fixed summation order per body, and
        # shards of <= 384 bodies all use the same number of column slices in the blendshape gradient
        g3 = gpu_grads(layer, dev, b[:16], p[:16], c[:16], [u[:16] for u in ups])
        assert all(torch.equal(a[:16], b_) for a, b_ in zip(g1, g3)), precision


@pytest.mark.parametrize("n", [9, 300])
def test_reusing_the_forward_workspace_changes_nothing(dev, models, n):
    """A and vposed kept from the forward == A and vposed recomputed: bitwise equal gradients."""
    m = models["sparse"] if n == 9 else synthetic.make_model(2, num_verts=1000)
    V = 6890 if n == 9 else 1000
    b, p, c = synthetic.make_inputs(n, 4)
    ups = upstream(n, V, 14)
    g_keep = gpu_grads(SMPL(m, save_forward_workspace=True).to(dev), dev, b, p, c, ups)
    g_re = gpu_grads(SMPL(m, save_forward_workspace=False).to(dev), dev, b, p, c, ups)
    assert all(torch.equal(x, y) for x, y in zip(g_keep, g_re))


def test_training_style_loss(dev, models):
    """A loss shaped like an HMR trainer's (L1 on 2-D keypoints + L2 on 3-D joints + priors)."""
    m = models["sparse"]
    n = 8
    b, p, c = synthetic.make_inputs(n, 13)
    rng = np.random.default_rng(5)
    tgt_k, tgt_j = rng.normal(size=(n, 24, 2)), rng.normal(size=(n, 24, 3)) * 0.3

    def loss_of(outs, tb, tp, dt, to):
        v, j, k = outs
        return ((k - to(tgt_k)).abs().mean() + ((j - to(tgt_j)) ** 2).mean()
                + 1e-3 * (tb ** 2).sum() + 1e-3 * (tp[:, 3:] ** 2).sum() + 1e-2 * v[:, ::50].pow(2).mean())

    layer = SMPL(m, precision="fp32").to(dev)
    tb, tp, tc = (torch.from_numpy(x).to(dev).requires_grad_() for x in (b, p, c))
    loss = loss_of(layer(tb, tp, tc), tb, tp, torch.float32, lambda a: torch.from_numpy(a.astype(np.float32)).to(dev))
    loss.backward()
    rb, rp, rc = (torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (b, p, c))
    ref_loss = loss_of(smpl_forward(m, rb, rp, rc, dtype=torch.float64), rb, rp, torch.float64,
                       lambda a: torch.from_numpy(a))
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item()) + 1e-6
    assert_grads((tb.grad, tp.grad, tc.grad), (rb.grad, rp.grad, rc.grad), "training loss")


def test_backward_full_size_properties(dev, models):
    """BASELINE.json's batch (4096 bodies, tcgen05 everywhere): size-independent properties of the
    gradient map, plus the float64 oracle on a slice of the batch."""
    m = models["sparse"]
    n = 4096
    b, p, c = synthetic.make_inputs(n, 61)
    layer = SMPL(m).to(dev)
    g = torch.Generator().manual_seed(9)
    uv, uj, uk = (torch.randn(n, d0, d1, generator=g).to(dev) for d0, d1 in ((6890, 3), (24, 3), (24, 2)))

    def grads(scale_v, scale_j, scale_k):
        tb, tp, tc = (torch.from_numpy(x).to(dev).requires_grad_() for x in (b, p, c))
        v, j, k = layer(tb, tp, tc)
        torch.autograd.backward([v, j, k], [uv * scale_v, uj * scale_j, uk * scale_k])
        torch.cuda.synchronize()
        return tb.grad, tp.grad, tc.grad

    g1 = grads(1.0, 1.0, 1.0)
    assert all(torch.isfinite(t).all() for t in g1)
    # homogeneity: scaling every upstream gradient by a power of two scales every sum exactly
    g2 = grads(2.0, 2.0, 2.0)
    assert all(torch.equal(2.0 * a, b_) for a, b_ in zip(g1, g2))
    # additivity over the outputs: g(v) + g(j) + g(k) == g(v, j, k) up to fp32 rounding
    gv, gj, gk = grads(1.0, 0.0, 0.0), grads(0.0, 1.0, 0.0), grads(0.0, 0.0, 1.0)
    for a, x, y, z in zip(g1, gv, gj, gk):
        assert (a - (x + y + z)).abs().max().item() <= 1e-5 * a.abs().max().item()
    # zero upstream -> zero gradient
    assert all(t.abs().max().item() == 0.0 for t in grads(0.0, 0.0, 0.0))
    # the oracle on the first 48 bodies (gradients are per body)
    k_ = 48
    ref = oracle_grads(m, b[:k_], p[:k_], c[:k_], [uv[:k_].cpu().numpy(), uj[:k_].cpu().numpy(), uk[:k_].cpu().numpy()])
    assert_grads(tuple(t[:k_] for t in g1), ref, "N=4096, first 48 bodies")


def test_train_step_is_cuda_graph_capturable(dev, models):
    """forward + loss + backward captured once and replayed: no host sync, no allocation outside the
    graph's pool, no stream-unsafe call anywhere in smplb200_forward / smplb200_backward."""
    m = models["sparse"]
    n = 16
    layer = SMPL(m).to(dev)
    b0, p0, c0 = synthetic.make_inputs(n, 31)
    b1, p1, c1 = synthetic.make_inputs(n, 32)
    sb, sp, sc = (torch.from_numpy(x).to(dev).requires_grad_() for x in (b0, p0, c0))

    def step():
        v, j, k = layer(sb, sp, sc)
        (v.square().mean() + j.square().mean() + k.abs().mean()).backward()

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            sb.grad = sp.grad = sc.grad = None
            step()
    torch.cuda.current_stream(dev).wait_stream(side)
    sb.grad = sp.grad = sc.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for arrs in ((b1, p1, c1), (b0, p0, c0)):
        with torch.no_grad():
            for t, a in zip((sb, sp, sc), arrs):
                t.copy_(torch.from_numpy(a))
        graph.replay()
        torch.cuda.synchronize()
        got = tuple(t.grad.clone() for t in (sb, sp, sc))
        eb, ep, ec = (torch.from_numpy(a).to(dev).requires_grad_() for a in arrs)
        v, j, k = layer(eb, ep, ec)
        (v.square().mean() + j.square().mean() + k.abs().mean()).backward()
        torch.cuda.synchronize()
        assert all(torch.equal(g, e.grad) for g, e in zip(got, (eb, ep, ec)))


def test_make_graphed_callables(dev, models):
    """torch.cuda.make_graphed_callables (PyTorch's API for graphing one module of an eager training
    loop) accepts the layer: forward and backward replay as two CUDA graphs, same gradients."""
    m = models["sparse"]
    n = 24
    layer = SMPL(m, precision="fp32", lbs="fma").to(dev)
    sample = tuple(torch.from_numpy(x).to(dev).requires_grad_() for x in synthetic.make_inputs(n, 51))
    graphed = torch.cuda.make_graphed_callables(layer, sample)
    arrs = synthetic.make_inputs(n, 52)
    outs_g, outs_e = [], []
    for mod, sink in ((graphed, outs_g), (layer, outs_e)):
        b, p, c = (torch.from_numpy(x).to(dev).requires_grad_() for x in arrs)
        v, j, k = mod(b, p, c)
        (v.square().mean() + j.square().mean() + k.abs().mean()).backward()
        torch.cuda.synchronize()
        sink.extend([v.detach().clone(), j.detach().clone(), k.detach().clone(), b.grad, p.grad, c.grad])
    assert all(torch.equal(a, b_) for a, b_ in zip(outs_g, outs_e))


def test_backward_errors_and_empty(dev, models):
    layer = SMPL(models["sparse"], precision="fp32").to(dev)
    h = layer.handle(dev)
    z = lambda *s: torch.zeros(*s, device=dev)
    b, p, gk = z(2, 10), z(2, 72), z(2, 24, 2)
    gb, gp = z(2, 10), z(2, 72)
    lib = capi.lib()
    # g_kp2d without cam
    st = lib.smplb200_backward(h.ptr, b.data_ptr(), p.data_ptr(), None, 2, None, None, None, gk.data_ptr(),
                               gb.data_ptr(), gp.data_ptr(), None, None, 0, None, 0, layer.flags, None)
    assert st == 1
    # vertex path without workspace
    gv = z(2, 6890, 3)
    st = lib.smplb200_backward(h.ptr, b.data_ptr(), p.data_ptr(), None, 2, None, gv.data_ptr(), None, None,
                               gb.data_ptr(), gp.data_ptr(), None, None, 0, None, 0, layer.flags, None)
    assert st == 3
    assert lib.smplb200_backward_workspace_bytes(h.ptr, 2, layer.flags, 1) > 2 * 3 * 6912 * 4
    assert lib.smplb200_backward_launch_count(h.ptr, 2, layer.flags, 1, 0) == 5
    assert lib.smplb200_backward_launch_count(h.ptr, 2, layer.flags, 1, 1) == 3
    assert lib.smplb200_backward_launch_count(h.ptr, 2, layer.flags, 0, 0) == 1
    # empty batch
    tb, tp = z(0, 10).requires_grad_(), z(0, 72).requires_grad_()
    v, j = layer(tb, tp)
    (v.sum() + j.sum()).backward()
    assert tb.grad.shape == (0, 10) and tp.grad.shape == (0, 72)
