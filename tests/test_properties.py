"""Property tests of the host-side logic (CPU, hypothesis)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from human_3d_reconstruction_b200 import capi, sharding, synthetic
from oracle.smpl_backward_np import rodrigues_backward


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 100000), world=st.integers(1, 64))
def test_shard_bounds_partition_the_batch(n, world):
    """Contiguous, disjoint, complete, balanced to within one ceil-chunk, empty tails allowed."""
    prev_end, sizes = 0, []
    for r in range(world):
        lo, hi = sharding.shard_bounds(n, world, r)
        assert lo == prev_end and lo <= hi <= n
        prev_end = hi
        sizes.append(hi - lo)
    assert prev_end == n and sum(sizes) == n
    chunk = -(-n // world)
    assert max(sizes) <= chunk and all(s in (0, chunk) or i == max(j for j, t in enumerate(sizes) if t) for i, s in enumerate(sizes))


@settings(max_examples=60, deadline=None)
@given(p=st.sampled_from(list(capi.PRECISIONS)), l=st.sampled_from(list(capi.LBS_PATHS)),
       j=st.sampled_from(["kinematic", "regressed"]), rb=st.booleans())
def test_flag_fields_do_not_overlap(p, l, j, rb):
    f = capi.make_flags(p, j, rb, l)
    assert f & capi.PREC_MASK == capi.PRECISIONS[p]
    assert f & (3 << 5) == capi.LBS_PATHS[l]
    assert bool(f & capi.JOINTS_REGRESSED) == (j == "regressed") and bool(f & capi.ROTATE_BASE) == rb
    assert f < (1 << 7)


@settings(max_examples=50, deadline=None)
@given(seed=st.integers(0, 10_000), scale=st.floats(1e-3, 3.0))
def test_rodrigues_backward_is_the_transpose_of_the_forward_jacobian(seed, scale):
    """<g, dR> == <rodrigues_backward(theta, g), dtheta> for small dtheta (directional derivative)."""
    rng = np.random.default_rng(seed)
    theta = rng.normal(size=3) * scale
    g = rng.normal(size=(3, 3))
    d = rng.normal(size=3)

    def R(t):
        a = np.linalg.norm(t)
        k = t / a
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K

    h = 1e-6
    lhs = ((R(theta + h * d) - R(theta - h * d)) / (2 * h) * g).sum()
    rhs = rodrigues_backward(theta, g) @ d
    assert abs(lhs - rhs) <= 1e-5 * (1 + abs(lhs))


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 1000), nv=st.integers(24, 400))
def test_synthetic_model_invariants(seed, nv):
    m = synthetic.make_model(seed, num_verts=nv)
    w, jr = np.asarray(m["weights"]), np.asarray(m["J_regressor"])
    assert w.shape == (nv, 24) and np.allclose(w.sum(1), 1.0, atol=1e-5) and (w >= 0).all()
    assert np.allclose(jr.sum(0), 1.0, atol=1e-4)
    par = np.asarray(m["parents"]).astype(np.int64)
    assert all(0 <= par[j] < j for j in range(1, 24))
