"""CPU tests of the oracle itself (no GPU): the two independent restatements agree, the analytic
known-answer tests of SURVEY.md App. A.9 hold, and the committed golden vectors reproduce.

PARITY UNPINNED: the reference ships no SMPL code or fixtures; these tests
pin the oracle to the published formulation via a second, differently-formulated restatement.
"""
import os

import numpy as np
import pytest
import torch

from human_3d_reconstruction_b200 import synthetic
from oracle.smpl_np64 import rodrigues_expm, smpl_forward_np64
from oracle.smpl_ref import batch_rodrigues, smpl_forward

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smpl_golden_v1.npz")


@pytest.mark.parametrize("rotate_base", [False, True])
@pytest.mark.parametrize("joints_from", ["kinematic", "regressed"])
def test_two_restatements_agree_fp64(small_model, rotate_base, joints_from):
    betas, pose, cam = synthetic.make_inputs(9, 5)
    a = smpl_forward(small_model, betas, pose, cam, dtype=torch.float64,
                     rotate_base=rotate_base, joints_from=joints_from)
    b = smpl_forward_np64(small_model, betas, pose, cam, rotate_base=rotate_base, joints_from=joints_from)
    for x, y in zip(a, b):
        # limited only by the 1e-8 guard inside the HMR-idiom Rodrigues
        np.testing.assert_allclose(x.numpy(), y, rtol=0, atol=5e-8)


def test_rodrigues_matches_matrix_exponential():
    rng = np.random.default_rng(0)
    theta = rng.normal(0, 1.0, size=(200, 3))
    theta[0] = 0
    theta[1] = [np.pi - 1e-6, 0, 0]
    R = batch_rodrigues(torch.from_numpy(theta)).numpy()
    np.testing.assert_allclose(R, rodrigues_expm(theta), atol=5e-8)
    # proper rotations
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), R.shape), atol=1e-12)
    np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-12)


def test_kat_zero_pose_is_shape_only(small_model):
    """A.9(i)/(ii): pose = 0 => verts == v_shaped; betas = 0 too => verts == v_template."""
    n = 3
    betas, _, _ = synthetic.make_inputs(n, 2)
    pose = np.zeros((n, 72), np.float32)
    v, j, inter = smpl_forward(small_model, betas, pose, dtype=torch.float64, return_intermediates=True)
    np.testing.assert_allclose(v.numpy(), inter["v_shaped"].numpy(), atol=1e-7)
    np.testing.assert_allclose(j.numpy(), inter["J_rest"].numpy(), atol=1e-7)
    v0, _ = smpl_forward(small_model, np.zeros_like(betas), pose, dtype=torch.float64)
    np.testing.assert_allclose(v0.numpy(), np.broadcast_to(small_model["v_template"], v0.shape), atol=1e-7)


def test_kat_root_rotation_is_rigid_about_root_joint(small_model):
    """A.9(iii): rotating only the root by theta about z moves every vertex rigidly about J_0."""
    betas, _, _ = synthetic.make_inputs(2, 3)
    pose = np.zeros((2, 72), np.float32)
    ang = 0.7
    pose[:, 2] = ang
    v, j, inter = smpl_forward(small_model, betas, pose, dtype=torch.float64, return_intermediates=True)
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    vs, J0 = inter["v_shaped"].numpy(), inter["J_rest"].numpy()[:, :1]
    # pose blendshapes are zero because only the root rotates (pose_feature skips joint 0)
    np.testing.assert_allclose(v.numpy(), (vs - J0) @ Rz.T + J0, atol=1e-7)


def test_kat_one_hot_weights_follow_their_joint(small_model):
    """A.9(iv): with one-hot skinning weights a vertex follows exactly its joint's transform."""
    m = dict(small_model)
    V = m["v_template"].shape[0]
    w = np.zeros((V, 24), np.float32)
    owner = np.arange(V) % 24
    w[np.arange(V), owner] = 1.0
    m["weights"] = w
    betas, pose, _ = synthetic.make_inputs(3, 4)
    v, _, inter = smpl_forward(m, betas, pose, dtype=torch.float64, return_intermediates=True)
    A = inter["A"].numpy()[:, owner]  # [N,V,4,4]
    vp = np.concatenate([inter["v_posed"].numpy(), np.ones((3, V, 1))], axis=2)
    np.testing.assert_allclose(v.numpy(), np.einsum("nvab,nvb->nva", A, vp)[:, :, :3], atol=1e-12)


def test_kat_equal_transforms_make_weights_irrelevant(small_model):
    """A.9(v): rows of W sum to 1, so if only the root moves the result does not depend on W."""
    betas, _, _ = synthetic.make_inputs(2, 6)
    pose = np.zeros((2, 72), np.float32)
    pose[:, :3] = [0.3, -0.2, 0.5]
    dense = synthetic.make_model(3, num_verts=300, weights="dense")
    sparse = synthetic.make_model(3, num_verts=300, weights="sparse")
    a, _ = smpl_forward(dense, betas, pose, dtype=torch.float64)
    b, _ = smpl_forward(sparse, betas, pose, dtype=torch.float64)
    np.testing.assert_allclose(a.numpy(), b.numpy(), atol=2e-7)  # fp32 weight rows sum to 1 +- 1e-7


def test_kat_identity_camera(small_model):
    """A.9(vii): cam = (1, 0, 0) => kp2d == joints_xy."""
    betas, pose, _ = synthetic.make_inputs(4, 7)
    cam = np.tile(np.array([[1.0, 0.0, 0.0]], np.float32), (4, 1))
    _, j, k = smpl_forward(small_model, betas, pose, cam, dtype=torch.float64)
    np.testing.assert_array_equal(k.numpy(), j.numpy()[:, :, :2])


def test_fp32_oracle_close_to_fp64(small_model):
    betas, pose, cam = synthetic.make_inputs(16, 8)
    a = smpl_forward(small_model, betas, pose, cam, dtype=torch.float32)
    b = smpl_forward(small_model, betas, pose, cam, dtype=torch.float64)
    for x, y in zip(a, b):
        assert (x.double() - y).abs().max().item() < 2e-6


def test_shard_equivalence_of_oracle(small_model):
    """A.9(viii): bodies are independent -- forward(N) == concat(forward(shards)) bit for bit."""
    betas, pose, cam = synthetic.make_inputs(10, 9)
    full = smpl_forward(small_model, betas, pose, cam, dtype=torch.float64)
    parts = [smpl_forward(small_model, betas[s], pose[s], cam[s], dtype=torch.float64)
             for s in (slice(0, 4), slice(4, 10))]
    for i in range(3):
        np.testing.assert_allclose(full[i].numpy(), torch.cat([p[i] for p in parts]).numpy(), atol=1e-15)


def test_golden_vectors_reproduce():
    g = np.load(GOLDEN)
    idx = g["vert_idx"]
    for wmode in ("sparse", "dense"):
        model = synthetic.make_model(int(g["model_seed"]), weights=wmode)
        betas, pose, cam = synthetic.make_inputs(int(g["n"]), int(g["input_seed"]))
        np.testing.assert_array_equal(betas, g["betas"])
        np.testing.assert_array_equal(pose, g["pose"])
        for rb in (False, True):
            for jf in ("kinematic", "regressed"):
                v, j, k = smpl_forward(model, betas, pose, cam, dtype=torch.float64,
                                       rotate_base=rb, joints_from=jf)
                tag = f"{wmode}_rb{int(rb)}_{jf}"
                np.testing.assert_allclose(v.numpy()[:, idx], g[f"verts_{tag}"], atol=1e-12)
                np.testing.assert_allclose(j.numpy(), g[f"joints_{tag}"], atol=1e-12)
                np.testing.assert_allclose(k.numpy(), g[f"kp2d_{tag}"], atol=1e-12)
    # and the independent numpy restatement hits the same golden numbers
    model = synthetic.make_model(int(g["model_seed"]), weights="sparse")
    v, j, k = smpl_forward_np64(model, g["betas"], g["pose"], g["cam"])
    np.testing.assert_allclose(v[:, idx], g["verts_sparse_rb0_kinematic"], atol=5e-8)


def test_synthetic_model_invariants(full_model):
    m = full_model
    assert m["v_template"].shape == (6890, 3) and m["shapedirs"].shape == (10, 20670)
    assert m["posedirs"].shape == (207, 20670) and m["weights"].shape == (6890, 24)
    np.testing.assert_allclose(m["weights"].sum(1), 1.0, atol=3e-7)
    np.testing.assert_allclose(m["J_regressor"].sum(0), 1.0, atol=3e-6)
    assert (m["weights"] >= 0).all() and ((m["weights"] > 0).sum(1) <= 4).all()
    p = m["parents"]
    assert p[0] == -1 and all(0 <= p[i] < i for i in range(1, 24))
