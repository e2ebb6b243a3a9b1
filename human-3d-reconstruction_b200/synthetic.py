"""Synthetic SMPL-shaped model tensors and per-person parameters.

The licensed SMPL model file and the datasets are not available offline
(SURVEY.md F1/F2, §8d), so every test and bench in this repo runs on a seeded
synthetic model with the real model's shapes and realistic magnitudes:

    v_template [V,3]      ~ U(-1,1) * (0.45, 0.9, 0.15) metres
    shapedirs  [NB,3V]    ~ N(0, 0.02^2)
    posedirs   [207,3V]   ~ N(0, 0.003^2)
    J_regressor[V,J]      columns >= 0, each sums to 1 (dense Dirichlet or sparse-nearest)
    weights    [V,J]      rows >= 0, each sums to 1 (dense softmax or <=4 nnz per vertex)
    parents    [J]        the standard 24-joint SMPL kinematic tree, root stored as -1

Buffer layouts are the ones the eager HMR-idiom layer keeps:
the flattened vertex axis of shapedirs/posedirs is ``3*v + c``.

The per-person parameter vectors mirror what the CenterNet-style heads of the
reference would regress (heads dict, reference src/lib/opts.py:248-258, gathered by
src/lib/models/utils.py:23-27): betas[N,10], pose[N,72], cam[N,3].
"""
from __future__ import annotations

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_FEATURES = 9 * (NUM_JOINTS - 1)  # 207

# kintree_table[0] of the public SMPL model; root is -1 here (2**32-1 in the pickle).
SMPL_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
    dtype=np.int32,
)


def make_model(seed: int = 0, num_verts: int = NUM_VERTS, num_betas: int = NUM_BETAS,
               weights: str = "sparse", regressor: str = "sparse") -> dict:
    """Build a seeded SMPL-shaped model dict of float32 numpy arrays.

    weights:   "sparse" (exactly <=4 non-zeros per vertex, like the real model) or "dense".
    regressor: "sparse" (each joint regressed from its ~32 nearest vertices) or "dense".
    """
    if weights not in ("sparse", "dense") or regressor not in ("sparse", "dense"):
        raise ValueError("weights/regressor must be 'sparse' or 'dense'")
    rng = np.random.default_rng(seed)
    V, J, NB, P = int(num_verts), NUM_JOINTS, int(num_betas), NUM_POSE_FEATURES
    scale = np.array([0.45, 0.9, 0.15])
    v_template = rng.uniform(-1.0, 1.0, size=(V, 3)) * scale
    shapedirs = rng.normal(0.0, 0.02, size=(NB, 3 * V))
    posedirs = rng.normal(0.0, 0.003, size=(P, 3 * V))

    # joint "centres" used only to give the sparse variants spatial coherence
    centres = rng.uniform(-0.8, 0.8, size=(J, 3)) * scale
    d2 = ((v_template[:, None, :] - centres[None, :, :]) ** 2).sum(-1)  # [V,J]

    if regressor == "dense":
        j_reg = rng.dirichlet(np.full(V, 0.5), size=J).T  # [V,J], columns sum to 1
    else:
        j_reg = np.zeros((V, J))
        k = min(32, V)
        for j in range(J):
            near = np.argpartition(d2[:, j], k - 1)[:k]
            w = rng.uniform(0.2, 1.0, size=k)
            j_reg[near, j] = w / w.sum()

    if weights == "dense":
        logits = rng.normal(0.0, 1.0, size=(V, J)) * 4.0
        logits -= logits.max(axis=1, keepdims=True)
        w = np.exp(logits)
        lbs_w = w / w.sum(axis=1, keepdims=True)
    else:
        lbs_w = np.zeros((V, J))
        nnz = rng.integers(1, 5, size=V)  # 1..4 influences per vertex
        order = np.argsort(d2, axis=1)[:, :4]  # 4 nearest joints
        raw = np.exp(-d2[np.arange(V)[:, None], order] * 8.0) + 1e-3
        raw *= (np.arange(4)[None, :] < nnz[:, None])
        raw /= raw.sum(axis=1, keepdims=True)
        lbs_w[np.arange(V)[:, None], order] = raw

    f32 = np.float32
    model = {
        "v_template": v_template.astype(f32),
        "shapedirs": shapedirs.astype(f32),
        "posedirs": posedirs.astype(f32),
        "J_regressor": j_reg.astype(f32),
        "weights": lbs_w.astype(f32),
        "parents": SMPL_PARENTS.copy(),
    }
    # renormalise after the float32 cast so rows/columns sum to 1 as tightly as fp32 allows
    model["weights"] /= model["weights"].sum(axis=1, keepdims=True, dtype=np.float64).astype(f32)
    model["J_regressor"] /= model["J_regressor"].sum(axis=0, keepdims=True, dtype=np.float64).astype(f32)
    return model


def make_inputs(n: int, seed: int = 1, num_betas: int = NUM_BETAS, edge_rows: bool = True):
    """betas[N,NB] ~ N(0,1) clipped +-3; pose[N,72] ~ N(0,0.3^2) rad; cam = (s, tx, ty).

    With ``edge_rows`` the first rows (as far as N allows) are the Rodrigues edge cases of
    SURVEY.md §8d: an all-zero pose, a pose with |theta| < 1e-6, and one joint with |theta| ~ pi.
    """
    rng = np.random.default_rng(seed)
    betas = np.clip(rng.normal(0.0, 1.0, size=(n, num_betas)), -3.0, 3.0)
    pose = rng.normal(0.0, 0.3, size=(n, 3 * NUM_JOINTS))
    cam = np.concatenate(
        [rng.uniform(0.5, 1.5, size=(n, 1)), rng.uniform(-0.5, 0.5, size=(n, 2))], axis=1)
    if edge_rows:
        if n > 0:
            pose[0] = 0.0
        if n > 1:
            pose[1] = rng.normal(0.0, 1.0, size=72) * 2e-7
        if n > 2:
            axis = np.array([0.6, -0.48, 0.64])
            pose[2, 3 * 5:3 * 5 + 3] = axis * (np.pi - 1e-4)
    f32 = np.float32
    return betas.astype(f32), pose.astype(f32), cam.astype(f32)
