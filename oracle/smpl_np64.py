"""Independent float64 numpy restatement of SMPL -- TEST INFRASTRUCTURE (checker of the oracle).

PARITY UNPINNED (see oracle/smpl_ref.py header: the reference snapshot has no SMPL code).
This file deliberately uses *different* textbook formulas from oracle/smpl_ref.py so that the
two only agree if both implement the published model (Loper et al. 2015):
  * Rodrigues as the matrix exponential  R = I + sin(a) K + (1 - cos(a)) K^2  (not a quaternion);
  * forward kinematics as explicit world rotations / joint positions
        Rw_j = Rw_p R_j,   Jp_j = Jp_p + Rw_p (J_j - J_p)      (no 4x4 matrices);
  * skinning as "rotate each vertex about each posed joint"
        v' = sum_j w_vj ( Rw_j (v - J_j) + Jp_j )              (no rest-pose-removed A, no T).
Pure numpy loops over joints; only used on small N in tests.
"""
from __future__ import annotations

import numpy as np


def rodrigues_expm(theta):
    """[M,3] -> [M,3,3] via I + sin(a) K + (1-cos(a)) K^2, K the unit-axis cross matrix."""
    theta = np.asarray(theta, dtype=np.float64)
    a = np.linalg.norm(theta, axis=1)
    safe = np.where(a > 0, a, 1.0)
    k = theta / safe[:, None]
    K = np.zeros((theta.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(a)[:, None, None], np.cos(a)[:, None, None]
    return np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)


def smpl_forward_np64(model, betas, pose, cam=None, rotate_base=False, joints_from="kinematic"):
    f = lambda x: np.asarray(x, dtype=np.float64)
    vt, sd, pd = f(model["v_template"]), f(model["shapedirs"]), f(model["posedirs"])
    jr, w = f(model["J_regressor"]), f(model["weights"])
    parents = [int(p) for p in np.asarray(model["parents"]).astype(np.int64)]
    betas, pose = f(betas), f(pose)
    N, V, J = betas.shape[0], vt.shape[0], w.shape[1]

    v_shaped = vt[None] + np.einsum("nb,bvc->nvc", betas, sd.reshape(-1, V, 3))
    Jrest = np.einsum("nvc,vj->njc", v_shaped, jr)
    R = rodrigues_expm(pose.reshape(-1, 3)).reshape(N, J, 3, 3)
    pf = (R[:, 1:] - np.eye(3)).reshape(N, -1)
    v_posed = v_shaped + np.einsum("np,pvc->nvc", pf, pd.reshape(-1, V, 3))

    Rw = np.zeros((N, J, 3, 3))
    Jp = np.zeros((N, J, 3))
    root = R[:, 0]
    if rotate_base:
        root = root @ np.diag([1.0, -1.0, -1.0])
    Rw[:, 0], Jp[:, 0] = root, Jrest[:, 0]
    for i in range(1, J):
        p = parents[i]
        Rw[:, i] = Rw[:, p] @ R[:, i]
        Jp[:, i] = Jp[:, p] + np.einsum("nab,nb->na", Rw[:, p], Jrest[:, i] - Jrest[:, p])

    verts = np.zeros((N, V, 3))
    for j in range(J):
        moved = np.einsum("nab,nvb->nva", Rw[:, j], v_posed - Jrest[:, None, j]) + Jp[:, None, j]
        verts += w[None, :, j, None] * moved
    if joints_from == "kinematic":
        joints = Jp
    else:
        joints = np.einsum("nvc,vj->njc", verts, jr)
    out = [verts, joints]
    if cam is not None:
        cam = f(cam)
        out.append(cam[:, None, 0:1] * (joints[:, :, :2] + cam[:, None, 1:3]))
    return tuple(out)
