"""Hand-derived SMPL backward in float64 numpy -- TEST INFRASTRUCTURE (checker of the CUDA backward).

States, step by step, the reverse-mode algorithm the CUDA backward kernels implement
(human-3d-reconstruction_b200/csrc/k_backward.cuh) so that the derivation itself is pinned against
torch autograd of the oracle (tests/test_oracle_backward.py) independently of any GPU:

  1. projection      kp2d = s (J_xy + t)
  2. skinning        verts_v = sum_j w_vj ( Rw_j vp_v + ta_j ),  A_j = [Rw_j | ta_j]
  3. blendshapes     v_posed = coef . basis
  4. kinematic chain Rw_j = Rw_p R_j,  tw_j = Rw_p (Jr_j - Jr_p) + tw_p,  ta_j = tw_j - Rw_j Jr_j
  5. folded regressor Jr = J_template + betas . J_shapedirs
  6. Rodrigues       R_j = quat2mat(normalize([cos h, sin h * n])),  h = |theta + 1e-8| / 2

PARITY UNPINNED (like oracle/smpl_ref.py): the reference has no SMPL layer, so "the reference's
gradients" are torch autograd through our restatement of the published model.
"""
from __future__ import annotations

import numpy as np


def _quat_from_theta(theta):
    a = np.linalg.norm(theta + 1e-8)
    n = theta / a
    h = 0.5 * a
    q = np.concatenate([[np.cos(h)], np.sin(h) * n])
    return q / np.linalg.norm(q), a, n, h


def _rot_from_quat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z],
                     [2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x],
                     [2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z]])


def rodrigues_backward(theta, gR):
    """dL/dtheta (3) from dL/dR (3x3) for the HMR-idiom Rodrigues."""
    q, a, n, h = _quat_from_theta(theta)
    w, x, y, z = q
    g = gR
    gq = np.array([
        2 * w * (g[0, 0] + g[1, 1] + g[2, 2]) + 2 * (-z * g[0, 1] + y * g[0, 2] + z * g[1, 0] - x * g[1, 2] - y * g[2, 0] + x * g[2, 1]),
        2 * x * (g[0, 0] - g[1, 1] - g[2, 2]) + 2 * (y * g[0, 1] + z * g[0, 2] + y * g[1, 0] - w * g[1, 2] + z * g[2, 0] + w * g[2, 1]),
        2 * y * (-g[0, 0] + g[1, 1] - g[2, 2]) + 2 * (x * g[0, 1] + w * g[0, 2] + x * g[1, 0] + z * g[1, 2] - w * g[2, 0] + z * g[2, 1]),
        2 * z * (-g[0, 0] - g[1, 1] + g[2, 2]) + 2 * (-w * g[0, 1] + x * g[0, 2] + w * g[1, 0] + y * g[1, 2] + x * g[2, 0] + y * g[2, 1]),
    ])
    gq = gq - np.dot(gq, q) * q            # through q / |q| (|q| = 1)
    s, c = np.sin(h), np.cos(h)
    # q_w = cos h, q_v = sin h * n, h = a/2, n = theta/a  (a = |theta + eps| treated as |theta|)
    nnT = np.outer(n, n)
    dqv = 0.5 * c * nnT + (s / a) * (np.eye(3) - nnT)     # d q_v / d theta
    dqw = -0.5 * s * n                                     # d q_w / d theta
    return dqw * gq[0] + dqv.T @ gq[1:]


def smpl_backward_np(model, betas, pose, cam, gV=None, gJ=None, gK=None, rotate_base=False,
                     joints_from="kinematic"):
    """Returns (g_betas[N,NB], g_pose[N,72], g_cam[N,3]).

    joints_from='regressed': joints = J_regressor^T vertices, so g_joints / g_kp2d flow into the
    vertex gradient instead of the chain translations."""
    regressed = joints_from == "regressed"
    f = lambda x: np.asarray(x, dtype=np.float64)
    vt, sd, pd = f(model["v_template"]).reshape(-1), f(model["shapedirs"]), f(model["posedirs"])
    jr, W = f(model["J_regressor"]), f(model["weights"])
    parents = [int(p) for p in np.asarray(model["parents"]).astype(np.int64)]
    betas, pose = f(betas), f(pose)
    N, NB = betas.shape
    V, J = W.shape
    jt = (jr.T @ vt.reshape(V, 3))                                     # [J,3]   folded regressor
    jsd = np.einsum("vj,kvc->kjc", jr, sd.reshape(NB, V, 3))           # [NB,J,3]
    flip = np.diag([1.0, -1.0, -1.0])
    g_betas, g_pose = np.zeros((N, NB)), np.zeros((N, 3 * J))
    g_cam = np.zeros((N, 3))
    for b in range(N):
        # ---- forward recompute
        R = np.stack([_rot_from_quat(_quat_from_theta(pose[b, 3 * j:3 * j + 3])[0]) for j in range(J)])
        Jr = jt + np.einsum("k,kjc->jc", betas[b], jsd)
        pf = (R[1:] - np.eye(3)).reshape(-1)
        vp = (vt + betas[b] @ sd + pf @ pd).reshape(V, 3)
        Rw, tw = np.zeros((J, 3, 3)), np.zeros((J, 3))
        Rw[0], tw[0] = (R[0] @ flip if rotate_base else R[0]), Jr[0]
        for j in range(1, J):
            p = parents[j]
            Rw[j] = Rw[p] @ R[j]
            tw[j] = Rw[p] @ (Jr[j] - Jr[p]) + tw[p]
        ta = tw - np.einsum("jab,jb->ja", Rw, Jr)
        # ---- 1. projection
        gJb = np.zeros((J, 3)) if gJ is None else f(gJ[b]).copy()
        if gK is not None:
            s, t = cam[b, 0], f(cam[b, 1:3])
            gk = f(gK[b])
            if regressed:
                verts = np.einsum("vj,jab,vb->va", W, Rw, vp) + W @ ta
                jout = jr.T @ verts
            else:
                jout = tw
            g_cam[b, 0] = np.sum(gk * (jout[:, :2] + t))
            g_cam[b, 1:3] = s * gk.sum(0)
            gJb[:, :2] += s * gk
        # ---- 2. skinning
        gRw, gta = np.zeros((J, 3, 3)), np.zeros((J, 3))
        gvp = np.zeros((V, 3))
        gVb = None if gV is None else f(gV[b])
        if regressed:
            gVb = (0.0 if gVb is None else gVb) + jr @ gJb             # joints = J_regressor^T verts
            gJb = np.zeros((J, 3))
        if gVb is not None:
            g = gVb                                                    # [V,3]
            TR = np.einsum("vj,jab->vab", W, Rw)
            gvp = np.einsum("vab,va->vb", TR, g)                       # T_R^T g
            gRw += np.einsum("vj,va,vb->jab", W, g, vp)
            gta += W.T @ g
        # ---- 3. blendshapes
        gcoef_b = sd @ gvp.reshape(-1)                                 # [NB]
        gcoef_p = pd @ gvp.reshape(-1)                                 # [207]
        # ---- 4. chain (children before parents)
        gtw = gJb + gta                                                # joints = tw ; ta = tw - Rw Jr
        gRw -= np.einsum("ja,jb->jab", gta, Jr)
        gJr = -np.einsum("jab,ja->jb", Rw, gta)
        gR = np.zeros((J, 3, 3))
        gR[1:] += gcoef_p.reshape(J - 1, 3, 3)
        for j in range(J - 1, 0, -1):
            p = parents[j]
            d = Jr[j] - Jr[p]
            gRw[p] += gRw[j] @ R[j].T + np.outer(gtw[j], d)
            gR[j] += Rw[p].T @ gRw[j]
            gd = Rw[p].T @ gtw[j]
            gJr[j] += gd
            gJr[p] -= gd
            gtw[p] += gtw[j]
        gR[0] += gRw[0] @ flip.T if rotate_base else gRw[0]
        gJr[0] += gtw[0]
        # ---- 5/6. regressor + Rodrigues
        g_betas[b] = gcoef_b + np.einsum("jc,kjc->k", gJr, jsd)
        for j in range(J):
            g_pose[b, 3 * j:3 * j + 3] = rodrigues_backward(pose[b, 3 * j:3 * j + 3], gR[j])
    return g_betas, g_pose, g_cam
