"""CPU oracle for the SMPL forward hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED.  The mounted snapshot of Aaron20127/human-3d-reconstruction contains no
SMPL layer at all (`grep -ri 'smpl|rodrigues|shapedirs|posedirs'` over
/root/reference hits only image-blend helpers), ships no tests or golden vectors for this
path and imports no third-party SMPL package.  This file is therefore
a restatement of the *published* SMPL formulation (Loper et al., SIGGRAPH Asia 2015) in the
eager-PyTorch "HMR layer" idiom that BASELINE.json's north_star vocabulary describes
(SURVEY.md Appendix A.1-A.8), written from that specification.  It is pinned only against
(a) an independent float64 numpy restatement using different formulas (`oracle/smpl_np64.py`)
and (b) analytic known-answer tests (tests/test_oracle.py).  Anchors that DO exist in the
reference and that this layer's calling convention follows:
  * the caller that would feed it: gather of per-person head vectors,
    reference src/lib/models/utils.py:12-27 and top-K decode src/lib/models/decode.py:26-41;
  * the heads dict that would carry pose72/shape10/cam3: src/lib/opts.py:248-258,
    src/lib/models/model.py:450-473;
  * nn.Module-over-native-op boundary: src/lib/models/DCNv2/dcn_v2.py:57-128.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg
may import this module.  The product path (human-3d-reconstruction_b200/) never does.

All functions are dtype-generic (float32 is "the reference's CPU path", float64 the arbiter).
"""
from __future__ import annotations

import torch


def quat_to_rotmat(quat: torch.Tensor) -> torch.Tensor:
    """[M,4] (w,x,y,z) -> [M,3,3]; normalises first."""
    q = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    rot = torch.stack(
        [w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
         2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
         2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1)
    return rot.view(-1, 3, 3)


def batch_rodrigues(theta: torch.Tensor) -> torch.Tensor:
    """Axis-angle [M,3] -> rotation [M,3,3] through a half-angle quaternion."""
    angle = torch.norm(theta + 1e-8, p=2, dim=1, keepdim=True)
    axis = theta / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    return quat_to_rotmat(quat)


def batch_global_rigid_transformation(Rs, Js, parents, rotate_base=False):
    """Kinematic chain.

    Rs [N,J,3,3], Js [N,J,3] -> (J_posed [N,J,3], A [N,J,4,4]) where A has the rest pose
    removed: rotation of G_j, translation t_j - R_j J_j.
    """
    N, J = Rs.shape[0], Rs.shape[1]
    dt = Rs.dtype
    root = Rs[:, 0]
    if rotate_base:
        flip = torch.tensor([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=dt)
        root = torch.matmul(root, flip)
    Js = Js.unsqueeze(-1)  # [N,J,3,1]

    def make_G(R, t):  # R [N,3,3], t [N,3,1] -> [N,4,4]
        top = torch.cat([R, t], dim=2)
        bottom = torch.zeros(N, 1, 4, dtype=dt)
        bottom[:, :, 3] = 1
        return torch.cat([top, bottom], dim=1)

    chain = [make_G(root, Js[:, 0])]
    for i in range(1, J):
        p = int(parents[i])
        local = make_G(Rs[:, i], Js[:, i] - Js[:, p])
        chain.append(torch.matmul(chain[p], local))
    G = torch.stack(chain, dim=1)  # [N,J,4,4]
    J_posed = G[:, :, :3, 3]
    J_h = torch.cat([Js, torch.zeros(N, J, 1, 1, dtype=dt)], dim=2)  # [N,J,4,1]
    init_bone = torch.matmul(G, J_h)  # [N,J,4,1]
    init_bone = torch.nn.functional.pad(init_bone, (3, 0))  # -> [N,J,4,4], only last column set
    A = G - init_bone
    return J_posed, A


def smpl_forward(model: dict, betas, pose, cam=None, *, dtype=torch.float32,
                 rotate_base: bool = False, joints_from: str = "kinematic",
                 return_intermediates: bool = False):
    """Eager SMPL forward.

    model: dict with v_template[V,3], shapedirs[NB,3V], posedirs[207,3V], J_regressor[V,J],
           weights[V,J], parents[J] (numpy or torch).
    Returns (vertices[N,V,3], joints[N,J,3]) and kp2d[N,J,2] when cam is given.
    """
    t = lambda a: torch.as_tensor(a).to(dtype)
    v_template, shapedirs, posedirs = t(model["v_template"]), t(model["shapedirs"]), t(model["posedirs"])
    j_reg, weights = t(model["J_regressor"]), t(model["weights"])
    parents = [int(p) for p in torch.as_tensor(model["parents"]).to(torch.int64).tolist()]
    betas, pose = t(betas), t(pose)
    N, V, J = betas.shape[0], v_template.shape[0], weights.shape[1]

    # A.2 shape blend
    v_shaped = torch.matmul(betas, shapedirs).view(N, V, 3) + v_template
    # A.3 joint regression, one coordinate at a time
    Jx = torch.matmul(v_shaped[:, :, 0], j_reg)
    Jy = torch.matmul(v_shaped[:, :, 1], j_reg)
    Jz = torch.matmul(v_shaped[:, :, 2], j_reg)
    Jrest = torch.stack([Jx, Jy, Jz], dim=2)  # [N,J,3]
    # A.4 Rodrigues
    Rs = batch_rodrigues(pose.reshape(-1, 3)).view(N, J, 3, 3)
    # A.5 pose blend
    pose_feature = (Rs[:, 1:] - torch.eye(3, dtype=dtype)).reshape(N, 9 * (J - 1))
    v_posed = torch.matmul(pose_feature, posedirs).view(N, V, 3) + v_shaped
    # A.6 chain
    J_posed, A = batch_global_rigid_transformation(Rs, Jrest, parents, rotate_base=rotate_base)
    # A.7 linear blend skinning (materialises T exactly like the eager idiom)
    T = torch.matmul(weights, A.reshape(N, J, 16)).view(N, V, 4, 4)
    v_h = torch.cat([v_posed, torch.ones(N, V, 1, dtype=dtype)], dim=2)
    verts = torch.matmul(T, v_h.unsqueeze(-1))[:, :, :3, 0]
    # A.6/A.7 joint output
    if joints_from == "kinematic":
        joints = J_posed
    elif joints_from == "regressed":
        joints = torch.stack([torch.matmul(verts[:, :, c], j_reg) for c in range(3)], dim=2)
    else:
        raise ValueError("joints_from must be 'kinematic' or 'regressed'")
    out = [verts, joints]
    # A.8 weak-perspective projection
    if cam is not None:
        cam = t(cam)
        out.append(cam[:, None, 0:1] * (joints[:, :, :2] + cam[:, None, 1:3]))
    if return_intermediates:
        out.append({"v_shaped": v_shaped, "J_rest": Jrest, "Rs": Rs, "pose_feature": pose_feature,
                    "v_posed": v_posed, "A": A, "J_posed": J_posed})
    return tuple(out)


def smpl_forward_chunked(model, betas, pose, cam=None, *, chunk=1024, **kw):
    """Same as smpl_forward but bounds the eager T[N,V,4,4] intermediate (441 KB/body)."""
    outs = []
    n = len(betas)
    for s in range(0, n, chunk):
        c = None if cam is None else cam[s:s + chunk]
        outs.append(smpl_forward(model, betas[s:s + chunk], pose[s:s + chunk], c, **kw))
    return tuple(torch.cat([o[i] for o in outs], dim=0) for i in range(len(outs[0])))
