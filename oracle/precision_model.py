"""TEST INFRASTRUCTURE -- numpy model of the OPERAND ROUNDING each tensor-core mode of the SMPL path applies.

Not a restatement of anything in the reference (the reference computes in fp32 throughout):
this file exists so that the looser vertex bounds DESIGN.md §4 states for the tcgen05 modes can be re-derived on
a CPU, from the arithmetic the kernels are documented to do, independently of any GPU run.  Everything but the
operand rounding is float64, so what comes out is the part of the error that is DESIGNED IN (operand bits); the
fp32 accumulation of the tensor core adds ~1e-6 m on top (measured).

Modes (csrc/k_blend_tc.cuh, csrc/k_lbs_tc.cuh, csrc/k_fused_tc.cuh; include/smpl_b200.h SMPLB200_PREC_*):
  'bf16' / 'tf32'     every blendshape operand rounded once (template carried as three pieces x 1.0)
  'bf16x3' / 'f16x3'  3-term split  hi*hi + hi*lo + lo*hi  of both blendshape operands
  'f16'               the fused kernel: pose rows ONE fp16 product, shape rows 3-term fp16 split, template three
                      fp16 pieces; skinning blend 3-term fp16 split of W and A
  skinning of the unfused modes: 3-term TF32 split of W and A (k_lbs_tc).

Only tests/ import this file.
"""
from __future__ import annotations

import numpy as np

from .smpl_np64 import rodrigues_expm


def _round_f16(x):
    return np.asarray(x, np.float64).astype(np.float16).astype(np.float64)


def _round_bf16(x):
    """fp32 -> bf16, round to nearest even (what the host packer and k2 do)."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def _trunc_tf32(x):
    """fp32 -> tf32 by dropping the low 13 mantissa bits (the tensor core ignores them)."""
    u = np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32).astype(np.float64)


ROUND = {"f16": _round_f16, "bf16": _round_bf16, "tf32": _trunc_tf32}


def split(x, kind, terms=2):
    """x (fp32 values) -> `terms` pieces of `kind` whose sum approximates x; piece k rounds the residual."""
    r = np.asarray(x, np.float32).astype(np.float64)
    out = []
    for _ in range(terms):
        p = ROUND[kind](r.astype(np.float32))
        out.append(p)
        r = r - p
    return out


def product(a, b, kind, terms):
    """`a @ b` with both operands rounded to `kind`: terms=1 -> hi*hi; terms=3 -> hi*hi + hi*lo + lo*hi."""
    ah, al = split(a, kind)
    bh, bl = split(b, kind)
    if terms == 1:
        return ah @ bh
    return ah @ bh + ah @ bl + al @ bh


def smpl_forward_mode(model, betas, pose, mode):
    """Vertices [N,V,3] (float64) with the operand rounding of `mode`; kinematic chain in float64 -> fp32 like k2."""
    f = lambda x: np.asarray(x, dtype=np.float64)
    vt, sd, pd = f(model["v_template"]), f(model["shapedirs"]), f(model["posedirs"])
    jr, w = f(model["J_regressor"]), f(model["weights"])
    parents = [int(p) for p in np.asarray(model["parents"]).astype(np.int64)]
    betas, pose = f(betas), f(pose)
    N, V, J = betas.shape[0], vt.shape[0], w.shape[1]

    # k2 (fp32 on the GPU, exact here): rotations, rest joints from the folded regressor, chain, A
    R = rodrigues_expm(pose.reshape(-1, 3)).reshape(N, J, 3, 3)
    pf = (R[:, 1:] - np.eye(3)).reshape(N, -1).astype(np.float32)
    v_shaped = vt[None] + (betas @ sd).reshape(N, V, 3)
    Jrest = np.einsum("nvc,vj->njc", v_shaped, jr)
    Rw = np.zeros((N, J, 3, 3))
    Jp = np.zeros((N, J, 3))
    Rw[:, 0], Jp[:, 0] = R[:, 0], Jrest[:, 0]
    for i in range(1, J):
        p = parents[i]
        Rw[:, i] = Rw[:, p] @ R[:, i]
        Jp[:, i] = Jp[:, p] + np.einsum("nab,nb->na", Rw[:, p], Jrest[:, i] - Jrest[:, p])
    A = np.concatenate([Rw, (Jp - np.einsum("njab,njb->nja", Rw, Jrest))[..., None]], axis=3)   # [N,J,3,4]
    A = A.astype(np.float32)

    # k1: v_posed = template + betas @ shapedirs + pose_feature @ posedirs
    kind = {"bf16": "bf16", "bf16x3": "bf16", "tf32": "tf32", "f16x3": "f16", "f16": "f16"}[mode]
    t_pieces = split(vt.reshape(-1), kind, terms=3)
    template = t_pieces[0] + t_pieces[1] + t_pieces[2]
    if mode == "f16":
        blend = product(betas.astype(np.float32), sd, kind, 3) + product(pf, pd, kind, 1)
    else:
        terms = 3 if mode.endswith("x3") else 1
        blend = product(betas.astype(np.float32), sd, kind, terms) + product(pf, pd, kind, terms)
    v_posed = (template[None] + blend).reshape(N, V, 3)

    # k3: T = W @ A (3-term split: fp16 in the fused kernel, TF32 in k_lbs_tc), then fp32 FMAs
    T = product(w, A.reshape(N, J * 12).reshape(N, J, 12).transpose(1, 0, 2).reshape(J, N * 12),
                "f16" if mode == "f16" else "tf32", 3)                                           # [V, N*12]
    T = T.reshape(V, N, 3, 4).transpose(1, 0, 2, 3)
    return np.einsum("nvab,nvb->nva", T[..., :3], v_posed) + T[..., 3]
