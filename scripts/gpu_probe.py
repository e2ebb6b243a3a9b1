"""GPU bring-up probe: runs one kernel family per process and prints its error vs the CPU oracle.

Usage (on the GPU box): python scripts/gpu_probe.py <case> [n]
Each case runs in its own process so a faulting kernel cannot poison the others' CUDA context.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from human_3d_reconstruction_b200 import SMPL, synthetic, capi
from human_3d_reconstruction_b200 import smpl as ops
from oracle.smpl_ref import smpl_forward


def planar_to_interleaved(vp, V):
    return vp[:, :, :V].permute(0, 2, 1).contiguous()


def main():
    case = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 70
    dev = torch.device("cuda:0")
    wmode = "dense" if "dense" in case else "sparse"
    model = synthetic.make_model(0, weights=wmode)
    betas, pose, cam = synthetic.make_inputs(n, 1)
    t0 = time.time()
    ref = smpl_forward(model, betas, pose, cam, dtype=torch.float32, return_intermediates=True)
    ref_v, ref_j, ref_k, inter = ref
    tb, tp, tc = (torch.from_numpy(x).to(dev) for x in (betas, pose, cam))
    layer = SMPL(model).to(dev)
    V = layer.num_verts
    err = lambda a, b: (a.detach().cpu().double() - b.double()).abs().max().item()
    if case == "chain":
        coef, A, joints = ops.pose_chain(layer, tb, tp)
        torch.cuda.synchronize()
        NB = layer.num_betas
        print("coef betas", err(coef[:, :NB], torch.from_numpy(betas)))
        print("coef pf", err(coef[:, NB:NB + 207], inter["pose_feature"]))
        print("coef one/pad", coef[:, NB + 207:NB + 210].min().item(), coef[:, NB + 210:].abs().max().item())
        print("A", err(A.view(n, 24, 3, 4), inter["A"][:, :, :3, :]))
        print("joints", err(joints, ref_j))
    elif case.startswith("blend_"):
        prec = case.split("_")[1]
        coef, A, joints = ops.pose_chain(layer, tb, tp)
        vp = ops.blendshapes(layer, coef, flags=capi.make_flags(precision=prec))
        torch.cuda.synchronize()
        print(f"vposed[{prec}]", err(planar_to_interleaved(vp, V), inter["v_posed"]))
    elif case.startswith("lbs_"):
        path = case.split("_")[1]
        vp = torch.zeros((n, 3, layer.handle(dev).padded_verts), device=dev)
        vp[:, :, :V] = inter["v_posed"].permute(0, 2, 1).to(dev)
        A = inter["A"][:, :, :3, :].reshape(n, 24, 12).contiguous().to(dev)
        v, kp = ops.lbs(layer, vp, A, joints=ref_j.to(dev).contiguous(), cam=tc,
                        flags=capi.make_flags(lbs=path))
        torch.cuda.synchronize()
        print(f"verts[{path}]", err(v, ref_v), "kp2d", err(kp, ref_k))
    elif case.startswith("full_"):
        _, prec, path = case.split("_")[:3]
        layer = SMPL(model, precision=prec, lbs=path).to(dev)
        v, j, k = layer(tb, tp, tc)
        torch.cuda.synchronize()
        print(f"full[{prec},{path}] verts", err(v, ref_v), "joints", err(j, ref_j), "kp2d", err(k, ref_k))
    elif case == "regress":
        layer = SMPL(model, precision="fp32", lbs="fma", joints="regressed").to(dev)
        v, j, k = layer(tb, tp, tc)
        torch.cuda.synchronize()
        r = smpl_forward(model, betas, pose, cam, joints_from="regressed")
        print("regressed joints", err(j, r[1]), "kp2d", err(k, r[2]), "verts", err(v, r[0]))
    else:
        raise SystemExit(f"unknown case {case}")
    print(f"case {case} n={n} done in {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
