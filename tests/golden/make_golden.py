"""Generates tests/golden/smpl_golden_v1.npz.

PARITY UNPINNED: the reference snapshot has no SMPL layer and no fixtures for this path
(SURVEY.md F1, §8c), so these vectors come from OUR oracle (oracle/smpl_ref.py) run in float64
on the seeded synthetic model -- they pin the oracle and the CUDA kernels against regressions,
not against upstream.  Run from the repo root:  python tests/golden/make_golden.py

Only the seeds and the float64 outputs on a fixed vertex subset are stored (the model itself is
regenerated from its seed, 19 MB would not belong in git).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from human_3d_reconstruction_b200 import synthetic  # noqa: E402
from oracle.smpl_ref import smpl_forward  # noqa: E402

MODEL_SEED, INPUT_SEED, N = 0, 1, 4
VERT_IDX = np.arange(0, synthetic.NUM_VERTS, 53)  # 130 vertices spread over all tiles


def main():
    out = {"model_seed": MODEL_SEED, "input_seed": INPUT_SEED, "n": N, "vert_idx": VERT_IDX}
    for wmode in ("sparse", "dense"):
        model = synthetic.make_model(MODEL_SEED, weights=wmode)
        betas, pose, cam = synthetic.make_inputs(N, INPUT_SEED)
        for rb in (False, True):
            for jf in ("kinematic", "regressed"):
                v, j, k = smpl_forward(model, betas, pose, cam, dtype=torch.float64,
                                       rotate_base=rb, joints_from=jf)
                tag = f"{wmode}_rb{int(rb)}_{jf}"
                out[f"verts_{tag}"] = v.numpy()[:, VERT_IDX]
                out[f"joints_{tag}"] = j.numpy()
                out[f"kp2d_{tag}"] = k.numpy()
    out["betas"], out["pose"], out["cam"] = betas, pose, cam
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "smpl_golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
