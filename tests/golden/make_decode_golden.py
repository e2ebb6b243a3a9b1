"""Generates tests/golden/decode_golden_v1.npz from the UNMODIFIED reference functions.

Runs only in the build container (needs /root/reference):  python tests/golden/make_decode_golden.py
Imports reference src/lib/models/decode.py (`_nms`, `_topk`) and src/lib/models/utils.py
(`_transpose_and_gather_feat`) read-only and records their outputs on seeded inputs; the inputs are
regenerated from the seeds by `decode_case()` so only the outputs are stored.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [  # (seed, B, C, H, W, head channels, K)
    (101, 2, 1, 32, 32, (72, 10, 3), 8),
    (102, 3, 3, 24, 40, (5,), 16),
    (103, 1, 1, 128, 128, (72, 10, 3), 32),
    (104, 2, 2, 17, 13, (4, 1), 5),
]


def decode_case(seed, B, C, H, W, chans, K):
    """Seeded inputs: sigmoid-ed heat map in (0,1) with distinct values, random head maps."""
    rng = np.random.default_rng(seed)
    heat = 1.0 / (1.0 + np.exp(-rng.normal(0.0, 2.0, size=(B, C, H, W))))
    heads = [rng.normal(0.0, 1.0, size=(B, ch, H, W)).astype(np.float32) for ch in chans]
    return heat.astype(np.float32), heads


def main():
    sys.path.insert(0, "/root/reference/src/lib")
    from models.decode import _nms, _topk                      # noqa: E402  (the real reference)
    from models.utils import _transpose_and_gather_feat        # noqa: E402
    out = {}
    for ci, (seed, B, C, H, W, chans, K) in enumerate(CASES):
        heat, heads = decode_case(seed, B, C, H, W, chans, K)
        th = torch.from_numpy(heat)
        scores, inds, clses, ys, xs = _topk(_nms(th), K=K)
        out[f"c{ci}_scores"], out[f"c{ci}_inds"] = scores.numpy(), inds.numpy()
        out[f"c{ci}_clses"], out[f"c{ci}_ys"], out[f"c{ci}_xs"] = clses.numpy(), ys.numpy(), xs.numpy()
        for hi, h in enumerate(heads):
            out[f"c{ci}_head{hi}"] = _transpose_and_gather_feat(torch.from_numpy(h), inds).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decode_golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
