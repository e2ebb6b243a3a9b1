"""Generates tests/golden/dcn_golden_v1.npz with torchvision.ops.deform_conv2d (CPU, float64).

The reference's own extension cannot be compiled on torch 2.x (THC headers), so the golden outputs
come from torchvision's independent implementation of the same operator (same offset/mask channel
convention as reference src/lib/models/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:137-190).
Run:  python tests/golden/make_dcn_golden.py
"""
import os

import numpy as np
import torch
from torchvision.ops import deform_conv2d

CASES = [  # B, Ci, Co, H, W, offset scale
    (2, 32, 16, 9, 11, 1.5),
    (1, 32, 32, 12, 12, 3.0),
    (2, 32, 48, 7, 5, 8.0),     # large offsets: many samples fall outside the map
]

out = {}
g = torch.Generator().manual_seed(317)
for n, (B, Ci, Co, H, W, sc) in enumerate(CASES):
    f32 = lambda t: t.float().double()      # inputs are stored as float32: round them BEFORE computing y
    x = f32(torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64))
    w = f32(torch.randn(Co, Ci, 3, 3, generator=g, dtype=torch.float64) / (Ci * 9) ** 0.5)
    b = f32(torch.randn(Co, generator=g, dtype=torch.float64))
    off = f32(torch.randn(B, 18, H, W, generator=g, dtype=torch.float64) * sc)
    m = f32(torch.rand(B, 9, H, W, generator=g, dtype=torch.float64))
    y = deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m)
    for k, v in (("x", x), ("w", w), ("b", b), ("off", off), ("m", m), ("y", y)):
        out[f"c{n}_{k}"] = v.numpy().astype(np.float32 if k != "y" else np.float64)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dcn_golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
