"""Fused decode -> gather: head maps -> the per-person (pose, shape, cam) vectors the SMPL layer eats.

One CUDA launch replaces the reference's `_nms` + `_topk` (src/lib/models/decode.py:6-41) and the
per-head `_transpose_and_gather_feat` (src/lib/models/utils.py:12-27) -- the steps
`multi_pose_decode` runs at src/lib/models/decode.py:80-88 -- without the NHWC permute-copy of every
head.  CUDA-only, forward-only, bit-exact against the reference for tie-free scores.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import capi


def decode_gather(heat: torch.Tensor, heads, K: int):
    """heat [B,C,H,W] (sigmoid-ed), heads: list of [B,ch,H,W] -> (scores, inds, clses, ys, xs, [gathered])."""
    if heat.device.type != "cuda":
        raise RuntimeError("decode_gather (B200) needs CUDA tensors; there is no CPU fallback")
    if heat.dtype != torch.float32 or heat.dim() != 4:
        raise TypeError("heat must be a float32 [B,C,H,W] tensor")
    heads = list(heads)
    if len(heads) > 8:
        raise ValueError("at most 8 heads per call")
    B, Cc, H, W = (int(x) for x in heat.shape)
    dev = heat.device
    heat = heat.contiguous()
    hs = []
    for h in heads:
        if h.device != dev or h.dtype != torch.float32 or h.dim() != 4 or h.shape[0] != B or tuple(h.shape[2:]) != (H, W):
            raise ValueError("every head must be float32 [B,ch,H,W] on the heat map's device")
        hs.append(h.contiguous())
    with torch.cuda.device(dev):
        scores = torch.empty((B, K), dtype=torch.float32, device=dev)
        inds = torch.empty((B, K), dtype=torch.int64, device=dev)
        clses = torch.empty((B, K), dtype=torch.int32, device=dev)
        ys = torch.empty((B, K), dtype=torch.float32, device=dev)
        xs = torch.empty((B, K), dtype=torch.float32, device=dev)
        outs = [torch.empty((B, K, int(h.shape[1])), dtype=torch.float32, device=dev) for h in hs]
        n = len(hs)
        src = (C.c_void_p * max(n, 1))(*[h.data_ptr() for h in hs])
        dst = (C.c_void_p * max(n, 1))(*[o.data_ptr() for o in outs])
        chs = (C.c_int32 * max(n, 1))(*[int(h.shape[1]) for h in hs])
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        capi.check(capi.lib().smplb200_decode_gather(
            idx, heat.data_ptr(), B, Cc, H, W, src, chs, n, int(K), scores.data_ptr(), inds.data_ptr(),
            clses.data_ptr(), ys.data_ptr(), xs.data_ptr(), dst, torch.cuda.current_stream(dev).cuda_stream),
            "smplb200_decode_gather")
    return scores, inds, clses, ys, xs, outs
