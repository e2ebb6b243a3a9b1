"""CPU oracle for the decode -> gather producer of the SMPL path -- TEST INFRASTRUCTURE.

Restates, in plain torch, the reference functions that turn head maps into per-person parameter
vectors (the caller of the SMPL layer in inference):

  * `_nms`                       reference src/lib/models/decode.py:6-13
  * `_topk`                      reference src/lib/models/decode.py:26-41
  * `_gather_feat`               reference src/lib/models/utils.py:12-21
  * `_transpose_and_gather_feat` reference src/lib/models/utils.py:23-27

PARITY PINNED: unlike the SMPL layer these functions exist in the reference snapshot.
tests/golden/decode_golden_v1.npz holds outputs of the UNMODIFIED reference functions (generated in
the build container by tests/golden/make_decode_golden.py, which imports /root/reference) and
tests/test_decode_oracle.py checks this restatement against them bit for bit.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def nms(heat: torch.Tensor, kernel: int = 3) -> torch.Tensor:
    """Keep a value where it equals the max of its kernel x kernel neighbourhood, else heat*0."""
    pad = (kernel - 1) // 2
    hmax = F.max_pool2d(heat, (kernel, kernel), stride=1, padding=pad)
    keep = (hmax == heat).float()
    return heat * keep


def gather_feat(feat: torch.Tensor, ind: torch.Tensor) -> torch.Tensor:
    """feat [B, M, C], ind [B, K] -> [B, K, C]."""
    dim = feat.size(2)
    ind = ind.unsqueeze(2).expand(ind.size(0), ind.size(1), dim)
    return feat.gather(1, ind)


def transpose_and_gather_feat(feat: torch.Tensor, ind: torch.Tensor) -> torch.Tensor:
    """feat [B, C, H, W] (NCHW), ind [B, K] into H*W -> [B, K, C]."""
    feat = feat.permute(0, 2, 3, 1).contiguous()
    feat = feat.view(feat.size(0), -1, feat.size(3))
    return gather_feat(feat, ind)


def topk(scores: torch.Tensor, K: int):
    """Two-stage top-K: per class over H*W, then over the C*K survivors."""
    batch, cat, height, width = scores.size()
    topk_scores, topk_inds = torch.topk(scores.view(batch, cat, -1), K)
    topk_inds = topk_inds % (height * width)
    topk_ys = torch.div(topk_inds, width, rounding_mode="floor").int().float()
    topk_xs = (topk_inds % width).int().float()
    topk_score, topk_ind = torch.topk(topk_scores.view(batch, -1), K)
    topk_clses = torch.div(topk_ind, K, rounding_mode="floor").int()
    topk_inds = gather_feat(topk_inds.view(batch, -1, 1), topk_ind).view(batch, K)
    topk_ys = gather_feat(topk_ys.view(batch, -1, 1), topk_ind).view(batch, K)
    topk_xs = gather_feat(topk_xs.view(batch, -1, 1), topk_ind).view(batch, K)
    return topk_score, topk_inds, topk_clses, topk_ys, topk_xs


def decode_gather(heat: torch.Tensor, heads, K: int):
    """heat [B,C,H,W] (already sigmoid-ed, as the reference's decode expects) and a list of head
    maps [B,Ch,H,W] -> (scores[B,K], inds[B,K] int64, clses[B,K] int32, ys[B,K], xs[B,K],
    [gathered_h [B,K,Ch] ...]) exactly as `multi_pose_decode` obtains them
    (reference src/lib/models/decode.py:77-88)."""
    heat = nms(heat)
    scores, inds, clses, ys, xs = topk(heat, K)
    return scores, inds, clses, ys, xs, [transpose_and_gather_feat(h, inds) for h in heads]


def check_equivalent(got, heat, heads, K: int):
    """Is `got` (a decode_gather result) what the reference computes on (heat, heads), up to the ORDER OF
    EQUAL SCORES?  `torch.topk` leaves the order (and, at the cut, the choice) of tied scores unspecified
    (it differs between torch's CPU and CUDA kernels), so on head maps with ties an index-for-index
    comparison is not defined; this checks everything that is:
      1. the K scores equal the reference's, bit for bit, in order;
      2. every returned index is distinct within its image and points at a pixel whose NMS-ed score is
         exactly the returned score (so the selected SET is a valid top-K);
      3. ys / xs / clses are the coordinates of that index;
      4. each gathered vector is exactly the head's channel vector at that pixel.
    Without ties 1-4 imply equality with the reference index for index.  Returns (ok, n_images_with_ties).
    """
    scores, inds, clses, ys, xs, gathered = got
    B, C, H, W = heat.shape
    hn = nms(heat)
    ref_scores = topk(hn, K)[0]
    ok = torch.equal(scores, ref_scores)
    b_idx = torch.arange(B).view(B, 1).expand(B, K)
    ok &= bool((hn[b_idx, clses.long(), ys.long(), xs.long()] == scores).all())
    ok &= bool((ys.long() * W + xs.long() == inds).all())
    for b in range(B):
        ok &= len(set((clses[b].long() * H * W + inds[b]).tolist())) == K
    for h, g in zip(heads, gathered):
        ok &= torch.equal(h[b_idx, :, ys.long(), xs.long()], g)
    top = torch.topk(hn.view(B, -1), min(K + 1, C * H * W)).values
    ties = int((top[:, 1:] == top[:, :-1]).any(dim=1).sum())
    return bool(ok), ties
