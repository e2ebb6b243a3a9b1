"""Batch sharding of the SMPL forward across the GPUs of one box.

Bodies are independent, so rank r of R owns the contiguous rows
``[r*ceil(N/R), min(N, (r+1)*ceil(N/R)))`` of betas/pose/cam -- the explicit form of the batch
split the reference gets implicitly from ``nn.DataParallel`` (reference
src/lib/trains/trainer.py:176; intended uneven ``chunk_sizes`` at src/lib/opts.py:198-207).
The forward needs NO collective.  The only optional exchange is an all-gather of the small
per-body outputs (joints 288 B + kp2d 192 B per body) over NCCL/NVLink; vertices (82,680 B per
body) always stay on the rank that produced them.

Everything here is backend-agnostic ``torch.distributed`` so the host logic is testable with
``gloo`` on CPU (tests/test_sharding.py); on the GPU box the process group is NCCL.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_size(n: int, world_size: int) -> int:
    return (int(n) + world_size - 1) // world_size


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous shard [lo, hi) of rank `rank`; trailing ranks may be empty."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    per = shard_size(n, world_size)
    lo = min(int(n), rank * per)
    hi = min(int(n), lo + per)
    return lo, hi


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather row shards produced by `shard_bounds` back into [n_total, ...] on every rank.

    Shards are padded to the common ceil(N/R) rows so one fixed-size
    ``all_gather_into_tensor`` moves everything (a single NCCL launch), then trimmed.
    """
    world = dist.get_world_size(group)
    per = shard_size(n_total, world)
    tail = tuple(local.shape[1:])
    if local.shape[0] > per:
        raise ValueError("local shard larger than ceil(N/R)")
    send = local
    if local.shape[0] != per:
        send = local.new_zeros((per,) + tail)
        send[: local.shape[0]] = local
    out = local.new_empty((per * world,) + tail)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    return out[:n_total]


class ShardedSMPL:
    """Runs ``forward_fn`` (e.g. an ``SMPL`` module) on this rank's shard of a global batch.

    ``forward(betas, pose, cam, gather=True)`` takes the GLOBAL [N, .] parameter arrays (every
    rank holds them, as after the reference's decode stage) and returns
    ``(local_vertices, joints, kp2d, (lo, hi))`` where joints/kp2d are global when ``gather``
    is set and local otherwise.  The gather is issued on a side stream right after the
    forward so callers can keep consuming the local vertices meanwhile.
    """

    def __init__(self, forward_fn, group=None):
        self.forward_fn = forward_fn
        self.group = group

    def forward(self, betas, pose, cam=None, gather: bool = True):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = int(betas.shape[0])
        lo, hi = shard_bounds(n, world, rank)
        c = None if cam is None else cam[lo:hi]
        out = self.forward_fn(betas[lo:hi], pose[lo:hi], c)
        verts, joints = out[0], out[1]
        kp2d = out[2] if len(out) > 2 else None
        if gather and world > 1:
            joints = all_gather_rows(joints, n, self.group)
            if kp2d is not None:
                kp2d = all_gather_rows(kp2d, n, self.group)
        return verts, joints, kp2d, (lo, hi)
