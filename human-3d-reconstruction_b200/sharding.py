"""Batch sharding of the SMPL forward across the GPUs of one box.

Bodies are independent, so rank r of R owns the contiguous rows
``[r*ceil(N/R), min(N, (r+1)*ceil(N/R)))`` of betas/pose/cam -- the explicit form of the batch
split the reference gets implicitly from ``nn.DataParallel`` (reference
src/lib/trains/trainer.py:176; intended uneven ``chunk_sizes`` at src/lib/opts.py:198-207).
The forward needs NO collective.  The only optional exchange makes the small per-body outputs
(joints 288 B + kp2d 192 B per body) of every rank visible on every rank; vertices (82,680 B per
body) always stay on the rank that produced them.  Two transports:

  * `PeerExchange` (the B200 path): each rank PUSHES its rows into every peer's copy of a
    symmetric buffer over NVLink / NVSwitch and raises a per-rank flag; no NCCL launch anywhere.
    Default ('peer', what 'auto' picks): plain peer stores from one small kernel
    (`smplb200_push_rows`, csrc/k_exchange.cuh).  Those stores share every SM's store path with the
    compute kernels and cost the step about the NVLink transfer time (~1.2 us per MB pushed, measured).
    Opt-in ('dma'): copy-engine peer copies + stream memory operations, no kernel at all
    (`smplb200_exchange_rows_dma`): 2 GPUs 58.5 M bodies/s against 57.3 M ('peer') and 58.9 M (no exchange)
    on one box, correct at 2 and 4 GPUs -- but its only 8-GPU run did not finish (GPU budget ran out
    before it could be debugged), so it is NOT the default.
    Either way it runs on a side stream that waits only for the forward's "joints ready" event
    (recorded right after the ~10 us chain kernel), so it overlaps the blendshape / skinning kernels.
  * `all_gather_rows`: one fixed-size ``all_gather_into_tensor`` -- what `PeerExchange` falls back
    to (still on its side stream) when peer mapping is unavailable, and what the ``gloo`` tests
    exercise on CPU.

The host logic is backend-agnostic ``torch.distributed`` (tests/test_sharding.py, world size 2,
gloo); on the GPU box the process group is NCCL and is used for rendezvous / barriers only.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

import os
_NOWAIT = os.environ.get("SMPLB200_XCHG_NOWAIT") == "1"      # measurement knob: push only, nobody waits for the peers
ROW = 120   # floats per body in the gathered buffer: joints 72 | kp2d 48


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


class PeerExchange:
    """All ranks' joints | kp2d rows on every rank, off the compute stream.

    ``exchange(joints, kp2d, ready)`` is called right after the forward was enqueued with
    ``joints_ready=ready``; it returns ``(joints_all[N,24,3], kp2d_all[N,24,2], done_event)``.  The
    returned tensors are views of the current slot of the gathered buffer and are complete once
    ``done_event`` has fired (``stream.wait_event(done_event)`` before reading them).  Two slots
    alternate: a view stays valid until the next-but-one exchange, provided its consumer runs in
    stream order before the next ``exchange`` call is enqueued (the usual pipeline).

    transport 'peer' ('auto' picks it): symmetric memory (``torch.distributed._symmetric_memory``: one
    rendezvous at construction maps every rank's buffer into this process) + the library's peer-store kernels.
    transport 'dma' (opt-in, validated at 2 and 4 GPUs only): the same mapping + copy-engine peer copies and
    stream memory operations (``smplb200_exchange_rows_dma``): nothing runs on the SMs.
    transport 'collective': ``all_gather_into_tensor`` on the side stream (fallback; CPU/gloo tests).
    """

    def __init__(self, n_total: int, device, group=None, transport: str = "auto", slots: int = 2):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n_total, self.per = int(n_total), shard_size(n_total, self.world)
        self.device = torch.device(device)
        self.slots, self.epoch = int(slots), 0
        self.rows = self.per * self.world
        self.transport = "collective"
        self.why_not_peer = None
        cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=self.device) if cuda else None
        if transport in ("auto", "peer", "dma") and cuda and self.world > 1:
            try:
                self._setup_peer()
                self.transport = "peer"
                if transport == "dma":
                    # probe the copy-engine path once (epoch 0: writes 0 into flags that are 0, waits for >= 0)
                    st = self._lib.smplb200_exchange_rows_dma(
                        self._dev_index, None, None, 0, 0, self.rows, self._peer_slots[0], self._peer_flags,
                        self.world, self.rank, 0, self.stream.cuda_stream)
                    self._check(st, "smplb200_exchange_rows_dma")
                    self.transport = "dma"
            except Exception as e:      # no IPC between these processes / unsupported build: say so, fall back
                if transport in ("peer", "dma"):
                    raise
                self.why_not_peer = f"{type(e).__name__}: {e}"
        if self.transport == "collective":
            self.buf = torch.zeros((self.slots, self.rows, ROW), dtype=torch.float32, device=self.device)
        self._done = [torch.cuda.Event() for _ in range(self.slots)] if cuda else None

    def _setup_peer(self):
        import torch.distributed._symmetric_memory as symm_mem
        from . import capi
        slot_bytes = self.rows * ROW * 4
        self._flags_off = (self.slots * slot_bytes + 255) // 256 * 256
        total = self._flags_off + 512                      # uint32 flags[16] | uint32 counter at +256
        t = symm_mem.empty(total // 4, dtype=torch.float32, device=self.device)
        t.zero_()
        hdl = symm_mem.rendezvous(t, self.group.group_name)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                           # every rank's flags are zero before anyone pushes
        self._symm, self._hdl = t, hdl
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        self._peer_slots = [(C.c_void_p * self.world)(*[p + s * slot_bytes for p in ptrs]) for s in range(self.slots)]
        self._peer_flags = (C.c_void_p * self.world)(*[p + self._flags_off for p in ptrs])
        self._my_flags = ptrs[self.rank] + self._flags_off
        self._counter = ptrs[self.rank] + self._flags_off + 256
        self.buf = t[: self.slots * self.rows * ROW].view(self.slots, self.rows, ROW)
        self._lib = capi.lib()
        self._check = capi.check
        self._push, self._wait = self._lib.smplb200_push_rows, self._lib.smplb200_wait_rows
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()

    def exchange(self, joints: torch.Tensor, kp2d, ready=None):
        self.epoch += 1
        slot = self.epoch % self.slots
        n = int(joints.shape[0])
        lo = self.rank * self.per
        out = self.buf[slot]
        done = None
        if self.stream is not None:
            if ready is not None:
                self.stream.wait_event(ready)
            else:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self.transport == "dma":           # copy-engine copies + stream memory ops: no kernel, one C call
            st = self._lib.smplb200_exchange_rows_dma(
                self._dev_index, joints.data_ptr(), None if kp2d is None else kp2d.data_ptr(), n, lo, self.rows,
                self._peer_slots[slot], self._peer_flags, self.world, self.rank, self.epoch & 0xFFFFFFFF,
                self.stream.cuda_stream)
            if st:
                self._check(st, "smplb200_exchange_rows_dma")
            done = self._done[slot]
            done.record(self.stream)
            flat = self.buf[slot].view(-1)        # slot layout: joints block [rows][72] | kp2d block [rows][48]
            j_all = flat[: self.rows * 72].view(self.rows, 24, 3)[: self.n_total]
            k_all = flat[self.rows * 72:].view(self.rows, 24, 2)[: self.n_total]
            return j_all, k_all, done
        if self.transport == "peer":          # two C calls on the side stream; no torch stream context needed
            idx, s, ep = self._dev_index, self.stream.cuda_stream, self.epoch & 0xFFFFFFFF
            st = self._push(idx, joints.data_ptr(), None if kp2d is None else kp2d.data_ptr(), n, lo,
                            self._peer_slots[slot], self._peer_flags, self.world, self.rank, ep, self._counter, s)
            if not _NOWAIT:
                st = st or self._wait(idx, self._my_flags, self.world, ep, s)
            if st:
                self._check(st, "smplb200_push_rows / smplb200_wait_rows")
            done = self._done[slot]
            done.record(self.stream)
            rows = out[: self.n_total] if self.per * self.world != self.n_total else out
            return rows[:, :72].unflatten(1, (24, 3)), rows[:, 72:].unflatten(1, (24, 2)), done
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            send = joints.new_zeros((self.per, ROW))
            send[:n, :72] = joints.reshape(n, 72)
            if kp2d is not None:
                send[:n, 72:] = kp2d.reshape(n, 48)
            if self.world > 1:
                dist.all_gather_into_tensor(out.view(-1), send.view(-1), group=self.group)
            else:
                out.copy_(send)
            if self.stream is not None:
                done = torch.cuda.Event()
                done.record(self.stream)
        rows = out[: self.n_total] if self.per * self.world != self.n_total else out
        return rows[:, :72].unflatten(1, (24, 3)), rows[:, 72:].unflatten(1, (24, 2)), done


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ShardedSMPL:
    """Runs ``forward_fn`` (e.g. an ``SMPL`` module) on this rank's shard of a global batch.

    ``forward(betas, pose, cam, gather=True)`` takes the GLOBAL [N, .] parameter arrays (every
    rank holds them, as after the reference's decode stage) and returns
    ``(local_vertices, joints, kp2d, (lo, hi))`` where joints/kp2d are global when ``gather``
    is set and local otherwise.  With ``exchange=PeerExchange(...)`` the gather runs on the
    exchange's side stream as soon as the joints exist and the current stream waits for it before
    returning the views; without one it is a plain in-line ``all_gather_rows``.
    """

    def __init__(self, forward_fn, group=None, exchange: "PeerExchange | None" = None):
        self.forward_fn = forward_fn
        self.group = group
        self.exchange = exchange

    def forward(self, betas, pose, cam=None, gather: bool = True):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = int(betas.shape[0])
        lo, hi = shard_bounds(n, world, rank)
        c = None if cam is None else cam[lo:hi]
        ex = self.exchange if (gather and world > 1) else None
        ready = None
        if ex is not None and ex.stream is not None:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(ex.device))      # creates the CUDA handle
            out = self.forward_fn(betas[lo:hi], pose[lo:hi], c, joints_ready=ready)
        else:
            out = self.forward_fn(betas[lo:hi], pose[lo:hi], c)
        verts, joints = out[0], out[1]
        kp2d = out[2] if len(out) > 2 else None
        if ex is not None:
            joints, kp2d_all, done = ex.exchange(joints, kp2d, ready)
            kp2d = kp2d_all if kp2d is not None else None
            if done is not None:
                torch.cuda.current_stream(ex.device).wait_event(done)
        elif gather and world > 1:
            joints = all_gather_rows(joints, n, self.group)
            if kp2d is not None:
                kp2d = all_gather_rows(kp2d, n, self.group)
        return verts, joints, kp2d, (lo, hi)
