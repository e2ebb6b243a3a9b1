"""Drop-in SMPL layer: ``forward(betas, pose, cam=None) -> (vertices, joints[, kp2d])``.

The reference snapshot names an HMR trainer but ships no SMPL module; this class
is the layer BASELINE.json's north_star asks for, shaped like the reference's one native-op
module (nn.Module over an autograd-free native call, reference
src/lib/models/DCNv2/dcn_v2.py:57-128) and fed the way its decode stage would feed it: per-person
vectors gathered by ``_transpose_and_gather_feat`` (reference src/lib/models/utils.py:23-27) from
heads ``{'pose': 72, 'shape': 10, 'cam': 3}`` (reference src/lib/opts.py:248-258).

Model tensors are registered buffers so ``nn.DataParallel`` (reference
src/lib/trains/trainer.py:176) replicates them; the packed device-side model (one
``SmplB200Model*`` per device) is created lazily from the buffers on first use.

Differentiable: when an input requires grad the call goes through ``_SMPLFunction`` whose
backward is ``smplb200_backward`` (hand-written kernels, csrc/k_backward.cuh) -- what the reference
trainer needs to put a loss on the layer's outputs (reference src/lib/trains/trainer.py:31-37,
102-104).  Outputs that do not reach the loss cost nothing: without a vertex gradient the backward
is one small kernel.  CUDA-only: a CPU tensor raises -- there is no CPU path in the product.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch
import torch.nn as nn

from . import capi, model_io, synthetic

_BUFFERS = ("v_template", "shapedirs", "posedirs", "J_regressor", "weights")


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _check_in(name, t, n, width, device):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.device != device:
        raise RuntimeError(f"{name} is on {t.device}, expected {device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 2 or t.shape[0] != n or t.shape[1] != width:
        raise ValueError(f"{name} must have shape [{n}, {width}], got {tuple(t.shape)}")
    return t.contiguous()


class _SMPLFunction(torch.autograd.Function):
    """autograd node: forward = smplb200_forward, backward = smplb200_backward (first order only)."""

    @staticmethod
    def forward(ctx, layer, flags, betas, pose, cam):
        keep = layer.save_forward_workspace
        outs = layer._forward_impl(betas, pose, cam, flags, return_workspace=keep)
        ctx.fwd_ws = None
        if keep:
            outs, ws = outs[:-1], outs[-1]
            # kept alive for the backward (A and vposed are reused instead of recomputed) unless big
            ctx.fwd_ws = ws if (ws is not None and ws.numel() <= layer.save_forward_workspace_max_bytes) else None
        ctx.layer, ctx.flags, ctx.has_cam = layer, flags, cam is not None
        ctx.set_materialize_grads(False)      # an output that does not reach the loss stays None
        ctx.save_for_backward(betas, pose, cam, outs[1])
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_verts, g_joints, g_kp2d=None):
        betas, pose, cam, joints = ctx.saved_tensors
        layer, flags = ctx.layer, ctx.flags
        device, n = betas.device, int(betas.shape[0])
        h = layer.handle(device)

        def prep(g):
            return None if g is None else g.to(torch.float32).contiguous()

        g_verts, g_joints, g_kp2d = prep(g_verts), prep(g_joints), prep(g_kp2d)
        regressed = bool(flags & capi.JOINTS_REGRESSED)
        vertex_path = g_verts is not None or (regressed and (g_joints is not None or g_kp2d is not None))
        # (no torch.cuda.device() switch: every allocation names its device and the library runs on
        #  the handle's device whatever the caller's current one is)
        g_betas = torch.empty_like(betas)
        g_pose = torch.empty_like(pose)
        g_cam = torch.empty_like(cam) if cam is not None else None
        if n > 0:
            wsb = h.backward_workspace_bytes(n, flags, vertex_path)
            if wsb == 0:
                raise RuntimeError("smplb200_backward_workspace_bytes rejected the flag combination")
            ws = torch.empty(wsb, dtype=torch.uint8, device=device) if vertex_path else None
            fws = ctx.fwd_ws
            capi.check(capi.lib().smplb200_backward(
                h.ptr, _ptr(betas), _ptr(pose), _ptr(cam), n, _ptr(joints),
                _ptr(g_verts), _ptr(g_joints), _ptr(g_kp2d),
                _ptr(g_betas), _ptr(g_pose), _ptr(g_cam),
                _ptr(fws), 0 if fws is None else fws.numel(),
                _ptr(ws), wsb if vertex_path else 0, flags, _stream_ptr(device)), "smplb200_backward")
        ctx.fwd_ws = None
        need = ctx.needs_input_grad   # (layer, flags, betas, pose, cam)
        return (None, None, g_betas if need[2] else None, g_pose if need[3] else None,
                g_cam if (ctx.has_cam and need[4]) else None)


class SMPL(nn.Module):
    """SMPL body model on B200.

    Args:
      model: dict, or path to an ``.npz``/pickle, with ``v_template[V,3]``, ``shapedirs[NB,3V]``,
             ``posedirs[207,3V]``, ``J_regressor[V,24]``, ``weights[V,24]``, ``parents[24]`` in the
             eager layer's layouts -- or the official SMPL file layout
             (``shapedirs[V,3,NB]``, ``J_regressor[24,V]`` dense/sparse, ``kintree_table``), see model_io.
      precision: 'auto' | 'fp32' | 'bf16' | 'tf32' | 'bf16x3' | 'f16x3' | 'f16'  (blendshape operands;
             'auto' = fp32 FMA below 32 bodies, 'f16x3' (split fp16, bound 4e-6 m) from there; 'f16' =
             the fused blendshapes+skinning kernel, stated vertex bound 5e-5 m)
      joints: 'kinematic' (J_posed of the chain, default) | 'regressed' (HMR-style, from vertices)
      rotate_base: HMR's root pre-rotation by diag(1,-1,-1); default False
      lbs: 'auto' | 'fma' | 'tc' | 'dense'  (skinning kernel)
    """

    def __init__(self, model, precision="auto", joints="kinematic", rotate_base=False, lbs="auto",
                 save_forward_workspace=True, save_forward_workspace_max_bytes=1 << 30):
        super().__init__()
        # training: keep the forward's scratch (A, vposed: ~0.2 MB/body) for the backward instead of
        # recomputing it, up to this many bytes per call
        self.save_forward_workspace = bool(save_forward_workspace)
        self.save_forward_workspace_max_bytes = int(save_forward_workspace_max_bytes)
        if isinstance(model, (str, bytes)):      # .npz / pickle, official SMPL layout or the eager layer's
            model = model_io.load_model(model)
        elif "kintree_table" in model or np.asarray(model["shapedirs"]).ndim == 3:
            model = model_io.from_official_layout(model)
        for k in _BUFFERS:
            self.register_buffer(k, torch.as_tensor(np.asarray(model[k], dtype=np.float32)).clone())
        self.register_buffer(
            "parents", torch.as_tensor(np.asarray(model["parents"]).astype(np.int64)).clone())
        self.num_verts = int(self.v_template.shape[0])
        self.num_joints = int(self.weights.shape[1])
        self.num_betas = int(self.shapedirs.shape[0])
        self.flags = capi.make_flags(precision, joints, rotate_base, lbs)
        self.precision, self.joints_from, self.rotate_base, self.lbs = precision, joints, rotate_base, lbs
        # shared by DataParallel replicas (replicate() shallow-copies __dict__): device index -> handle
        self._handles = {}
        self._handles_lock = threading.Lock()

    # -- packed-model cache hygiene --------------------------------------------------------------
    # The per-device SmplB200Model is packed ONCE from the registered buffers.  Anything that can
    # change the buffers drops the cache so the next forward re-packs: load_state_dict(), .to() /
    # .float() / .cuda() (`_apply`), and `invalidate()` for callers that edit a buffer in place.
    def invalidate(self):
        """Drop the packed per-device models; the next forward re-packs them from the buffers."""
        with self._handles_lock:
            self._handles.clear()

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def __getstate__(self):
        # copy.deepcopy / torch.save(module): device handles and the lock are per-process state
        state = self.__dict__.copy()
        state.pop("_handles", None)
        state.pop("_handles_lock", None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._handles = {}
        self._handles_lock = threading.Lock()

    # -- construction helpers ------------------------------------------------------------------
    @classmethod
    def synthetic(cls, seed: int = 0, weights: str = "sparse", regressor: str = "sparse", **kw):
        """Seeded SMPL-shaped random model (the licensed model file is not available offline)."""
        return cls(synthetic.make_model(seed, weights=weights, regressor=regressor), **kw)

    def model_dict(self) -> dict:
        d = {k: getattr(self, k).detach().cpu().numpy() for k in _BUFFERS}
        d["parents"] = self.parents.detach().cpu().numpy().astype(np.int32)
        return d

    def handle(self, device) -> capi.ModelHandle:
        if not isinstance(device, torch.device):
            device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("SMPL (B200) runs on CUDA devices only; there is no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            with self._handles_lock:
                h = self._handles.get(idx)
                if h is None:
                    h = capi.ModelHandle(self.model_dict(), idx)
                    self._handles[idx] = h
        return h

    # -- the forward pass ------------------------------------------------------------------------
    def forward(self, betas, pose, cam=None, *, return_kp2d=None, flags=None, joints_ready=None):
        """betas[N,NB], pose[N,72] (axis-angle), cam[N,3]=(s,tx,ty) or None.

        Returns (vertices[N,V,3], joints[N,24,3]) and, when ``cam`` is given, kp2d[N,24,2].
        ``return_kp2d`` (SURVEY.md §8b signature): None = follow ``cam``; True requires ``cam``;
        False drops the projection even when ``cam`` is passed.
        ``joints_ready``: an (already once-recorded) ``torch.cuda.Event`` that the library re-records on the
        current stream as soon as joints and kp2d are final -- right after the ~10 us chain kernel, long
        before the vertices -- so a side stream can start exchanging them (sharding.PeerExchange).
        """
        if return_kp2d and cam is None:
            raise ValueError("return_kp2d=True needs cam[N,3]")
        if return_kp2d is False:
            cam = None
        if not isinstance(betas, torch.Tensor) or betas.device.type != "cuda":
            raise RuntimeError("SMPL (B200) needs CUDA tensors; there is no CPU fallback")
        device = betas.device
        n = int(betas.shape[0])
        betas = _check_in("betas", betas, n, self.num_betas, device)
        pose = _check_in("pose", pose, n, 3 * self.num_joints, device)
        if cam is not None:
            cam = _check_in("cam", cam, n, 3, device)
        flags = self.flags if flags is None else int(flags)
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (betas, pose, cam)):
            if joints_ready is not None:
                raise ValueError("joints_ready is an inference-path option (no_grad)")
            return _SMPLFunction.apply(self, flags, betas, pose, cam)
        return self._forward_impl(betas, pose, cam, flags, joints_ready=joints_ready)

    def _forward_impl(self, betas, pose, cam, flags, return_workspace=False, joints_ready=None):
        device, n = betas.device, int(betas.shape[0])
        h = self.handle(device)
        verts = torch.empty((n, self.num_verts, 3), dtype=torch.float32, device=device)
        joints = torch.empty((n, self.num_joints, 3), dtype=torch.float32, device=device)
        kp2d = None if cam is None else torch.empty((n, self.num_joints, 2), dtype=torch.float32, device=device)
        ws = None
        if n > 0:
            ws_bytes = h.workspace_bytes(n, flags)
            if ws_bytes == 0:
                raise RuntimeError("smplb200_workspace_bytes rejected the flag combination")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
            capi.check(capi.lib().smplb200_forward_opts(
                h.ptr, _ptr(betas), _ptr(pose), _ptr(cam), n, _ptr(verts), _ptr(joints), _ptr(kp2d),
                _ptr(ws), ws_bytes, flags, _stream_ptr(device), capi.forward_opts(joints_ready)),
                "smplb200_forward")
        elif joints_ready is not None:
            joints_ready.record(torch.cuda.current_stream(device))
        outs = (verts, joints) if cam is None else (verts, joints, kp2d)
        return outs + (ws,) if return_workspace else outs

    def launch_count(self, n: int, with_projection: bool, device=None) -> int:
        h = self.handle(device if device is not None else self.v_template.device)
        return h.launch_count(n, self.flags, with_projection)


class GraphedSMPL:
    """CUDA-graph replay of the forward for a fixed batch size (small-batch / latency regime).

    At N <= ~256 the forward is launch-bound (three ~5-20 us kernels behind ~40 us of Python and
    launch overhead per call); capturing the k2 -> k1 -> k3 sequence once and replaying it removes
    the per-call host cost.  Inputs are static device tensors (``betas``, ``pose``, ``cam``): write
    new parameters into them (``copy_``), call ``replay()``, read ``vertices`` / ``joints`` / ``kp2d``.
    """

    def __init__(self, smpl: "SMPL", n: int, device, with_cam: bool = True):
        self.device = torch.device(device)
        dev = self.device
        self.betas = torch.zeros((n, smpl.num_betas), dtype=torch.float32, device=dev)
        self.pose = torch.zeros((n, 3 * smpl.num_joints), dtype=torch.float32, device=dev)
        self.cam = torch.zeros((n, 3), dtype=torch.float32, device=dev) if with_cam else None
        if self.cam is not None:
            self.cam[:, 0] = 1.0
        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):       # warm-up: creates the per-device handle, primes caches
                for _ in range(2):
                    smpl(self.betas, self.pose, self.cam)
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                out = smpl(self.betas, self.pose, self.cam)
        self.vertices, self.joints = out[0], out[1]
        self.kp2d = out[2] if len(out) > 2 else None

    def replay(self):
        self.graph.replay()
        return (self.vertices, self.joints) if self.kp2d is None else (self.vertices, self.joints, self.kp2d)


class StaticSMPL:
    """Inference runner with STATIC device buffers for a fixed batch size: inputs (``betas``, ``pose``, ``cam``),
    outputs and workspace are allocated once, so ``run()`` is a single C call (``smplb200_forward_opts``) --
    ~10 us of host time instead of the ~60 us the general ``SMPL.forward`` spends validating, allocating and
    marshalling.  A serving loop at 4096 bodies is otherwise HOST-bound (the device step is 0.14 ms).
    Write new parameters into the input tensors (``copy_``), call ``run()``, read the outputs in stream order.
    """

    def __init__(self, smpl: "SMPL", n: int, device, with_cam: bool = True):
        self.smpl, self.n = smpl, int(n)
        self.device = dev = torch.device(device)
        self.h = smpl.handle(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.betas = torch.zeros((n, smpl.num_betas), **f32)
        self.pose = torch.zeros((n, 3 * smpl.num_joints), **f32)
        self.cam = torch.zeros((n, 3), **f32) if with_cam else None
        self.vertices = torch.empty((n, smpl.num_verts, 3), **f32)
        self.joints = torch.empty((n, smpl.num_joints, 3), **f32)
        self.kp2d = torch.empty((n, smpl.num_joints, 2), **f32) if with_cam else None
        self.ws_bytes = self.h.workspace_bytes(n, smpl.flags)
        if self.ws_bytes == 0:
            raise RuntimeError("smplb200_workspace_bytes rejected the flag combination")
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self._fn = capi.lib().smplb200_forward_opts
        self._args = (self.h.ptr, _ptr(self.betas), _ptr(self.pose), _ptr(self.cam), self.n, _ptr(self.vertices),
                      _ptr(self.joints), _ptr(self.kp2d), _ptr(self.ws), self.ws_bytes, smpl.flags)
        self._opts = {}

    def run(self, stream=None, joints_ready=None):
        s = _stream_ptr(self.device) if stream is None else stream.cuda_stream
        opts = None
        if joints_ready is not None:
            key = id(joints_ready)
            opts = self._opts.get(key)
            if opts is None:          # the ctypes struct is built once per event object
                opts = self._opts[key] = (capi.forward_opts(joints_ready), joints_ready)
            opts = opts[0]
        st = self._fn(*self._args, s, opts)
        if st:
            capi.check(st, "smplb200_forward")
        return (self.vertices, self.joints) if self.kp2d is None else (self.vertices, self.joints, self.kp2d)


class HostRunner:
    """End-to-end runner over HOST buffers through ``smplb200_forward_host``.

    Owns pinned host input/output buffers and one device staging arena sized for ``n`` bodies;
    each ``run()`` copies the inputs host->device, runs the forward and copies the requested
    outputs device->host on the given stream (no synchronisation inside).
    """

    def __init__(self, smpl: SMPL, n: int, device, with_vertices: bool = False, with_cam: bool = True):
        self.smpl, self.n = smpl, int(n)
        self.device = torch.device(device)
        self.h = smpl.handle(self.device)
        pin = dict(dtype=torch.float32, pin_memory=True)
        self.betas = torch.empty((n, smpl.num_betas), **pin)
        self.pose = torch.empty((n, 3 * smpl.num_joints), **pin)
        self.cam = torch.empty((n, 3), **pin) if with_cam else None
        self.vertices = torch.empty((n, smpl.num_verts, 3), **pin) if with_vertices else None
        self.joints = torch.empty((n, smpl.num_joints, 3), **pin)
        self.kp2d = torch.empty((n, smpl.num_joints, 2), **pin) if with_cam else None
        self.staging_bytes = self.h.host_staging_bytes(n, smpl.flags)
        with torch.cuda.device(self.device):
            self.staging = torch.empty(self.staging_bytes, dtype=torch.uint8, device=self.device)
        # device copies of the small outputs inside the staging arena (what a multi-GPU caller exchanges)
        oj, ok = C.c_size_t(), C.c_size_t()
        capi.check(capi.lib().smplb200_host_staging_layout(self.h.ptr, self.n, smpl.flags, C.byref(oj), C.byref(ok)),
                   "smplb200_host_staging_layout")
        nj = smpl.num_joints
        self.joints_dev = self.staging[oj.value: oj.value + n * nj * 12].view(torch.float32).view(n, nj, 3)
        self.kp2d_dev = self.staging[ok.value: ok.value + n * nj * 8].view(torch.float32).view(n, nj, 2)

    @property
    def h2d_bytes(self) -> int:
        return sum(t.numel() * 4 for t in (self.betas, self.pose, self.cam) if t is not None)

    @property
    def d2h_bytes(self) -> int:
        return sum(t.numel() * 4 for t in (self.vertices, self.joints, self.kp2d) if t is not None)

    def run(self, stream=None, joints_ready=None):
        s = _stream_ptr(self.device) if stream is None else stream.cuda_stream
        with torch.cuda.device(self.device):
            capi.check(capi.lib().smplb200_forward_host_opts(
                self.h.ptr, _ptr(self.betas), _ptr(self.pose), _ptr(self.cam), self.n,
                _ptr(self.vertices), _ptr(self.joints), _ptr(self.kp2d),
                _ptr(self.staging), self.staging_bytes, self.smpl.flags, s, capi.forward_opts(joints_ready)),
                "smplb200_forward_host")


# ---- per-kernel entry points (unit parity, ncu) --------------------------------------------------
def pose_chain(smpl: SMPL, betas, pose, flags=None):
    """k2 -> (coef[N,224], A[N,24,12], joints[N,24,3])."""
    device, n = betas.device, int(betas.shape[0])
    h = smpl.handle(device)
    flags = smpl.flags if flags is None else flags
    with torch.cuda.device(device):
        coef = torch.empty((n, capi.COEF_K), dtype=torch.float32, device=device)
        A = torch.empty((n, smpl.num_joints, 12), dtype=torch.float32, device=device)
        joints = torch.empty((n, smpl.num_joints, 3), dtype=torch.float32, device=device)
        capi.check(capi.lib().smplb200_pose_chain(
            h.ptr, _ptr(betas.contiguous()), _ptr(pose.contiguous()), n, _ptr(coef), _ptr(A), _ptr(joints),
            flags, _stream_ptr(device)), "smplb200_pose_chain")
    return coef, A, joints


def blendshapes(smpl: SMPL, coef, flags=None):
    """k1 -> vposed planar [N,3,VP]."""
    device, n = coef.device, int(coef.shape[0])
    h = smpl.handle(device)
    flags = smpl.flags if flags is None else flags
    with torch.cuda.device(device):
        vposed = torch.empty((n, 3, h.padded_verts), dtype=torch.float32, device=device)
        wsb = int(capi.lib().smplb200_blendshapes_workspace_bytes(h.ptr, n, flags))
        ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=device)
        capi.check(capi.lib().smplb200_blendshapes(
            h.ptr, _ptr(coef.contiguous()), n, _ptr(vposed), _ptr(ws), wsb, flags,
            _stream_ptr(device)), "smplb200_blendshapes")
    return vposed


def lbs(smpl: SMPL, vposed, A, joints=None, cam=None, flags=None):
    """k3 (+k4) -> vertices[N,V,3] (and kp2d when joints and cam are given)."""
    device, n = vposed.device, int(vposed.shape[0])
    h = smpl.handle(device)
    flags = smpl.flags if flags is None else flags
    with torch.cuda.device(device):
        verts = torch.empty((n, smpl.num_verts, 3), dtype=torch.float32, device=device)
        kp2d = None
        if cam is not None and joints is not None:
            kp2d = torch.empty((n, smpl.num_joints, 2), dtype=torch.float32, device=device)
        wsb = int(capi.lib().smplb200_lbs_workspace_bytes(h.ptr, n, flags))
        ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=device)
        capi.check(capi.lib().smplb200_lbs(
            h.ptr, _ptr(vposed.contiguous()), _ptr(A.contiguous()), n, _ptr(verts),
            _ptr(joints), _ptr(cam), _ptr(kp2d), _ptr(ws), wsb, flags, _stream_ptr(device)),
            "smplb200_lbs")
    return verts if kp2d is None else (verts, kp2d)


def blend_skin(smpl: SMPL, coef, A):
    """k1 + k3 fused (precision 'f16') -> vertices[N,V,3] from k2's coef[N,224] and A[N,24,12]."""
    device, n = coef.device, int(coef.shape[0])
    h = smpl.handle(device)
    with torch.cuda.device(device):
        verts = torch.empty((n, smpl.num_verts, 3), dtype=torch.float32, device=device)
        wsb = int(capi.lib().smplb200_blend_skin_workspace_bytes(h.ptr, n))
        if wsb == 0:
            raise RuntimeError("fused blendshapes+skinning kernel unavailable for this model (needs <= 13 betas)")
        ws = torch.empty(wsb, dtype=torch.uint8, device=device)
        capi.check(capi.lib().smplb200_blend_skin(
            h.ptr, _ptr(coef.contiguous()), _ptr(A.contiguous()), n, _ptr(verts), _ptr(ws), wsb,
            _stream_ptr(device)), "smplb200_blend_skin")
    return verts


def regress_joints(smpl: SMPL, vertices, cam=None):
    device, n = vertices.device, int(vertices.shape[0])
    h = smpl.handle(device)
    with torch.cuda.device(device):
        joints = torch.empty((n, smpl.num_joints, 3), dtype=torch.float32, device=device)
        kp2d = None if cam is None else torch.empty((n, smpl.num_joints, 2), dtype=torch.float32, device=device)
        capi.check(capi.lib().smplb200_regress_joints(
            h.ptr, _ptr(vertices.contiguous()), n, _ptr(joints), _ptr(cam), _ptr(kp2d),
            _stream_ptr(device)), "smplb200_regress_joints")
    return joints if kp2d is None else (joints, kp2d)
