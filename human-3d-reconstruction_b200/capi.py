"""ctypes binding of include/smpl_b200.h (libsmpl_b200.so).

This is the thin shim of SURVEY.md §8(b): validate, pass raw pointers + the current CUDA stream,
turn a non-zero status into ``RuntimeError``.  It mirrors what the reference's pybind module does
for its one native op (reference src/lib/models/DCNv2/src/vision.cpp:4-9 and the Python import at
src/lib/models/DCNv2/dcn_v2.py:13) but binds a plain C ABI so no torch headers are involved.

There is no CPU fallback: if the shared library is missing the import of this module still
succeeds (so CPU-only tooling can inspect the package) but every call raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SMPLB200_LIB: A/B timing of two builds of the same source inside one GPU session (scripts/ab_fused.sh)
LIB_PATH = os.environ.get("SMPLB200_LIB") or os.path.join(_HERE, "libsmpl_b200.so")

# ---- constants mirrored from smpl_b200.h -------------------------------------------------------
OK = 0
PREC_AUTO, PREC_FP32, PREC_BF16, PREC_TF32, PREC_BF16X3, PREC_F16, PREC_F16X3 = 0, 1, 2, 3, 4, 5, 6
PREC_MASK = 0x7
JOINTS_KINEMATIC, JOINTS_REGRESSED = 0, 1 << 3
ROTATE_BASE = 1 << 4
LBS_AUTO, LBS_FMA, LBS_TC, LBS_DENSE = 0, 1 << 5, 2 << 5, 3 << 5
TC_MIN_BATCH = 32        # AUTO: tcgen05 (f16x3) blendshapes from this many bodies
TC_LBS_MIN_BATCH = 384   # AUTO: tcgen05 skinning blend from this many bodies
DCN_INPUT_NHWC = 1       # smplb200_dcn_v2_forward flag: the input tensor is already channels-last
COEF_K = 224

PRECISIONS = {"auto": PREC_AUTO, "fp32": PREC_FP32, "bf16": PREC_BF16, "tf32": PREC_TF32,
              "bf16x3": PREC_BF16X3, "f16": PREC_F16, "f16x3": PREC_F16X3}
LBS_PATHS = {"auto": LBS_AUTO, "fma": LBS_FMA, "tc": LBS_TC, "dense": LBS_DENSE}


class ModelDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("num_verts", C.c_int32),
        ("num_joints", C.c_int32),
        ("num_betas", C.c_int32),
        ("v_template", C.c_void_p),
        ("shapedirs", C.c_void_p),
        ("posedirs", C.c_void_p),
        ("j_regressor", C.c_void_p),
        ("weights", C.c_void_p),
        ("parents", C.c_void_p),
    ]


class ForwardOpts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("joints_ready_event", C.c_void_p)]


def forward_opts(joints_ready_event=None):
    """SmplB200ForwardOpts for the `_opts` entry points; `joints_ready_event`: a recorded torch.cuda.Event."""
    if joints_ready_event is None:
        return None
    handle = joints_ready_event.cuda_event
    if not handle:
        raise RuntimeError("joints_ready event has no CUDA handle yet: call event.record() once before passing it")
    return C.pointer(ForwardOpts(C.sizeof(ForwardOpts), handle))      # pointer() keeps the struct alive


# every symbol include/smpl_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _u32, _sz, _int = C.c_void_p, C.c_int64, C.c_uint32, C.c_size_t, C.c_int
SYMBOLS = {
    "smplb200_model_create": (_int, [C.POINTER(ModelDesc), C.POINTER(_vp)]),
    "smplb200_model_destroy": (None, [_vp]),
    "smplb200_model_num_verts": (C.c_int32, [_vp]),
    "smplb200_model_num_joints": (C.c_int32, [_vp]),
    "smplb200_model_num_betas": (C.c_int32, [_vp]),
    "smplb200_model_device": (C.c_int32, [_vp]),
    "smplb200_model_max_weight_nnz": (C.c_int32, [_vp]),
    "smplb200_model_device_bytes": (_sz, [_vp]),
    "smplb200_workspace_bytes": (_sz, [_vp, _i64, _u32]),
    "smplb200_forward": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "smplb200_forward_opts": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _u32, _vp, _vp]),
    "smplb200_forward_host_opts": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _u32, _vp, _vp]),
    "smplb200_host_staging_layout": (_int, [_vp, _i64, _u32, C.POINTER(_sz), C.POINTER(_sz)]),
    "smplb200_push_rows": (_int, [C.c_int32, _vp, _vp, _i64, _i64, _vp, _vp, C.c_int32, C.c_int32, _u32, _vp, _vp]),
    "smplb200_wait_rows": (_int, [C.c_int32, _vp, C.c_int32, _u32, _vp]),
    "smplb200_exchange_rows_dma": (_int, [C.c_int32, _vp, _vp, _i64, _i64, _i64, _vp, _vp, C.c_int32, C.c_int32, _u32, _vp]),
    "smplb200_probe_fp32_fma": (_int, [C.c_int32, C.c_int32, _vp, C.POINTER(C.c_double), _vp]),
    "smplb200_host_staging_bytes": (_sz, [_vp, _i64, _u32]),
    "smplb200_forward_host": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "smplb200_padded_verts": (_i64, [_vp]),
    "smplb200_pose_chain": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _u32, _vp]),
    "smplb200_blendshapes_workspace_bytes": (_sz, [_vp, _i64, _u32]),
    "smplb200_blendshapes": (_int, [_vp, _vp, _i64, _vp, _vp, _sz, _u32, _vp]),
    "smplb200_lbs_workspace_bytes": (_sz, [_vp, _i64, _u32]),
    "smplb200_lbs": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _u32, _vp]),
    "smplb200_regress_joints": (_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "smplb200_blend_skin_workspace_bytes": (_sz, [_vp, _i64]),
    "smplb200_blend_skin": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "smplb200_backward_workspace_bytes": (_sz, [_vp, _i64, _u32, _int]),
    "smplb200_backward": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz,
                                 _u32, _vp]),
    "smplb200_backward_launch_count": (_int, [_vp, _i64, _u32, _int, _int]),
    "smplb200_decode_gather": (_int, [C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp,
                                      C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "smplb200_dcn_v2_workspace_bytes": (_sz, [C.c_int32] * 5 + [_u32]),
    "smplb200_dcn_v2_forward": (_int, [C.c_int32, _vp, _vp, _vp, _vp, _vp] + [C.c_int32] * 14 + [_vp, _vp, _sz, _u32, _vp]),
    "smplb200_strerror": (C.c_char_p, [_int]),
    "smplb200_version": (_int, []),
    "smplb200_last_cuda_error": (_int, []),
    "smplb200_forward_launch_count": (_int, [_vp, _i64, _u32, _int]),
}

_lib = None
_lib_lock = threading.Lock()


def lib():
    """Load libsmpl_b200.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                    "g.build()'` (nvcc, sm_100a). There is no CPU fallback for the SMPL kernels.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def strerror(status: int) -> str:
    return lib().smplb200_strerror(int(status)).decode()


def check(status: int, what: str) -> None:
    if status != OK:
        l = lib()
        extra = ""
        if status == 5:
            extra = f" [cudaError {l.smplb200_last_cuda_error()}]"
        raise RuntimeError(f"{what} failed: {strerror(status)} (status {status}){extra}")


def make_flags(precision="auto", joints="kinematic", rotate_base=False, lbs="auto") -> int:
    try:
        f = PRECISIONS[precision] | LBS_PATHS[lbs]
    except KeyError as e:
        raise ValueError(f"unknown precision/lbs option {e}") from None
    if joints == "regressed":
        f |= JOINTS_REGRESSED
    elif joints != "kinematic":
        raise ValueError("joints must be 'kinematic' or 'regressed'")
    if rotate_base:
        f |= ROTATE_BASE
    return f


class ModelHandle:
    """Owns one SmplB200Model* (one per device).  Host arrays are only read during create."""

    def __init__(self, model: dict, device: int):
        import numpy as np

        def f32(key):
            a = np.ascontiguousarray(np.asarray(model[key], dtype=np.float32))
            return a

        vt, sd, pd = f32("v_template"), f32("shapedirs"), f32("posedirs")
        jr, w = f32("J_regressor"), f32("weights")
        parents = np.ascontiguousarray(np.asarray(model["parents"]).astype(np.int64).astype(np.int32))
        V, NB, J = vt.shape[0], sd.shape[0], w.shape[1]
        if vt.shape != (V, 3) or sd.shape != (NB, 3 * V) or pd.shape != (9 * (J - 1), 3 * V) \
                or jr.shape != (V, J) or w.shape != (V, J) or parents.shape != (J,):
            raise ValueError("model tensors have inconsistent shapes (see SURVEY.md App. A.1)")
        desc = ModelDesc(C.sizeof(ModelDesc), int(device), V, J, NB,
                         vt.ctypes.data, sd.ctypes.data, pd.ctypes.data, jr.ctypes.data,
                         w.ctypes.data, parents.ctypes.data)
        out = C.c_void_p()
        check(lib().smplb200_model_create(C.byref(desc), C.byref(out)), "smplb200_model_create")
        self.ptr = out
        self._sizes = {}
        self.device = int(device)
        self.num_verts, self.num_joints, self.num_betas = V, J, NB
        self.padded_verts = int(lib().smplb200_padded_verts(out))
        self.max_weight_nnz = int(lib().smplb200_model_max_weight_nnz(out))
        self.device_bytes = int(lib().smplb200_model_device_bytes(out))

    def close(self):
        if getattr(self, "ptr", None):
            lib().smplb200_model_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def workspace_bytes(self, n: int, flags: int) -> int:
        key = (n, flags)
        v = self._sizes.get(key)
        if v is None:      # pure function of (n, flags): cached, the hot path makes no ctypes call for it
            v = self._sizes[key] = int(lib().smplb200_workspace_bytes(self.ptr, int(n), int(flags)))
        return v

    def backward_workspace_bytes(self, n: int, flags: int, vertex_path: bool) -> int:
        key = (n, flags, bool(vertex_path))
        v = self._sizes.get(key)
        if v is None:
            v = self._sizes[key] = int(lib().smplb200_backward_workspace_bytes(
                self.ptr, int(n), int(flags), int(bool(vertex_path))))
        return v

    def host_staging_bytes(self, n: int, flags: int) -> int:
        return int(lib().smplb200_host_staging_bytes(self.ptr, int(n), int(flags)))

    def launch_count(self, n: int, flags: int, with_projection: bool) -> int:
        return int(lib().smplb200_forward_launch_count(self.ptr, int(n), int(flags), int(with_projection)))
