"""B200-native SMPL body-model forward pass (drop-in for the reference's PyTorch SMPL layer).

Public surface:
    SMPL                 nn.Module, forward(betas, pose, cam=None, ...) -> (vertices, joints[, kp2d])
    capi                 ctypes binding of include/smpl_b200.h (libsmpl_b200.so)
    synthetic            seeded SMPL-shaped model tensors / parameters
    sharding             batch sharding across ranks (+ optional NCCL gather of joints)
    decode_gather        fused NMS + top-K + head gather (the producer of the per-person vectors)
"""
from . import synthetic  # noqa: F401
from . import model_io  # noqa: F401
from . import capi  # noqa: F401
from .smpl import SMPL, GraphedSMPL, HostRunner, StaticSMPL  # noqa: F401
from . import sharding  # noqa: F401
from .decode import decode_gather  # noqa: F401
from .dcn import DCN, DCNv2, dcn_v2_conv  # noqa: F401

__all__ = ["SMPL", "GraphedSMPL", "HostRunner", "StaticSMPL", "decode_gather", "DCN", "DCNv2", "dcn_v2_conv", "capi", "synthetic",
           "sharding", "model_io"]
