"""DCNv2 forward on B200: drop-in for the reference's one native op (inference).

Mirrors the Python surface of reference src/lib/models/DCNv2/dcn_v2.py -- ``dcn_v2_conv(input, offset,
mask, weight, bias, stride, padding, dilation, deformable_groups)`` (:16-54), ``DCNv2`` (:57-98) and
``DCN`` (:101-130, the class the network instantiates at reference src/lib/models/model.py:355) --
over ``smplb200_dcn_v2_forward`` (include/smpl_b200.h), one fused implicit-GEMM kernel
(csrc/k_dcn.cuh).  The reference's extension itself no longer compiles on torch 2.x (THC headers), so
this is also what lets the DLA neck run with ``USE_DCN = True`` on this stack.

Forward only: the DCN backward (gradients w.r.t. input/offset/mask/weight/bias, reference
dcn_v2.py:36-54) is not built; asking for gradients raises.  CUDA only, no CPU fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import capi


def _pair(x):
    return (int(x), int(x)) if not isinstance(x, (tuple, list)) else (int(x[0]), int(x[1]))


def dcn_v2_conv(input, offset, mask, weight, bias, stride=1, padding=1, dilation=1, deformable_groups=1):
    """output[B,Co,Ho,Wo]; offset[B,2*dg*kh*kw,Ho,Wo] (dh, dw interleaved per tap), mask[B,dg*kh*kw,Ho,Wo]."""
    if not isinstance(input, torch.Tensor) or input.device.type != "cuda":
        raise RuntimeError("dcn_v2_conv (B200) needs CUDA tensors; there is no CPU fallback")
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (input, offset, mask, weight, bias)):
        raise RuntimeError("dcn_v2_conv (B200) is forward-only: wrap the call in torch.no_grad() "
                           "(the DCN backward is not built)")
    dev = input.device
    tensors = [input, offset, mask, weight] + ([bias] if bias is not None else [])
    for t in tensors:
        if t.device != dev or t.dtype != torch.float32:
            raise TypeError("dcn_v2_conv: all tensors must be float32 on the input's device")
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    B, Ci, H, W = input.shape
    Co, Ciw, kh, kw = weight.shape
    if Ciw != Ci:
        raise ValueError(f"input has {Ci} channels, weight expects {Ciw}")
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    dg = int(deformable_groups)
    if tuple(offset.shape) != (B, 2 * dg * kh * kw, Ho, Wo) or tuple(mask.shape) != (B, dg * kh * kw, Ho, Wo):
        raise ValueError("offset / mask shapes do not match the output size and kernel")
    # a channels_last input is sampled in place (its memory IS [B,H,W,Ci]); anything else is copied
    flags = 0
    if not input.is_contiguous() and input.is_contiguous(memory_format=torch.channels_last):
        flags = capi.DCN_INPUT_NHWC
    else:
        input = input.contiguous()
    offset, mask, weight = (t.contiguous() for t in (offset, mask, weight))
    bias = None if bias is None else bias.contiguous()
    out = torch.empty((B, Co, Ho, Wo), dtype=torch.float32, device=dev)
    lib = capi.lib()
    wsb = int(lib.smplb200_dcn_v2_workspace_bytes(B, Ci, H, W, Co, flags))
    if wsb == 0:
        raise RuntimeError("dcn_v2_conv (B200): unsupported shape (needs 3x3, Ci % 32 == 0, Co % 16 == 0, Co <= 256)")
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    capi.check(lib.smplb200_dcn_v2_forward(
        idx, input.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
        offset.data_ptr(), mask.data_ptr(), B, Ci, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, dg,
        out.data_ptr(), ws.data_ptr(), wsb, flags, torch.cuda.current_stream(dev).cuda_stream),
        "smplb200_dcn_v2_forward")
    return out


class DCNv2(nn.Module):
    """Modulated deformable convolution with externally supplied offsets and mask."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__()
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.deformable_groups = int(deformable_groups)
        self.weight = nn.Parameter(torch.empty(self.out_channels, self.in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(self.out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        # same initialisation as the reference (dcn_v2.py:75-81): U(-1/sqrt(fan_in), 1/sqrt(fan_in)), zero bias
        bound = 1.0 / math.sqrt(self.in_channels * self.kernel_size[0] * self.kernel_size[1])
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            self.bias.zero_()

    def forward(self, input, offset, mask):
        return dcn_v2_conv(input, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups)


class DCN(DCNv2):
    """DCNv2 whose offsets and mask come from a plain convolution of the input (zero-initialised)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
        k = self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(self.in_channels, self.deformable_groups * 3 * k, kernel_size=self.kernel_size,
                                          stride=self.stride, padding=self.padding, bias=True)
        with torch.no_grad():
            self.conv_offset_mask.weight.zero_()
            self.conv_offset_mask.bias.zero_()

    def forward(self, input):
        o1, o2, m = torch.chunk(self.conv_offset_mask(input), 3, dim=1)
        return dcn_v2_conv(input, torch.cat((o1, o2), dim=1), torch.sigmoid(m), self.weight, self.bias, self.stride,
                           self.padding, self.dilation, self.deformable_groups)
