"""GPU parity of the fused DCNv2 forward (-m gpu) through the C ABI vs the CPU oracle (pinned against
torchvision and the reference's own known-answer test in tests/test_dcn_oracle.py) and the golden
fixture generated from torchvision.

Tolerance: the kernel multiplies split-bf16 operands (hi*hi + lo*hi + hi*lo, ~2^-16 relative per
product, fp32 accumulate): every output must agree with the float64 oracle to 5e-5 of the largest
|output| of the case (DCN_RTOL; measured 0.3-1.7e-5); measured errors are printed by bench.py / smoke().
"""
import os

import numpy as np
import pytest
import torch

from human_3d_reconstruction_b200 import DCN, DCNv2, capi, dcn_v2_conv
from oracle.dcn_ref import dcn_forward, dcn_v2_forward

pytestmark = pytest.mark.gpu
DCN_RTOL = 5e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dcn_golden_v1.npz")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def close(got, ref, what):
    got, ref = got.detach().cpu().double(), ref.double()
    assert got.shape == ref.shape, what
    assert torch.isfinite(got).all(), what
    scale = max(ref.abs().max().item(), 1e-12)
    err = (got - ref).abs().max().item()
    assert err <= DCN_RTOL * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"
    return err / scale


def rand_case(B, Ci, Co, H, W, off_scale, seed, stride=1, padding=1, dilation=1):
    g = torch.Generator().manual_seed(seed)
    Ho = (H + 2 * padding - (dilation * 2 + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * 2 + 1)) // stride + 1
    x = torch.randn(B, Ci, H, W, generator=g)
    w = torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5
    b = torch.randn(Co, generator=g)
    off = torch.randn(B, 18, Ho, Wo, generator=g) * off_scale
    m = torch.rand(B, 9, Ho, Wo, generator=g)
    return x, w, b, off, m


def test_golden_fixture(dev):
    z = np.load(GOLDEN)
    for n in range(3):
        x, w, b, off, m, y = (torch.from_numpy(z[f"c{n}_{k}"]) for k in ("x", "w", "b", "off", "m", "y"))
        with torch.no_grad():
            got = dcn_v2_conv(x.to(dev), off.to(dev), m.to(dev), w.to(dev), b.to(dev))
        close(got, y, f"golden case {n}")


@pytest.mark.parametrize("B,Ci,Co,H,W,sc", [
    (1, 32, 16, 5, 7, 1.0),        # one partial tile
    (2, 64, 64, 16, 16, 2.0),      # DLA level-0 channel count
    (3, 128, 256, 9, 13, 4.0),     # widest output, odd map, tiles straddle images
    (1, 512, 256, 8, 8, 1.0),      # deepest DLA projection (K = 4608)
    (2, 32, 48, 33, 31, 12.0),     # huge offsets: most samples outside the map
])
def test_vs_oracle(dev, B, Ci, Co, H, W, sc):
    x, w, b, off, m = rand_case(B, Ci, Co, H, W, sc, 100 + Ci + Co)
    ref = dcn_v2_forward(x, w, b, off, m, dtype=torch.float64)
    with torch.no_grad():
        got = dcn_v2_conv(x.to(dev), off.to(dev), m.to(dev), w.to(dev), b.to(dev))
    close(got, ref, f"B{B} Ci{Ci} Co{Co} {H}x{W}")


@pytest.mark.parametrize("stride,padding,dilation", [(2, 1, 1), (1, 2, 2), (1, 0, 1)])
def test_stride_padding_dilation(dev, stride, padding, dilation):
    x, w, b, off, m = rand_case(2, 32, 32, 14, 12, 1.5, 7, stride, padding, dilation)
    ref = dcn_v2_forward(x, w, b, off, m, stride, padding, dilation, dtype=torch.float64)
    with torch.no_grad():
        got = dcn_v2_conv(x.to(dev), off.to(dev), m.to(dev), w.to(dev), b.to(dev), stride, padding, dilation)
    close(got, ref, f"s{stride} p{padding} d{dilation}")


def test_reference_known_answer_and_plain_convolution(dev):
    """reference DCNv2/test.py:31-66 (zero offsets, mask 0.5, identity kernel => 2*out == input), at the
    smallest shape the kernel supports; and zero offsets + unit mask == conv2d."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 32, 4, 4, generator=g)
    w = torch.zeros(32, 32, 3, 3)
    for p in range(32):
        w[p, p, 1, 1] = 1.0
    with torch.no_grad():
        out = dcn_v2_conv(x.to(dev), torch.zeros(2, 18, 4, 4, device=dev), torch.full((2, 9, 4, 4), 0.5, device=dev),
                          w.to(dev), torch.zeros(32, device=dev))
    # the reference's fp32 GEMM meets 1e-10 here; the split-bf16 operands carry 16 mantissa bits:
    # |x - 2*out| <= 2^-16 * max|x|
    assert (x - 2 * out.cpu()).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()
    w2 = torch.randn(48, 32, 3, 3, generator=g) / 17.0
    b2 = torch.randn(48, generator=g)
    with torch.no_grad():
        got = dcn_v2_conv(x.to(dev), torch.zeros(2, 18, 4, 4, device=dev), torch.ones(2, 9, 4, 4, device=dev),
                          w2.to(dev), b2.to(dev))
    close(got, torch.nn.functional.conv2d(x.double(), w2.double(), b2.double(), padding=1), "plain conv")


def test_network_layer_shape_and_linearity(dev):
    """The network's largest DeformConv (64 -> 64 @ 128x128): oracle parity on a 2-image batch and the
    size-independent properties of the operator: exact homogeneity in the input and in the mask (powers
    of two), additivity in the input up to rounding, bias-only output for a zero mask."""
    x, w, b, off, m = rand_case(2, 64, 64, 128, 128, 2.0, 77)
    ref = dcn_v2_forward(x, w, b, off, m, dtype=torch.float64)
    xd, wd, bd, od, md = (t.to(dev) for t in (x, w, b, off, m))
    zb = torch.zeros_like(bd)
    with torch.no_grad():
        y = dcn_v2_conv(xd, od, md, wd, bd)
        close(y, ref, "64->64 @128x128")
        y0 = dcn_v2_conv(xd, od, md, wd, zb)
        assert torch.equal(dcn_v2_conv(2.0 * xd, od, md, wd, zb), 2.0 * y0)
        assert torch.equal(dcn_v2_conv(xd, od, 0.5 * md, wd, zb), 0.5 * y0)
        x2 = torch.randn_like(xd)
        ya = dcn_v2_conv(xd + x2, od, md, wd, zb)
        yb = y0 + dcn_v2_conv(x2, od, md, wd, zb)
        assert (ya - yb).abs().max().item() <= 1e-4 * ya.abs().max().item()
        only_bias = dcn_v2_conv(xd, od, torch.zeros_like(md), wd, bd)
        assert torch.equal(only_bias, bd.view(1, -1, 1, 1).expand_as(only_bias))


def test_channels_last_input_is_sampled_in_place(dev):
    x, w, b, off, m = rand_case(2, 64, 32, 10, 9, 2.0, 21)
    ref = dcn_v2_forward(x, w, b, off, m, dtype=torch.float64)
    xd = x.to(dev)
    with torch.no_grad():
        a = dcn_v2_conv(xd, off.to(dev), m.to(dev), w.to(dev), b.to(dev))
        c = dcn_v2_conv(xd.contiguous(memory_format=torch.channels_last), off.to(dev), m.to(dev), w.to(dev), b.to(dev))
    close(a, ref, "NCHW input")
    assert torch.equal(a, c), "the channels-last path must give the same bits"


def test_modules(dev):
    torch.manual_seed(11)
    layer = DCN(64, 32, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1).to(dev)
    with torch.no_grad():
        layer.conv_offset_mask.weight.normal_(0, 0.05)
        layer.conv_offset_mask.bias.normal_(0, 0.5)
        layer.bias.normal_()
    x = torch.randn(2, 64, 12, 10, device=dev)
    with torch.no_grad():
        got = layer(x)
    ref = dcn_forward(x.cpu(), layer.conv_offset_mask.weight.cpu(), layer.conv_offset_mask.bias.cpu(),
                      layer.weight.detach().cpu(), layer.bias.detach().cpu(), dtype=torch.float64)
    close(got, ref, "DCN module")
    plain = DCNv2(32, 16, 3, 1, 1).to(dev)
    assert {"weight", "bias"} == set(dict(plain.named_parameters()))
    with pytest.raises(RuntimeError, match="forward-only"):
        layer(x.requires_grad_())


def test_errors_and_unsupported(dev):
    lib = capi.lib()
    assert lib.smplb200_dcn_v2_workspace_bytes(1, 48, 4, 4, 32, 0) == 0          # Ci not a multiple of 32
    assert lib.smplb200_dcn_v2_workspace_bytes(1, 64, 4, 4, 512, 0) == 0         # Co > 256
    assert lib.smplb200_dcn_v2_workspace_bytes(2, 64, 8, 8, 32, 0) > lib.smplb200_dcn_v2_workspace_bytes(2, 64, 8, 8, 32, 1)
    z = lambda *s: torch.zeros(*s, device=dev)
    x, w, off, m, out = z(1, 32, 4, 4), z(16, 32, 3, 3), z(1, 18, 4, 4), z(1, 9, 4, 4), z(1, 16, 4, 4)
    ws = torch.empty(lib.smplb200_dcn_v2_workspace_bytes(1, 32, 4, 4, 16, 0), dtype=torch.uint8, device=dev)
    args = lambda dg, wsp, wsn: (0, x.data_ptr(), w.data_ptr(), None, off.data_ptr(), m.data_ptr(), 1, 32, 4, 4, 16,
                                  3, 3, 1, 1, 1, 1, 1, 1, dg, out.data_ptr(), wsp, wsn, 0, None)
    assert lib.smplb200_dcn_v2_forward(*args(2, ws.data_ptr(), ws.numel())) == 2      # deformable groups
    assert lib.smplb200_dcn_v2_forward(*args(1, None, 0)) == 3                        # workspace
    assert lib.smplb200_dcn_v2_forward(*args(1, ws.data_ptr(), ws.numel())) == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dcn_v2_conv(torch.zeros(1, 32, 4, 4), off.cpu(), m.cpu(), w.cpu(), None)
