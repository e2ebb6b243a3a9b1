"""DCNv2 forward oracle (oracle/dcn_ref.py) pinned against torchvision's CPU operator, the committed
golden fixture and the reference's own known-answer test.  CPU only."""
import os

import numpy as np
import pytest
import torch
from torchvision.ops import deform_conv2d

from oracle.dcn_ref import dcn_forward, dcn_v2_forward

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dcn_golden_v1.npz")


@pytest.mark.parametrize("stride,padding,dilation,dg", [(1, 1, 1, 1), (2, 1, 1, 1), (1, 2, 2, 1), (1, 1, 1, 2)])
def test_oracle_matches_torchvision(stride, padding, dilation, dg):
    g = torch.Generator().manual_seed(5)
    B, Ci, Co, H, W = 2, 8, 6, 9, 7
    x = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Co, Ci, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(Co, generator=g, dtype=torch.float64)
    Ho = (H + 2 * padding - (dilation * 2 + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * 2 + 1)) // stride + 1
    off = torch.randn(B, dg * 18, Ho, Wo, generator=g, dtype=torch.float64) * 2.5
    m = torch.rand(B, dg * 9, Ho, Wo, generator=g, dtype=torch.float64)
    ref = deform_conv2d(x, off, w, b, stride=stride, padding=padding, dilation=dilation, mask=m)
    got = dcn_v2_forward(x, w, b, off, m, stride, padding, dilation, dg, dtype=torch.float64)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_oracle_matches_golden_fixture():
    z = np.load(GOLDEN)
    for n in range(3):
        x, w, b, off, m, y = (torch.from_numpy(z[f"c{n}_{k}"]) for k in ("x", "w", "b", "off", "m", "y"))
        got = dcn_v2_forward(x, w, b, off, m, dtype=torch.float64)
        assert torch.allclose(got, y, rtol=1e-10, atol=1e-10), n
        got32 = dcn_v2_forward(x, w, b, off, m, dtype=torch.float32)
        assert torch.allclose(got32.double(), y, rtol=1e-4, atol=1e-4), n


def test_reference_known_answer_zero_offset_identity_kernel():
    """reference src/lib/models/DCNv2/test.py:31-66: zero offsets, mask = sigmoid(0) = 0.5, identity
    kernel  =>  2 * output == input."""
    g = torch.Generator().manual_seed(0)
    N, C, H, W = 2, 2, 4, 4
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.zeros(C, C, 3, 3)
    for p in range(C):
        w[p, p, 1, 1] = 1.0
    out = dcn_v2_forward(x, w, torch.zeros(C), torch.zeros(N, 18, H, W), torch.sigmoid(torch.zeros(N, 9, H, W)))
    assert (x - 2 * out).abs().max().item() < 1e-10


def test_zero_offsets_unit_mask_is_a_plain_convolution():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 8, 6, generator=g, dtype=torch.float64)
    w = torch.randn(7, 5, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(7, generator=g, dtype=torch.float64)
    got = dcn_v2_forward(x, w, b, torch.zeros(2, 18, 8, 6), torch.ones(2, 9, 8, 6), dtype=torch.float64)
    ref = torch.nn.functional.conv2d(x, w, b, padding=1)
    assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_dcn_module_forward_with_zero_initialised_offset_conv():
    """reference dcn_v2.py:113-116 zero-initialises conv_offset_mask: DCN(x) == 0.5 * conv(x)."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 4, 6, 6, generator=g)
    w = torch.randn(3, 4, 3, 3, generator=g)
    b = torch.randn(3, generator=g)
    got = dcn_forward(x, torch.zeros(27, 4, 3, 3), torch.zeros(27), w, b)
    ref = 0.5 * torch.nn.functional.conv2d(x, w, None, padding=1) + b.view(1, 3, 1, 1)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_oracle_vs_torchvision_random_configurations():
    """Random small shapes / strides / paddings / dilations / groups / offset scales (seeded)."""
    rng = torch.Generator().manual_seed(2026)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=rng))
    for _ in range(12):
        dg = ri(1, 2)
        B, Ci, Co = ri(1, 2), dg * ri(1, 3), ri(1, 5)
        H, W = ri(3, 9), ri(3, 9)
        stride, padding, dilation = ri(1, 2), ri(0, 2), ri(1, 2)
        Ho = (H + 2 * padding - (dilation * 2 + 1)) // stride + 1
        Wo = (W + 2 * padding - (dilation * 2 + 1)) // stride + 1
        if Ho < 1 or Wo < 1:
            continue
        x = torch.randn(B, Ci, H, W, generator=rng, dtype=torch.float64)
        w = torch.randn(Co, Ci, 3, 3, generator=rng, dtype=torch.float64)
        b = torch.randn(Co, generator=rng, dtype=torch.float64)
        off = torch.randn(B, dg * 18, Ho, Wo, generator=rng, dtype=torch.float64) * float(ri(0, 6))
        m = torch.rand(B, dg * 9, Ho, Wo, generator=rng, dtype=torch.float64)
        ref = deform_conv2d(x, off, w, b, stride=stride, padding=padding, dilation=dilation, mask=m)
        got = dcn_v2_forward(x, w, b, off, m, stride, padding, dilation, dg, dtype=torch.float64)
        assert torch.allclose(got, ref, rtol=1e-11, atol=1e-11), (B, Ci, Co, H, W, stride, padding, dilation, dg)
