"""oracle/dla34_ref.py (the configs[4] producer network) vs the fixture generated from the UNMODIFIED
reference `dla_net` (tests/golden/make_dla_golden.py: equal seeded weights, bit-identical head maps)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_dla_golden import SEED, golden_input, param_checksums   # noqa: E402
from oracle.dla34_ref import HEADS_HMR, dla_net                    # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dla_golden_v1.npz")


def test_seeded_init_and_head_maps_match_the_reference_fixture():
    g = np.load(GOLDEN)
    net = dla_net(dict(HEADS_HMR), seed=SEED).eval()
    assert sum(p.numel() for p in net.parameters()) == int(g["num_params"][0])
    for grp, (s, q, n) in param_checksums(net).items():
        ref = g[f"param_{grp}"]
        assert n == int(ref[2]), grp
        np.testing.assert_allclose([s, q], ref[:2], rtol=1e-9, atol=1e-9, err_msg=f"seeded init of {grp}")
    with torch.no_grad():
        out = net(golden_input())[0]
    for h, ch in HEADS_HMR.items():
        ref = g[f"head_{h}"]
        assert out[h].shape == ref.shape == (1, ch, 16, 16)
        # same ops in the same order: equal up to the conv kernels oneDNN picks for this CPU
        np.testing.assert_allclose(out[h].numpy(), ref, rtol=1e-4, atol=1e-6, err_msg=h)


def test_network_shape_is_the_reference_dla34():
    net = dla_net(dict(HEADS_HMR), seed=1)
    from oracle.dla34_ref import DeformConv
    necks = [m for m in net.modules() if isinstance(m, DeformConv)]
    assert len(necks) == 16                                # 12 in dla_up + 4 in ida_up (reference model.py:365-416)
    shapes = sorted({(m.conv.in_channels, m.conv.out_channels) for m in necks})
    assert shapes == [(64, 64), (128, 64), (128, 128), (256, 64), (256, 128), (256, 256), (512, 256)]
    with torch.no_grad():
        out = net.eval()(torch.zeros(1, 3, 96, 64))[0]
    assert out["pose"].shape == (1, 72, 24, 16) and out["hm"].shape == (1, 1, 24, 16)   # down_ratio 4


import pytest  # noqa: E402


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/lib"), reason="reference tree only exists in the build container")
def test_restatement_equals_the_unmodified_reference_network():
    """Where the reference is importable: same state_dict keys, equal seeded weights, bit-identical head maps."""
    import contextlib
    import io
    sys.path.insert(0, "/root/reference/src/lib")
    from models.model import dla_net as dla_reference
    torch.manual_seed(SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = dla_reference(dict(HEADS_HMR), num_layers=34, head_conv=256, down_ratio=4, not_use_dcn=True).eval()
    mine = dla_net(dict(HEADS_HMR), seed=SEED).eval()
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys())
    assert all(torch.equal(sd_ref[k], sd_mine[k]) for k in sd_ref)
    x = golden_input()
    with torch.no_grad():
        a, b = ref(x)[0], mine(x)[0]
    assert all(torch.equal(a[h], b[h]) for h in HEADS_HMR)
