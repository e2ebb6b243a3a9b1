"""The decode -> gather oracle (oracle/decode_ref.py) against golden vectors produced by the
UNMODIFIED reference functions (tests/golden/make_decode_golden.py), bit for bit."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from oracle.decode_ref import decode_gather

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_decode_golden", os.path.join(HERE, "golden", "make_decode_golden.py"))
gold = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gold)
GOLDEN = np.load(os.path.join(HERE, "golden", "decode_golden_v1.npz"))


@pytest.mark.parametrize("ci", range(len(gold.CASES)))
def test_oracle_matches_reference_golden(ci):
    seed, B, C, H, W, chans, K = gold.CASES[ci]
    heat, heads = gold.decode_case(seed, B, C, H, W, chans, K)
    scores, inds, clses, ys, xs, feats = decode_gather(torch.from_numpy(heat), [torch.from_numpy(h) for h in heads], K)
    np.testing.assert_array_equal(scores.numpy(), GOLDEN[f"c{ci}_scores"])
    np.testing.assert_array_equal(inds.numpy(), GOLDEN[f"c{ci}_inds"])
    np.testing.assert_array_equal(clses.numpy(), GOLDEN[f"c{ci}_clses"])
    np.testing.assert_array_equal(ys.numpy(), GOLDEN[f"c{ci}_ys"])
    np.testing.assert_array_equal(xs.numpy(), GOLDEN[f"c{ci}_xs"])
    for hi, f in enumerate(feats):
        np.testing.assert_array_equal(f.numpy(), GOLDEN[f"c{ci}_head{hi}"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/lib"), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference/src/lib")
    try:
        from models.decode import _nms, _topk
        from models.utils import _transpose_and_gather_feat
    finally:
        sys.path.pop(0)
    heat, heads = gold.decode_case(777, 2, 2, 48, 36, (7, 3), 12)
    th = torch.from_numpy(heat)
    ref = _topk(_nms(th), K=12)
    mine = decode_gather(th, [torch.from_numpy(h) for h in heads], 12)
    for a, b in zip(ref, mine[:5]):
        assert torch.equal(a, b)
    for h, f in zip(heads, mine[5]):
        assert torch.equal(_transpose_and_gather_feat(torch.from_numpy(h), ref[1]), f)
