"""Generates tests/golden/dla_golden_v1.npz from the UNMODIFIED reference network.

Runs only in the build container (needs /root/reference):  python tests/golden/make_dla_golden.py
Imports reference src/lib/models/model.py (`dla_net`, not_use_dcn=True -- the DCN extension does not
compile on torch 2.x) read-only and
  1. checks that oracle/dla34_ref.py, seeded the same way (317 = reference opts.py:37), builds
     parameters that are EQUAL tensor by tensor under the SAME state_dict keys;
  2. checks that, in eval mode on a seeded input, both produce bit-identical head maps on the CPU;
  3. records: per-head float64 checksums of the reference's parameters, and the reference's head maps
     for one small seeded input (1 x 3 x 64 x 64 -> six heads at 16 x 16) -- what
     tests/test_dla_oracle.py re-checks on any box without the reference.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dla_golden_v1.npz")
SEED = 317


def golden_input(seed=5, shape=(1, 3, 64, 64)):
    return torch.from_numpy(np.random.default_rng(seed).normal(0.0, 1.0, size=shape).astype(np.float32))


def param_checksums(net):
    """float64 (sum, sum of squares) over every state_dict tensor, grouped by top-level module."""
    groups = {}
    for k, v in net.state_dict().items():
        g = k.split(".")[0]
        s, q, n = groups.get(g, (0.0, 0.0, 0))
        v = v.double()
        groups[g] = (s + float(v.sum()), q + float((v * v).sum()), n + v.numel())
    return groups


def main():
    from oracle.dla34_ref import HEADS_HMR, dla_net as dla_restated
    sys.path.insert(0, "/root/reference/src/lib")
    from models.model import dla_net as dla_reference       # noqa: E402  (the real reference)
    torch.manual_seed(SEED)
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints its DCN mode
        ref = dla_reference(dict(HEADS_HMR), num_layers=34, head_conv=256, down_ratio=4, not_use_dcn=True)
    mine = dla_restated(dict(HEADS_HMR), seed=SEED)
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys()), "state_dict keys / order differ"
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_mine[k]), f"seeded initialisation differs at {k}"
    mine.load_state_dict(sd_ref, strict=True)
    ref.eval(); mine.eval()
    x = golden_input()
    with torch.no_grad():
        out_ref, out_mine = ref(x)[0], mine(x)[0]
    for h in HEADS_HMR:
        assert torch.equal(out_ref[h], out_mine[h]), f"head {h} differs from the reference"
    rec = {f"head_{h}": out_ref[h].numpy() for h in HEADS_HMR}
    for g, (s, q, n) in param_checksums(ref).items():
        rec[f"param_{g}"] = np.array([s, q, n], dtype=np.float64)
    rec["num_params"] = np.array([sum(p.numel() for p in ref.parameters())])
    np.savez_compressed(OUT, **rec)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", int(rec["num_params"][0]), "parameters;",
          "restatement == reference (weights at seed 317, and head maps bit for bit)")


if __name__ == "__main__":
    main()
