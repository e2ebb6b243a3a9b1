"""CPU derivation of the vertex bounds DESIGN.md §4 states for the tensor-core modes.

`oracle/precision_model.py` applies ONLY the operand rounding each mode is documented to do (everything else in
float64) and the result is compared with the float64 oracle.  The stated bounds (also asserted on the GPU by
tests/test_gpu_parity.py, tests/test_gpu_fused.py and in every bench run) must hold with room for the tensor core's
fp32 accumulation (~1e-6 m); the ordering of the modes must be what the design claims.
"""
import numpy as np
import pytest

from human_3d_reconstruction_b200 import synthetic
from oracle.precision_model import smpl_forward_mode, split
from oracle.smpl_np64 import smpl_forward_np64

# stated bounds, metres (VERT_ATOL in tests/test_gpu_parity.py, F16_ATOL in tests/test_gpu_fused.py)
STATED = {"f16x3": 4e-6, "bf16x3": 1e-5, "f16": 5e-5, "tf32": 5e-4, "bf16": 4e-3}
ACCUMULATION = 1.5e-6     # fp32 accumulation inside the tensor core, measured (f16x3: 1.9e-6 total)


@pytest.fixture(scope="module")
def errors(full_model):
    betas, pose, _ = synthetic.make_inputs(64, 1)
    ref = smpl_forward_np64(full_model, betas, pose)[0]
    return {m: float(np.abs(smpl_forward_mode(full_model, betas, pose, m) - ref).max()) for m in STATED}


@pytest.mark.parametrize("mode", list(STATED))
def test_operand_rounding_alone_stays_inside_the_stated_bound(errors, mode):
    assert errors[mode] + ACCUMULATION < STATED[mode], (mode, errors[mode])


def test_modes_rank_as_designed(errors):
    # split fp16 (11+11 bits) < split bf16 (8+8) < fused f16 (pose rows 11 bits) < tf32 (10 bits, truncated) < bf16 (8)
    assert errors["f16x3"] < errors["bf16x3"] < errors["f16"] < errors["tf32"] < errors["bf16"]
    # the fused kernel is an order of magnitude inside the regime BASELINE configs[2] names (TF32 / BF16 operands)
    assert errors["f16"] * 10 < errors["tf32"]


def test_fused_mode_meets_fp32_class_when_pose_is_zero(full_model):
    """Zero pose -> the single-product pose rows contribute nothing; only the exact splits remain."""
    betas, pose, _ = synthetic.make_inputs(16, 2)
    pose = np.zeros_like(pose)
    ref = smpl_forward_np64(full_model, betas, pose)[0]
    err = np.abs(smpl_forward_mode(full_model, betas, pose, "f16") - ref).max()
    assert err < 1e-6, err


@pytest.mark.parametrize("kind,bits", [("f16", 11), ("bf16", 8), ("tf32", 11)])
def test_split_pieces_reconstruct_the_value(kind, bits):
    rng = np.random.default_rng(0)
    x = rng.normal(0, 0.05, 4096).astype(np.float32)
    hi, lo = split(x, kind)
    # two pieces carry 2*bits significant bits (tf32 truncation: one bit less on the first piece)
    assert np.abs(hi + lo - x.astype(np.float64)).max() <= np.abs(x).max() * 2.0 ** (-2 * bits + 2)
    three = sum(split(x, kind, terms=3))
    # fp16 residuals below 2^-14 are sub-normal (spacing 2^-24): at most half a spacing = 3e-8 absolute is lost
    floor = 2.0 ** -25 if kind == "f16" else 0.0
    assert np.abs(three - x.astype(np.float64)).max() <= max(np.abs(x).max() * 2.0 ** -23, floor)
