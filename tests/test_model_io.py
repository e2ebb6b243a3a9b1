"""Model-file ingest: official SMPL layout <-> the eager layer's buffers."""
import pickle

import numpy as np
import torch

from human_3d_reconstruction_b200 import SMPL, model_io, synthetic
from oracle.smpl_ref import smpl_forward


def test_official_layout_round_trip(tmp_path, small_model):
    off = model_io.to_official_layout(small_model)
    V = small_model["v_template"].shape[0]
    assert off["shapedirs"].shape == (V, 3, 10) and off["posedirs"].shape == (V, 3, 207)
    assert off["J_regressor"].shape == (24, V) and off["kintree_table"][0, 0] == 2 ** 32 - 1
    back = model_io.from_official_layout(off)
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "parents"):
        np.testing.assert_array_equal(back[k], small_model[k])
    # .npz and pickle files, as a path
    np.savez(tmp_path / "m.npz", **off)
    with open(tmp_path / "m.pkl", "wb") as f:
        pickle.dump(off, f)
    for name in ("m.npz", "m.pkl"):
        got = model_io.load_model(str(tmp_path / name))
        np.testing.assert_array_equal(got["posedirs"], small_model["posedirs"])
        assert got["parents"][0] == -1
    layer = SMPL(str(tmp_path / "m.npz"))
    assert layer.num_verts == V and layer.num_betas == 10
    np.testing.assert_array_equal(layer.shapedirs.numpy(), small_model["shapedirs"])


def test_official_layout_column_order_matches_oracle(small_model):
    """shapedirs[V,3,NB] -> [NB,3V] must put vertex v, coordinate c at column 3v+c."""
    off = model_io.to_official_layout(small_model)
    betas, pose, _ = synthetic.make_inputs(3, 8)
    a = smpl_forward(model_io.from_official_layout(off), betas, pose, dtype=torch.float64)
    b = smpl_forward(small_model, betas, pose, dtype=torch.float64)
    assert torch.equal(a[0], b[0])
    # 300-dim shape space truncates to the first num_betas components
    off300 = dict(off)
    off300["shapedirs"] = np.concatenate([off["shapedirs"], np.ones_like(off["shapedirs"])], axis=2)
    np.testing.assert_array_equal(model_io.from_official_layout(off300)["shapedirs"], small_model["shapedirs"])


def test_sparse_regressor_is_densified(small_model):
    class FakeSparse:                      # stands in for scipy.sparse (official pickle)
        def __init__(self, a): self.a = a
        def toarray(self): return self.a
    off = model_io.to_official_layout(small_model)
    off["J_regressor"] = FakeSparse(off["J_regressor"])
    np.testing.assert_array_equal(model_io.from_official_layout(off)["J_regressor"], small_model["J_regressor"])


def test_official_pickle_loads_without_chumpy_or_scipy(tmp_path, small_model):
    """The release's .pkl holds chumpy.ch.Ch arrays and a scipy.sparse J_regressor; neither package is
    imported by the loader (chumpy is not installed here at all)."""
    import importlib.util
    import sys
    from _official_pkl import write_official_pickle
    assert importlib.util.find_spec("chumpy") is None
    for kind in ("fake-old-path", "scipy"):
        path = str(tmp_path / f"basicModel_{kind}.pkl")
        write_official_pickle(path, small_model, sparse=kind)
        assert "chumpy" not in sys.modules and "chumpy.ch" not in sys.modules
        got = model_io.load_model(path)
        for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights"):
            np.testing.assert_array_equal(got[k], small_model[k])
            assert got[k].dtype == np.float32
        assert got["parents"][0] == -1 and list(got["parents"][1:]) == list(small_model["parents"][1:])
    layer = SMPL(path)
    np.testing.assert_array_equal(layer.J_regressor.numpy(), small_model["J_regressor"])


def test_model_pickle_refuses_other_globals(tmp_path):
    """Unpickling a model file must not be able to run code: any global outside the whitelist raises."""
    import os
    import pytest

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned > %s" % (tmp_path / "pwned"),))

    p = tmp_path / "evil.pkl"
    with open(p, "wb") as f:
        pickle.dump({"v_template": Evil()}, f, protocol=2)
    with pytest.raises(pickle.UnpicklingError):
        model_io.load_model(str(p))
    assert not (tmp_path / "pwned").exists()
