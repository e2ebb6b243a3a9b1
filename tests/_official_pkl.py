"""Test helper: write a pickle shaped like the official SMPL release (chumpy arrays + a scipy-sparse
J_regressor, protocol 2) WITHOUT chumpy installed: a throw-away class is registered under the
module path ``chumpy.ch`` while pickling and removed again, so the loader sees exactly the globals
the real file references (``chumpy.ch.Ch``, ``scipy.sparse.csc.csc_matrix``) and cannot import them.
"""
import pickle
import sys
import types

import numpy as np

from human_3d_reconstruction_b200 import model_io


def write_official_pickle(path, model, sparse="fake-old-path"):
    off = model_io.to_official_layout(model)
    mods = {}

    def fake_module(name):
        m = types.ModuleType(name)
        mods[name] = sys.modules.get(name)
        sys.modules[name] = m
        return m

    try:
        ch = fake_module("chumpy.ch")
        pkg = fake_module("chumpy")
        pkg.ch = ch

        class Ch(object):                      # pickles as copy_reg._reconstructor + a state dict with 'x'
            def __init__(self, x):
                self.x = np.asarray(x)
                self._dirty_vars = set()      # the real class carries bookkeeping attributes as well
                self._itr = None

        Ch.__module__, Ch.__qualname__ = "chumpy.ch", "Ch"
        ch.Ch = Ch
        d = dict(off)
        for k in ("v_template", "shapedirs", "posedirs", "weights"):
            d[k] = Ch(off[k].astype(np.float64))          # the release stores float64
        jr = off["J_regressor"].astype(np.float64)
        if sparse == "scipy":
            import scipy.sparse as sp
            d["J_regressor"] = sp.csc_matrix(jr)
        else:
            spm = fake_module("scipy.sparse.csc")          # the module path old scipy pickles reference

            class csc_matrix(object):
                pass

            csc_matrix.__module__, csc_matrix.__qualname__ = "scipy.sparse.csc", "csc_matrix"
            spm.csc_matrix = csc_matrix
            obj = csc_matrix()
            cols = [np.nonzero(jr[:, j])[0] for j in range(jr.shape[1])]
            obj.indices = np.concatenate(cols).astype(np.int32)
            obj.indptr = np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.int32)
            obj.data = np.concatenate([jr[c, j] for j, c in enumerate(cols)])
            obj._shape = jr.shape
            obj.maxprint = 50
            d["J_regressor"] = obj
        d["f"] = np.zeros((4, 3), dtype=np.uint32)
        d["bs_style"], d["bs_type"] = "lbs", "lrotmin"
        with open(path, "wb") as f:
            pickle.dump(d, f, protocol=2)
    finally:
        for name, old in mods.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
    return off
