"""Model-file ingest: official SMPL layouts -> the eager layer's buffers.

The official SMPL release (and smplx / HMR conversions of it) stores

    v_template   [V, 3]
    shapedirs    [V, 3, NB']      (NB' = 10, or 300 for the full shape space)
    posedirs     [V, 3, 207]
    J_regressor  [24, V]          dense or scipy.sparse
    weights      [V, 24]
    kintree_table[2, 24]          row 0 = parents, root stored as 2**32 - 1
    f            [F, 3]           faces (not needed by the forward pass)

while the eager PyTorch layer this package replaces keeps

    shapedirs [NB, 3V]   posedirs [207, 3V]   J_regressor [V, 24]   parents [24] (root = -1)

with the flattened vertex axis ordered 3*v + c.  `load_model` accepts an ``.npz`` or a pickle of a
dict in either layout and returns the dict `SMPL(...)` takes.  The licensed model file itself is
not redistributable and is not available offline, so this path is exercised with synthetic models
written in the official layout (`to_official_layout`); the reference keeps an (empty) `models/`
directory for the real file (reference models/.gitignore:1-2).
"""
from __future__ import annotations

import pickle

import numpy as np


def _dense(a):
    if hasattr(a, "toarray"):          # scipy.sparse matrix (official pickle)
        a = a.toarray()
    if hasattr(a, "r"):                # chumpy array: .r is the numpy value
        a = a.r
    return np.asarray(a)


def from_official_layout(d: dict, num_betas: int = 10) -> dict:
    """Official SMPL arrays -> the eager layer's buffers (float32, root parent = -1)."""
    vt = _dense(d["v_template"]).astype(np.float32)
    V = vt.shape[0]
    sd = _dense(d["shapedirs"]).astype(np.float32)
    pd = _dense(d["posedirs"]).astype(np.float32)
    jr = _dense(d["J_regressor"]).astype(np.float32)
    w = _dense(d["weights"]).astype(np.float32)
    if sd.ndim == 3:                                   # [V,3,NB'] -> [NB,3V]
        sd = sd[:, :, :num_betas].reshape(V * 3, -1).T
    if pd.ndim == 3:                                   # [V,3,207] -> [207,3V]
        pd = pd.reshape(V * 3, -1).T
    if jr.shape == (w.shape[1], V):                    # [24,V] -> [V,24]
        jr = jr.T
    if "parents" in d:
        parents = np.asarray(_dense(d["parents"])).astype(np.int64)
    else:
        parents = np.asarray(_dense(d["kintree_table"]))[0].astype(np.int64)
    parents = parents.copy()
    parents[0] = -1                                    # 2**32-1 / -1 / anything: joint 0 is the root
    out = {"v_template": vt, "shapedirs": np.ascontiguousarray(sd), "posedirs": np.ascontiguousarray(pd),
           "J_regressor": np.ascontiguousarray(jr), "weights": w, "parents": parents.astype(np.int32)}
    J = w.shape[1]
    if out["shapedirs"].shape[1] != 3 * V or out["posedirs"].shape != (9 * (J - 1), 3 * V) \
            or out["J_regressor"].shape != (V, J) or parents.shape != (J,):
        raise ValueError("unrecognised SMPL model layout")
    return out


def to_official_layout(model: dict) -> dict:
    """Inverse of `from_official_layout` (used to write test fixtures in the official layout)."""
    V = model["v_template"].shape[0]
    parents = np.asarray(model["parents"]).astype(np.int64)
    kintree = np.stack([np.where(parents < 0, 2 ** 32 - 1, parents), np.arange(parents.shape[0])]).astype(np.uint32)
    return {"v_template": np.asarray(model["v_template"]),
            "shapedirs": np.asarray(model["shapedirs"]).T.reshape(V, 3, -1),
            "posedirs": np.asarray(model["posedirs"]).T.reshape(V, 3, -1),
            "J_regressor": np.asarray(model["J_regressor"]).T,
            "weights": np.asarray(model["weights"]), "kintree_table": kintree}


def load_model(path: str, num_betas: int = 10) -> dict:
    """Load an ``.npz`` or a pickled dict (official SMPL layout or the eager layer's) from disk."""
    if str(path).endswith(".npz"):
        with np.load(path, allow_pickle=False) as z:
            d = {k: z[k] for k in z.files}
    else:
        with open(path, "rb") as f:
            d = pickle.load(f, encoding="latin1")
    return from_official_layout(d, num_betas=num_betas)
