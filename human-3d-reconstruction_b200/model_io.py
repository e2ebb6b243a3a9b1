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
dict in either layout and returns the dict `SMPL(...)` takes.

The official ``.pkl`` is a Python-2 pickle whose arrays are ``chumpy.ch.Ch`` objects and whose
``J_regressor`` is a ``scipy.sparse.csc_matrix``.  Neither package is needed (chumpy is not
installable on current Pythons): the file is read by a RESTRICTED unpickler (`_ModelUnpickler`) that
resolves exactly four families of globals -- numpy's array reconstructors, copyreg's
``_reconstructor`` plus inert builtin containers, ``chumpy.*`` classes (mapped to a stub that keeps the ``x`` array) and scipy's
csc/csr/coo matrices (mapped to a stub that densifies from data/indices/indptr) -- and refuses
every other global, so loading a model file cannot execute code from it.  The licensed model file itself is
not redistributable and is not available offline, so this path is exercised with synthetic models
written in the official layout (`to_official_layout`); the reference keeps an (empty) `models/`
directory for the real file (reference models/.gitignore:1-2).
"""
from __future__ import annotations

import pickle

import numpy as np


class _ChumpyStub:
    """Stands in for any ``chumpy.*`` class: keeps the pickled attribute dict; ``.r`` is the value."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict, slots)
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)

    @property
    def r(self):
        if "x" not in self.__dict__:
            raise pickle.UnpicklingError("chumpy object without a stored value (.x): a computed "
                                         "expression, not a plain array -- re-export the model as .npz")
        return np.asarray(self.__dict__["x"])


class _SparseStub:
    """Stands in for scipy.sparse csc/csr/coo matrices: densifies from the pickled index arrays."""
    fmt = "csc"

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)

    def toarray(self):
        d = self.__dict__
        shape = tuple(int(x) for x in d.get("_shape", d.get("shape")))
        data = np.asarray(d["data"])
        out = np.zeros(shape, dtype=data.dtype)
        if self.fmt == "coo":
            np.add.at(out, (np.asarray(d["row"]), np.asarray(d["col"])), data)
            return out
        indices, indptr = np.asarray(d["indices"]), np.asarray(d["indptr"])
        major = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
        if self.fmt == "csc":
            np.add.at(out, (indices, major), data)
        else:
            np.add.at(out, (major, indices), data)
        return out


class _CscStub(_SparseStub):
    fmt = "csc"


class _CsrStub(_SparseStub):
    fmt = "csr"


class _CooStub(_SparseStub):
    fmt = "coo"


def _np_global(module, name):
    import importlib
    for mod in (module, module.replace("numpy.core", "numpy._core"), module.replace("numpy._core", "numpy.core")):
        try:
            return getattr(importlib.import_module(mod), name)
        except (ImportError, AttributeError):
            continue
    raise pickle.UnpicklingError(f"numpy global {module}.{name} not found")


class _ModelUnpickler(pickle.Unpickler):
    """Unpickler that can only build numpy arrays, chumpy stubs and sparse-matrix stubs."""

    _NUMPY = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
              ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
              ("numpy", "ndarray"), ("numpy", "dtype")}
    _BUILTINS = {"object", "set", "frozenset", "list", "dict", "tuple", "bytearray", "complex", "slice"}
    _SPARSE = {"csc_matrix": _CscStub, "csr_matrix": _CsrStub, "coo_matrix": _CooStub,
               "csc_array": _CscStub, "csr_array": _CsrStub, "coo_array": _CooStub}

    def find_class(self, module, name):
        if (module, name) in self._NUMPY:
            return _np_global(module, name)
        if module in ("copy_reg", "copyreg") and name == "_reconstructor":
            import copyreg
            return copyreg._reconstructor
        if module in ("__builtin__", "builtins") and name in self._BUILTINS:   # inert containers / scalars
            import builtins
            return getattr(builtins, name)
        if module == "_codecs" and name == "encode":      # how protocol-2 pickles written by Python 3 carry bytes
            import _codecs
            return _codecs.encode
        if module == "chumpy" or module.startswith("chumpy."):
            return _ChumpyStub
        if module.startswith("scipy.sparse") and name in self._SPARSE:
            return self._SPARSE[name]
        raise pickle.UnpicklingError(
            f"refusing to unpickle global {module}.{name}: SMPL model files may only contain numpy arrays, "
            "chumpy arrays and scipy sparse matrices")


def _dense(a):
    if hasattr(a, "toarray"):          # scipy.sparse matrix, or its stub (official pickle)
        a = a.toarray()
    if hasattr(a, "r"):                # chumpy array, or its stub: .r is the numpy value
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
        with open(path, "rb") as f:       # Python-2 pickles: latin1; restricted globals (see module doc)
            d = _ModelUnpickler(f, encoding="latin1").load()
        if not isinstance(d, dict):
            raise ValueError("SMPL model pickle must hold a dict of arrays")
    return from_official_layout(d, num_betas=num_betas)
