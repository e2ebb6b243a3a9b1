"""Importable alias of the `human-3d-reconstruction_b200/` package directory.

The package directory carries the reference's hyphenated name, which Python cannot import
directly; this shim points its ``__path__`` at that directory and runs its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "human-3d-reconstruction_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
