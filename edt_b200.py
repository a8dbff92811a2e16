"""`import edt_b200 as edt` -- importable alias of the package directory
`euclidean-distance-transform-3d_b200/` (whose name is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
  sys.path.insert(0, _root)
_pkg = importlib.import_module("euclidean-distance-transform-3d_b200")
sys.modules[__name__] = _pkg
