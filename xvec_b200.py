"""Import alias: ``import xvec_b200`` == the package directory ``speaker-recognition-x-vectors_b200/``
(whose name, fixed by the project layout, is not a valid Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("speaker-recognition-x-vectors_b200")
sys.modules[__name__] = _pkg
