"""Import alias: the package directory is named ``anytime-3d-reconstruction_b200`` (not a valid identifier), so
``import a3d`` loads it through importlib and re-exports it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('anytime-3d-reconstruction_b200')
sys.modules[__name__] = _pkg
