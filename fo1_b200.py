"""Importable alias of the ``vlm-fo1_b200`` package (a hyphen is not a valid identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("vlm-fo1_b200")
sys.modules[__name__] = _pkg
