"""Alias so that ``import sixdgs_b200`` works: the package directory is ``6dgs_b200`` (the name the
project brief fixes), which is not a valid Python identifier."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("6dgs_b200")
sys.modules[__name__] = _pkg
