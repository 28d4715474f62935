"""Importable alias for the package directory ``candidate-reranking-cir_b200/`` (a hyphen is
not a legal identifier): ``import cir_b200`` returns that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("candidate-reranking-cir_b200")
sys.modules[__name__] = _pkg
