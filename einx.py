"""`import einx` -> the package in ./ei-nexus_official_b200 (its directory name is not a Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("ei-nexus_official_b200")
