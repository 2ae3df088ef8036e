"""Import shim: ``import zk_toolkit_b200`` -> the package in ``zk-toolkit_b200/``."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
sys.modules[__name__] = importlib.import_module("zk-toolkit_b200")
