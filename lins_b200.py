"""Import alias: the package directory name (`lins---lidar-inertial-slam_b200`) is not a valid Python
identifier, so `import lins_b200` resolves to it through importlib."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("lins---lidar-inertial-slam_b200")
sys.modules[__name__] = _pkg
