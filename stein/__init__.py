"""`stein` -- import-compatible alias of `stein_b200` (reference package name:
stein/__init__.py, stein/{samplers,kernels,optimizers,utilities}/__init__.py)."""
import importlib
import sys

import stein_b200

for _sub in ("samplers", "kernels", "optimizers", "utilities", "log_p"):
    _m = importlib.import_module("stein_b200." + _sub)
    sys.modules[__name__ + "." + _sub] = _m
    setattr(sys.modules[__name__], _sub, _m)
__version__ = stein_b200.__version__
