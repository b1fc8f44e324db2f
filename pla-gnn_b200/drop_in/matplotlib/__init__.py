"""Stand-in for matplotlib, used ONLY when no real matplotlib is installed: the reference imports matplotlib at module level
(train.py:9, utils.py:8) but never plots on the hot path.  With drop_in/ first on sys.path this package would shadow a real
installation, so it first looks for one on the rest of sys.path and, when found, becomes it."""
import importlib.machinery
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.machinery.PathFinder.find_spec(
    "matplotlib", [p for p in sys.path if os.path.abspath(p or os.getcwd()) != _here])
if _spec is not None and _spec.origin and os.path.dirname(os.path.dirname(os.path.abspath(_spec.origin))) != _here:
    _real = importlib.util.module_from_spec(_spec)
    sys.modules["matplotlib"] = _real              # `import matplotlib.pyplot` continues in the real package
    _spec.loader.exec_module(_real)
