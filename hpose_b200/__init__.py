"""Import name for the package that lives in ``head-pose-estimation-model_b200/`` (hyphens are not
importable).  All code is in that directory; this file only redirects the package search path."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "head-pose-estimation-model_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
