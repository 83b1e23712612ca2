"""B200-native implementation of the Head-Pose-Estimation-Model hot path.

Import as ``hpose_b200`` (the directory name required by the build contract contains hyphens;
``hpose_b200/__init__.py`` at the repo root maps the import name onto this directory).
"""
from . import _lib  # noqa: F401
from ._lib import HposeError, build  # noqa: F401

__all__ = ["HposeError", "build"]
