"""Importable alias of the product package.

The package directory carries the reference's repository name
(``simple-implementation-of-structure-from-motion-and-multi-view-stereo-by-python_b200``),
which is not a valid Python identifier; ``import mvs_b200`` resolves its
submodules from that directory.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "simple-implementation-of-structure-from-motion-and-multi-view-stereo-by-python_b200")
__path__.append(_PKG_DIR)

from .context import MvsContext, MvsError  # noqa: E402,F401
