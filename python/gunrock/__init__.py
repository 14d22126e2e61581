"""``import gunrock`` -- the reference Python module's name (python/src/gunrock/__init__.py), served by
``gunrock_b200.pygunrock`` over the C ABI.  Put ``<repo>/python`` on ``sys.path`` (or PYTHONPATH)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from gunrock_b200.pygunrock import *  # noqa: E402,F401,F403
from gunrock_b200.pygunrock import __all__  # noqa: E402,F401
