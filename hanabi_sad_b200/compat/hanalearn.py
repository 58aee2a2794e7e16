"""Drop-in module named `hanalearn`: put this directory on sys.path in place of the reference's `build/`
(pyhanabi/set_path.py:11-19).  Everything is served by libhanabi_b200.so through hanabi_sad_b200.hanalearn."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from hanabi_sad_b200.hanalearn import HanabiEnv, HanabiThreadLoop, HanabiVecEnv  # noqa: F401,E402
from hanabi_sad_b200 import build as _build  # noqa: E402

__file__ = _build.LIB
