"""Drop-in module named `rela`: put this directory on sys.path in place of the reference's `build/rela`
(pyhanabi/set_path.py:11-19).  Everything is served by libhanabi_b200.so through hanabi_sad_b200.rela."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from hanabi_sad_b200.rela import *  # noqa: F401,F403,E402
from hanabi_sad_b200.rela import aggregate_priority, BatchRunner, Context, FFTransition, R2D2Actor, RNNPrioritizedReplay, RNNTransition, ThreadLoop  # noqa: F401,E402
from hanabi_sad_b200 import build as _build  # noqa: E402

# pyhanabi/create.py:20-21 asserts that the module is a compiled extension; the code behind this module is this library
__file__ = _build.LIB
