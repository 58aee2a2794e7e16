"""CPU-only: libhanabi_b200.so builds for sm_100a, loads, exports every symbol include/hanabi_b200.h declares,
and refuses to create an engine without a GPU (the product has no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hanabi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from hanabi_sad_b200 import _lib

    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libhanabi_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "hanabi_sad_b200/_lib.py has no ctypes signature for %s" % n
    assert L.hb_version() >= 1


def test_config_struct_matches_header_size():
    from hanabi_sad_b200 import _lib

    # hb_config: 9 int32 + pointer (8-aligned) + u64 + 13 x 4 bytes + 7 reserved
    assert ctypes.sizeof(_lib.HbConfig) == 40 + 8 + 8 + 13 * 4 + 7 * 4
    assert ctypes.sizeof(_lib.HbGameInfo) == 4 * (9 + 5 + 5 + 25 + 5 + 25 + 1)


@pytest.mark.skipif(has_gpu(), reason="this box has a GPU")
def test_no_cpu_fallback():
    from hanabi_sad_b200 import Engine, HbError

    with pytest.raises(HbError):
        Engine(4)


def test_bad_config_is_rejected_before_touching_the_device():
    from hanabi_sad_b200 import Engine, HbError

    with pytest.raises(HbError, match="players"):
        Engine(4, players=7)
    with pytest.raises(HbError, match="hand_size"):
        Engine(4, players=2, hand_size=9)
