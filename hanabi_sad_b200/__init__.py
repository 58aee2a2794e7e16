"""hanabi_sad_b200 -- B200-native Hanabi actor hot path (env step + observation encoding + R2D2 act forward +
prioritized episode replay as sm_100a CUDA kernels behind a C ABI), with Python facades that mirror the
reference's `hanalearn` / `rela` binding classes.  See DESIGN.md and include/hanabi_b200.h."""
from ._lib import lib, HbConfig, HbGameInfo, HbError, check  # noqa: F401
from .engine import Engine  # noqa: F401
