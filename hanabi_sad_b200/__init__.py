"""hanabi_sad_b200 -- B200-native Hanabi actor hot path (env step + observation encoding + R2D2 act forward +
prioritized episode replay as sm_100a CUDA kernels behind a C ABI), with Python facades that mirror the
reference's `hanalearn` / `rela` binding classes.  See DESIGN.md and include/hanabi_b200.h."""
from ._lib import lib, HbConfig, HbGameInfo, HbError, check  # noqa: F401
from .engine import Engine  # noqa: F401


def debug_gemm(A, B, bias=None, split=True, device=0):
    """Diagnostic: C = A @ B.T + bias through the tcgen05 GEMM template (hb_debug_gemm)."""
    import numpy as np

    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    M, K = A.shape
    N = B.shape[0]
    C = np.empty((M, N), np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    check(lib().hb_debug_gemm(device, A.ctypes.data, B.ctypes.data, None if b is None else b.ctypes.data, C.ctypes.data, M, N, K, int(bool(split))))
    return C
