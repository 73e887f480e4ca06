"""ctypes harness over the C ABI — K3 (lesson3): batched PL-ICP (sm_icp call site, plicp_odometry.cc:391)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .matcher import check, lib, f64, _d

_bound = False


def _bind():
    global _bound
    L = lib()
    if not _bound:
        dp = C.POINTER(C.c_double)
        L.b2s_plicp_match.argtypes = [C.POINTER(abi.IcpParams), C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double,
                                      dp, C.c_int, C.c_void_p, C.POINTER(abi.IcpResult)]
        _bound = True
    return L


def match(params: abi.IcpParams, ref_ranges, sens_ranges, theta, range_min, range_max, first_guess, device=0,
          stream=None):
    """-> (x[B,3], valid[B], iterations[B], nvalid[B], error[B])"""
    L = _bind()
    t = f64(theta)
    n = len(t)
    r, s = f64(ref_ranges).reshape(-1, n), f64(sens_ranges).reshape(-1, n)
    g = f64(first_guess).reshape(-1, 3)
    B = len(r)
    res = (abi.IcpResult * B)()
    check(L.b2s_plicp_match(C.byref(params), B, n, _d(r), _d(s), _d(t), range_min, range_max, _d(g), device,
                            C.c_void_p(stream) if stream else None, res))
    raw = np.frombuffer(res, dtype=np.dtype([("x", "<f8", 3), ("error", "<f8"), ("valid", "<i4"), ("iterations", "<i4"),
                                             ("nvalid", "<i4"), ("reserved", "<i4")]), count=B)
    return raw["x"].copy(), raw["valid"].copy(), raw["iterations"].copy(), raw["nvalid"].copy(), raw["error"].copy()
