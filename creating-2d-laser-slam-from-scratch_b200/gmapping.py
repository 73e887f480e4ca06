"""ctypes harness over the C ABI — K2b (lesson4 GMapping hit/visit map, gmapping.cc:127-242)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .matcher import check, lib, f64, _d

_bound = False


def _bind():
    global _bound
    L = lib()
    if _bound:
        return L
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.b2s_gmap_create.argtypes = [C.c_double] * 7 + [C.c_int, vp, C.POINTER(vp)]
    L.b2s_gmap_destroy.argtypes = [vp]
    L.b2s_gmap_destroy.restype = None
    L.b2s_gmap_size.argtypes = [vp, C.POINTER(C.c_int32)]
    L.b2s_gmap_compute_map.argtypes = [vp, dp, dp, C.c_int, dp, C.c_double, C.c_double]
    L.b2s_gmap_copy.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.b2s_gmap_copy_ros.argtypes = [vp, C.POINTER(C.c_int8)]
    _bound = True
    return L


class GMap:
    """Stand-in for gmapping::ScanMatcherMap as the lesson4 node uses it (one ComputeMap per scan)."""

    def __init__(self, xmin=-40.0, ymin=-40.0, xmax=40.0, ymax=40.0, delta=0.05, device=0, stream=None):
        self.L = _bind()
        self.bounds = (xmin, ymin, xmax, ymax, delta)
        cx, cy = (xmin + xmax) / 2.0, (ymin + ymax) / 2.0  # gmapping.cc:130-132
        self.h = C.c_void_p()
        check(self.L.b2s_gmap_create(cx, cy, xmin, ymin, xmax, ymax, delta, device,
                                     C.c_void_p(stream) if stream else None, C.byref(self.h)))
        s = (C.c_int32 * 2)()
        check(self.L.b2s_gmap_size(self.h, s))
        self.size_x, self.size_y = s[0], s[1]

    def compute_map(self, ranges, angles, laser_xy=(0.0, 0.0), max_range=30 - 0.01, max_urange=25.0):
        r, a = f64(ranges), f64(angles)
        lp = f64([laser_xy[0], laser_xy[1], 0.0])
        check(self.L.b2s_gmap_compute_map(self.h, _d(r), _d(a), len(r), _d(lp), max_range, max_urange))

    def cells(self):
        c = self.size_x * self.size_y
        n, v = np.zeros(c, np.int32), np.zeros(c, np.int32)
        ax, ay = np.zeros(c, np.float32), np.zeros(c, np.float32)
        ip, fp = C.POINTER(C.c_int32), C.POINTER(C.c_float)
        check(self.L.b2s_gmap_copy(self.h, n.ctypes.data_as(ip), v.ctypes.data_as(ip), ax.ctypes.data_as(fp),
                                   ay.ctypes.data_as(fp)))
        sh = (self.size_y, self.size_x)
        return n.reshape(sh), v.reshape(sh), ax.reshape(sh), ay.reshape(sh)

    def ros_map(self):
        xmin, ymin, xmax, ymax, delta = self.bounds
        w, h = int((xmax - xmin) / delta), int((ymax - ymin) / delta)
        out = np.zeros((h, w), np.int8)
        check(self.L.b2s_gmap_copy_ros(self.h, out.ctypes.data_as(C.POINTER(C.c_int8))))
        return out

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.b2s_gmap_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
