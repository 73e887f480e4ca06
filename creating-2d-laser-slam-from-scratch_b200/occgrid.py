"""ctypes harness over the C ABI — K2c, karto::OccupancyGrid::CreateFromScans (Karto.h:5659-5673)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .matcher import check, lib, f64, _d

_bound = False


def _bind():
    global _bound
    L = lib()
    if _bound:
        return L
    vp = C.c_void_p
    L.b2s_occ_grid_create_from_scans.argtypes = [C.POINTER(abi.Laser), C.c_int, C.POINTER(C.c_double),
                                                 C.POINTER(C.c_double), C.c_double, C.c_int, vp, C.POINTER(vp)]
    L.b2s_occ_grid_info_get.argtypes = [vp, C.POINTER(abi.OccGridInfo)]
    L.b2s_occ_grid_copy.argtypes = [vp, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.b2s_occ_grid_copy_ros.argtypes = [vp, C.POINTER(C.c_int8)]
    L.b2s_occ_grid_last_timing.argtypes = [vp, C.POINTER(C.c_double)]
    L.b2s_occ_grid_destroy.argtypes = [vp]
    L.b2s_occ_grid_destroy.restype = None
    _bound = True
    return L


class OccupancyGrid:
    """Stand-in for karto::OccupancyGrid* as returned by CreateFromScans."""

    def __init__(self, laser: abi.Laser, ranges, poses, resolution: float, device: int = 0, stream=None):
        self.L = _bind()
        p = f64(poses).reshape(-1, 3)
        r = f64(ranges).reshape(len(p), laser.n_readings)
        self.h = C.c_void_p()
        check(self.L.b2s_occ_grid_create_from_scans(C.byref(laser), len(p), _d(r), _d(p), resolution, device,
                                                    C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self.info = abi.OccGridInfo()
        if self.h.value:
            check(self.L.b2s_occ_grid_info_get(self.h, C.byref(self.info)))

    @property
    def is_null(self):
        return not self.h.value

    def arrays(self):
        i = self.info
        n = i.data_size
        cells, pas, hit = np.zeros(n, np.uint8), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        check(self.L.b2s_occ_grid_copy(self.h, cells.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       pas.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       hit.ctypes.data_as(C.POINTER(C.c_uint32))))
        sh = (i.height, i.width_step)
        return dict(width=i.width, height=i.height, width_step=i.width_step, offset=np.array(i.offset[:]),
                    cell_visits=int(i.cell_visits), cells=cells.reshape(sh), passes=pas.reshape(sh),
                    hits=hit.reshape(sh))

    def ros_map(self):
        out = np.zeros((self.info.height, self.info.width), dtype=np.int8)
        check(self.L.b2s_occ_grid_copy_ros(self.h, out.ctypes.data_as(C.POINTER(C.c_int8))))
        return out

    def last_timing(self):
        out = np.zeros(2)
        check(self.L.b2s_occ_grid_last_timing(self.h, _d(out)))
        return dict(raytrace_ms=out[0], threshold_ms=out[1])

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.b2s_occ_grid_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
