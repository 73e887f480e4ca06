"""ctypes harness over the C ABI — the lesson6 front end (karto::Mapper stand-in, b2s_mapper_*)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .matcher import check, lib


class MapperParams(C.Structure):
    """b2s_mapper_params (values as the Mapper stores them, Mapper.cpp:1448-1653)."""
    _fields_ = [("use_scan_matching", C.c_int32), ("use_scan_barycenter", C.c_int32),
                ("minimum_time_interval", C.c_double), ("minimum_travel_distance", C.c_double),
                ("minimum_travel_heading", C.c_double), ("scan_buffer_size", C.c_int32), ("do_loop_closing", C.c_int32),
                ("scan_buffer_maximum_scan_distance", C.c_double), ("link_match_minimum_response_fine", C.c_double),
                ("link_scan_maximum_distance", C.c_double), ("loop_search_maximum_distance", C.c_double),
                ("loop_match_minimum_chain_size", C.c_int32), ("reserved", C.c_int32),
                ("loop_match_maximum_variance_coarse", C.c_double), ("loop_match_minimum_response_coarse", C.c_double),
                ("loop_match_minimum_response_fine", C.c_double),
                ("sequential", abi.MatcherParams), ("loop", abi.MatcherParams)]


MATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                       C.c_int, C.c_int, C.POINTER(abi.MatchResult))


class ScanSolver(C.Structure):
    """b2s_scan_solver: karto::ScanSolver (Mapper.h:825-891) as a C vtable."""
    ADD_NODE = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.POINTER(C.c_double))
    ADD_CONSTRAINT = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double))
    COMPUTE = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double))
    CLEAR = C.CFUNCTYPE(None, C.c_void_p)
    _fields_ = [("user", C.c_void_p), ("add_node", ADD_NODE), ("add_constraint", ADD_CONSTRAINT),
                ("compute", COMPUTE), ("clear", CLEAR)]


_bound = False


def _bind():
    global _bound
    L = lib()
    if _bound:
        return L
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.b2s_mapper_default_params.argtypes = [C.POINTER(MapperParams), C.c_double]
    L.b2s_mapper_default_params.restype = None
    L.b2s_mapper_create.argtypes = [C.POINTER(MapperParams), C.POINTER(abi.Laser), C.c_int, C.POINTER(vp)]
    L.b2s_mapper_create_with_matcher.argtypes = [C.POINTER(MapperParams), C.POINTER(abi.Laser), MATCH_FN, vp, C.POINTER(vp)]
    L.b2s_mapper_destroy.argtypes = [vp]
    L.b2s_mapper_destroy.restype = None
    L.b2s_mapper_set_scan_solver.argtypes = [vp, C.POINTER(ScanSolver)]
    L.b2s_mapper_process.argtypes = [vp, dp, dp, C.c_double, C.POINTER(C.c_int32), dp]
    L.b2s_mapper_process_sensor.argtypes = [vp, C.c_char_p, dp, dp, C.c_double, C.POINTER(C.c_int32), dp]
    L.b2s_mapper_sensor_count.argtypes = [vp]
    L.b2s_mapper_get_scan_sensors.argtypes = [vp, C.POINTER(C.c_int32)]
    L.b2s_mapper_scan_count.argtypes = [vp]
    L.b2s_mapper_get_poses.argtypes = [vp, dp]
    L.b2s_mapper_edge_count.argtypes = [vp]
    L.b2s_mapper_get_edges.argtypes = [vp, C.POINTER(C.c_int32), dp, dp]
    L.b2s_mapper_stats.argtypes = [vp, dp]
    L.b2s_pose_graph_create.argtypes = [C.POINTER(vp)]
    L.b2s_pose_graph_destroy.argtypes = [vp]
    L.b2s_pose_graph_destroy.restype = None
    L.b2s_pose_graph_as_scan_solver.argtypes = [vp, C.POINTER(ScanSolver)]
    L.b2s_pose_graph_set_iterations.argtypes = [vp, C.c_int, C.c_int]
    L.b2s_pose_graph_stats.argtypes = [vp, dp]
    _bound = True
    return L


def default_params(range_threshold: float, **kw) -> MapperParams:
    """Mapper::InitializeParameters defaults; keyword overrides; `sequential_*` / `loop_*` reach the matcher structs."""
    p = MapperParams()
    _bind().b2s_mapper_default_params(C.byref(p), range_threshold)
    for k, v in kw.items():
        if k.startswith("sequential_"):
            setattr(p.sequential, k[len("sequential_"):], v)
        elif k.startswith("loop_") and hasattr(p.loop, k[len("loop_"):]) and not hasattr(p, k):
            setattr(p.loop, k[len("loop_"):], v)
        elif k.startswith("both_"):
            setattr(p.sequential, k[len("both_"):], v)
            setattr(p.loop, k[len("both_"):], v)
        elif hasattr(p, k):
            setattr(p, k, v)
        else:
            raise AttributeError(k)
    return p


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PoseGraph:
    """The library's dependency-free 2-D pose-graph optimiser (b2s_pose_graph_*), usable as a ScanSolver plug-in."""

    def __init__(self, lm_iterations=40, cg_iterations=400):
        self.L = _bind()
        self.h = C.c_void_p()
        check(self.L.b2s_pose_graph_create(C.byref(self.h)))
        check(self.L.b2s_pose_graph_set_iterations(self.h, lm_iterations, cg_iterations))

    def as_scan_solver(self) -> ScanSolver:
        s = ScanSolver()
        check(self.L.b2s_pose_graph_as_scan_solver(self.h, C.byref(s)))
        s._owner = self  # keep the graph alive as long as the vtable is referenced
        return s

    def stats(self):
        out = np.zeros(5)
        check(self.L.b2s_pose_graph_stats(self.h, _d(out)))
        return dict(nodes=int(out[0]), constraints=int(out[1]), chi2_before=out[2], chi2_after=out[3], lm_steps=int(out[4]))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.b2s_pose_graph_destroy(self.h)
            self.h = C.c_void_p()


class Mapper:
    """Stand-in for karto::Mapper: process(ranges, odometric_pose, time) per LaserScan."""

    def __init__(self, params: MapperParams, laser: abi.Laser, device=0, match_fn=None):
        self.L = _bind()
        self.h = C.c_void_p()
        self._keep = []
        if match_fn is None:
            check(self.L.b2s_mapper_create(C.byref(params), C.byref(laser), device, C.byref(self.h)))
        else:
            cb = MATCH_FN(match_fn)
            self._keep.append(cb)
            check(self.L.b2s_mapper_create_with_matcher(C.byref(params), C.byref(laser), cb, None, C.byref(self.h)))
        self.n = laser.n_readings

    def set_scan_solver(self, solver: ScanSolver):
        self._keep.append(solver)
        check(self.L.b2s_mapper_set_scan_solver(self.h, C.byref(solver)))

    def process(self, ranges, odometric_pose, time=0.0, sensor=None):
        """sensor: None = b2s_mapper_process (the sensor named "laser"); a name = b2s_mapper_process_sensor."""
        r = np.ascontiguousarray(ranges, np.float64)
        assert r.size == self.n
        o, out, ok = np.ascontiguousarray(odometric_pose, np.float64), np.zeros(3), C.c_int32(0)
        if sensor is None:
            check(self.L.b2s_mapper_process(self.h, _d(r), _d(o), float(time), C.byref(ok), _d(out)))
        else:
            check(self.L.b2s_mapper_process_sensor(self.h, str(sensor).encode(), _d(r), _d(o), float(time), C.byref(ok), _d(out)))
        return bool(ok.value), out

    def scan_sensors(self):
        """Per processed scan (unique-id order): the rank of its sensor in name order."""
        out = np.zeros(self.L.b2s_mapper_scan_count(self.h), np.int32)
        if len(out):
            check(self.L.b2s_mapper_get_scan_sensors(self.h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def poses(self):
        out = np.zeros((self.L.b2s_mapper_scan_count(self.h), 3))
        if len(out):
            check(self.L.b2s_mapper_get_poses(self.h, _d(out)))
        return out

    def edges(self):
        n = self.L.b2s_mapper_edge_count(self.h)
        ids, diff, cov = np.zeros((n, 2), np.int32), np.zeros((n, 3)), np.zeros((n, 9))
        if n:
            check(self.L.b2s_mapper_get_edges(self.h, ids.ctypes.data_as(C.POINTER(C.c_int32)), _d(diff), _d(cov)))
        return ids, diff, cov.reshape(n, 3, 3)

    def stats(self):
        out = np.zeros(5)
        check(self.L.b2s_mapper_stats(self.h, _d(out)))
        return dict(match_calls=int(out[0]), batches=int(out[1]), loop_candidates=int(out[2]), loops_closed=int(out[3]),
                    running_scans=int(out[4]))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.b2s_mapper_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
