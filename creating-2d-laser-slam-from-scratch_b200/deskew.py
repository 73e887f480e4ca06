"""ctypes harness over the C ABI — lesson5 motion de-skew (LidarUndistortion, lidar_undistortion.cc): the batched
CorrectLaserScan on the device plus the per-scan host preparation (IMU angle integration, odometry increment)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .matcher import check, lib, f64, _d


class DeskewScan(C.Structure):
    """b2s_deskew_scan (include/b200slam.h)."""
    _fields_ = [("time_start", C.c_double), ("time_increment", C.c_double), ("range_min", C.c_float), ("range_max", C.c_float),
                ("use_imu", C.c_int32), ("use_odom", C.c_int32), ("imu_last", C.c_int32), ("reserved", C.c_int32),
                ("odom_start_time", C.c_double), ("odom_end_time", C.c_double), ("odom_incre", C.c_float * 3),
                ("reserved2", C.c_float)]


_bound = False


def _bind():
    global _bound
    L = lib()
    if not _bound:
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        L.b2s_deskew_integrate_imu.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_int, dp, dp, dp, dp]
        L.b2s_deskew_integrate_imu.restype = C.c_int32
        L.b2s_deskew_odom_increment.argtypes = [dp, dp, fp]
        L.b2s_deskew_odom_increment.restype = None
        L.b2s_lidar_undistort.argtypes = [C.c_int, C.c_int, fp, C.c_double, C.c_double, C.POINTER(DeskewScan), dp, dp, dp, dp,
                                          C.c_int, fp, C.c_int, C.c_void_p]
        _bound = True
    return L


def integrate_imu(stamps, angular_velocity, scan_start, scan_end, capacity=2000, fn=None):
    """PruneImuDeque's integration -> (imu_last, imu_time, rot_x, rot_y, rot_z)."""
    fn = fn or _bind().b2s_deskew_integrate_imu
    st, av = f64(stamps), f64(angular_velocity).reshape(-1, 3)
    t, x, y, z = (np.zeros(capacity) for _ in range(4))
    last = fn(len(st), _d(st), _d(av), float(scan_start), float(scan_end), capacity, _d(t), _d(x), _d(y), _d(z))
    return int(last), t, x, y, z


def odom_increment(start_pose, end_pose, fn=None):
    fn = fn or _bind().b2s_deskew_odom_increment
    out = np.zeros(3, np.float32)
    fn(_d(f64(start_pose)), _d(f64(end_pose)), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def undistort(ranges, angle_min, angle_increment, scans, imu_time, rot_x, rot_y, rot_z, device=0, stream=None):
    """ranges [B, n] float32, scans: list of DeskewScan, IMU tables [B, stride] -> corrected cloud [B, n, 3] float32."""
    L = _bind()
    r = np.ascontiguousarray(ranges, np.float32)
    B, n = r.shape
    arr = (DeskewScan * B)(*scans)
    t, x, y, z = (np.ascontiguousarray(a, np.float64).reshape(B, -1) for a in (imu_time, rot_x, rot_y, rot_z))
    out = np.zeros((B, n, 3), np.float32)
    fp = C.POINTER(C.c_float)
    check(L.b2s_lidar_undistort(B, n, r.ctypes.data_as(fp), float(angle_min), float(angle_increment), arr, _d(t), _d(x), _d(y),
                                _d(z), t.shape[1], out.ctypes.data_as(fp), device, C.c_void_p(stream) if stream else None))
    return out
