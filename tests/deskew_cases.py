"""Shared synthetic inputs of the lesson5 de-skew tests: a 10 Hz scan of n beams during which the sensor turns and
drives, 200 Hz IMU messages and two odometry messages bracketing the scan."""
import numpy as np


def make_case(pkg, seed, n=1081, rate=(0.6, -0.2, 0.9), vel=(0.8, 0.1, 0.0), dropouts=True):
    DS = pkg.load("deskew")
    rng = np.random.default_rng(seed)
    t0 = 100.0 + rng.uniform(0, 5)
    dt = 0.1 / n
    ranges = rng.uniform(0.5, 25.0, n).astype(np.float32)
    if dropouts:
        ranges[rng.integers(0, n, 20)] = np.inf
        ranges[rng.integers(0, n, 10)] = np.nan
        ranges[rng.integers(0, n, 10)] = 0.01
        ranges[:3] = np.inf  # the first VALID reading is not beam 0
    imu_t = t0 - 0.05 + np.arange(45) * 0.005 + rng.uniform(0, 0.001)
    ang = np.asarray(rate)[None, :] * (1 + 0.1 * rng.standard_normal((45, 3)))
    t_end = t0 + dt * (n - 1)
    o_start, o_end = t0 - 0.013, t_end - 0.004
    start_pose = np.array([1.0, 2.0, 0.0, 0.0, 0.0, 0.3])
    d = o_end - o_start
    end_pose = start_pose + np.array([vel[0] * d, vel[1] * d, vel[2] * d, 0.0, 0.0, rate[2] * d])
    info = DS.DeskewScan()
    info.time_start, info.time_increment = t0, dt
    info.range_min, info.range_max = 0.1, 30.0
    info.use_imu = info.use_odom = 1
    info.odom_start_time, info.odom_end_time = o_start, o_end
    return dict(ranges=ranges, angle_min=-2.35619449, angle_increment=4.71238898 / (n - 1), info=info, imu_stamps=imu_t,
                imu_ang=ang, t_end=t_end, start_pose=start_pose, end_pose=end_pose)
