"""Host-only tests of the ROS-shaped input adapters (b2s_ros_*): the conversions lesson6's SlamKarto::getLaser / addScan
and lesson4's rosPointCloudToDataContainer apply to the message fields (no ROS, no GPU)."""
import numpy as np
import pytest


def test_karto_laser_and_readings(pkg):
    R, M = pkg.load("rosio"), pkg.load("matcher")
    n = 1080  # a 1081-sample scan spans 1080 increments: Karto's Custom sensor counts Round((max - min) / res), no + 1
    amin, inc = np.float32(-2.35619449), np.float32(0.00436332309619)
    amax = np.float32(amin + np.float32(n) * inc)
    ranges = np.linspace(0.5, 20.0, n).astype(np.float32)
    ranges[7] = np.inf
    ranges[9] = np.nan
    msg = R.make_msg(amin, amax, inc, 0.1, 30.0, ranges)
    laser = R.karto_laser(msg, (0.1, -0.02, 0.3), use_scan_range=12.0)
    assert laser.n_readings == n and laser.min_angle == float(amin) and laser.angular_resolution == float(inc)
    assert laser.range_threshold == 12.0 and list(laser.offset_pose) == [0.1, -0.02, 0.3]
    assert abs(laser.min_range - float(np.float32(0.1))) == 0 and laser.max_range == 30.0
    assert R.karto_laser(msg, use_scan_range=50.0).range_threshold == 30.0      # clipped into [range_min, range_max]
    assert R.karto_laser(msg, use_scan_range=0.01).range_threshold == float(np.float32(0.1))
    r = R.karto_readings(msg)
    assert r.dtype == np.float64 and np.array_equal(r[:7], ranges[:7].astype(np.float64)) and np.isinf(r[7]) and np.isnan(r[9])
    ri = R.karto_readings(msg, inverted=True)
    assert np.array_equal(ri[::-1][:7], r[:7]) and np.isinf(ri[n - 1 - 7])
    with pytest.raises(M.B2SError):  # 1081 ranges against 1080 expected readings: Validate would reject every scan
        R.karto_laser(R.make_msg(amin, amax, inc, 0.1, 30.0, np.ones(n + 1, np.float32)))


def test_hector_points_filters(pkg):
    R = pkg.load("rosio")
    pts = np.array([[1.0, 0.0, 0.0],     # kept
                    [0.05, 0.0, 0.0],    # too close (min_dist 0.2)
                    [-0.5, 0.1, 0.0],    # behind the robot within 0.5 m^2
                    [25.0, 0.0, 0.0],    # beyond use_max_scan_range 20
                    [3.0, 4.0, 2.0],     # above the height window
                    [0.0, 2.0, -0.5],    # kept
                    [40.0, 0.0, 0.0]], np.float32)  # beyond max_dist
    out, origo = R.hector_points(pts, laser_in_base=(0.2, 0.0, 0.3, np.pi / 2), scale_to_map=20.0)
    assert np.allclose(origo, [4.0, 0.0])
    # yaw 90 deg: (x, y) -> (-y, x) + (0.2, 0)
    assert out.shape == (2, 2)
    assert np.allclose(out[0], [(0.2 - 0.0) * 20, (1.0) * 20], atol=1e-4)
    assert np.allclose(out[1], [(0.2 - 2.0) * 20, 0.0], atol=1e-4)
    none, _ = R.hector_points(np.zeros((0, 3), np.float32))
    assert none.shape == (0, 2)
