"""The C restatement (oracle/karto_oracle.c) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.  Bit-exact on integers/bytes, 1e-12 on doubles."""
import hashlib
import os

import numpy as np
import pytest

from oracle import port

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = 0.01745329251994329577


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unpack(v):
    return v[0], v[1:4], v[4:].reshape(3, 3)


def close(a, b, tol=1e-12):
    return np.allclose(a, b, rtol=0, atol=tol)


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(G, "karto_small.npz"))


@pytest.mark.parametrize("tag", ["clean", "dropout"])
def test_small_full_dumps(pkg, small, tag):
    abi, synth = pkg.abi, pkg.synth
    laser = synth.Laser(type=2, n_readings=361, min_angle=synth.deg2rad(-90), max_angle=synth.deg2rad(90),
                        angular_resolution=synth.deg2rad(0.5), min_range=0.0, max_range=80.0, range_threshold=6.0)
    g = small
    pm = port.PortMatcher(abi.matcher_params(0.5, 0.05, 0.03, 6.0), abi.laser_from(laser))
    gi = g[f"{tag}_grid_info"]
    assert [pm.g.width, pm.g.height, pm.g.width_step, pm.g.data_size, pm.g.roi_x, pm.g.roi_y, pm.g.roi_w, pm.g.roi_h,
            pm.g.kernel_size] == list(gi)
    assert np.array_equal(pm.kernel(), g[f"{tag}_kernel"])
    pm.set_scan(g[f"{tag}_ranges"], g[f"{tag}_pose"])
    assert np.array_equal(pm.pts, g[f"{tag}_points"], equal_nan=True)
    base_pts = pm.point_readings(g[f"{tag}_base_ranges"], g[f"{tag}_base_pose"])
    assert np.array_equal(pm.find_valid_points(base_pts, pm.sp[:2]), g[f"{tag}_valid_points_base"], equal_nan=True)
    pm.add_scans(g[f"{tag}_base_ranges"], g[f"{tag}_base_pose"])
    assert np.array_equal(pm.grid_off, g[f"{tag}_grid_offset"])
    assert np.array_equal(pm.grid, g[f"{tag}_grid"])
    A, R = 10 * D, 1 * D
    assert np.array_equal(pm.compute_offsets(pm.sp[2], A, R), g[f"{tag}_lut"])
    for pen in (0, 1):
        se = abi.Search(0.25, 0.25, 0.05, 0.05, A, R, pen, 0)
        rc, res = pm.correlate_scan(pm.sp, se, want_sums=True)
        assert rc == 0
        assert np.array_equal(pm.last_sums, g[f"{tag}_sums"])
        r, mean, cov = unpack(g[f"{tag}_corr_pen{pen}"])
        pr = port.result_tuple(res)
        assert close(pr[0], r) and close(pr[1], mean) and close(pr[2], cov)
        sf = abi.Search(0.05, 0.05, 0.05, 0.05, 1 * D, 0.2 * D, pen, 1)
        rc, resf = pm.correlate_scan(pr[1], sf, cov_in=pr[2])
        assert rc == 0
        r, mean, cov = unpack(g[f"{tag}_fine_pen{pen}"])
        prf = port.result_tuple(resf)
        assert close(prf[0], r) and close(prf[1], mean) and close(prf[2], cov)
    rc, res = pm.match_scan(g[f"{tag}_ranges"], g[f"{tag}_pose"], g[f"{tag}_base_ranges"], g[f"{tag}_base_pose"])
    assert rc == 0
    r, mean, cov = unpack(g[f"{tag}_match"])
    pr = port.result_tuple(res)
    assert close(pr[0], r) and close(pr[1], mean) and close(pr[2], cov)
    og = port.occupancy_grid(abi.laser_from(laser), np.stack([g[f"{tag}_base_ranges"], g[f"{tag}_ranges"]]),
                             np.stack([g[f"{tag}_base_pose"], g[f"{tag}_pose"]]), 0.05)
    assert [og["width"], og["height"], og["width_step"]] == list(g[f"{tag}_occ_dims"])
    assert np.array_equal(og["offset"], g[f"{tag}_occ_offset"])
    assert np.array_equal(og["passes"], g[f"{tag}_occ_pass"])
    assert np.array_equal(og["hits"], g[f"{tag}_occ_hit"])
    assert np.array_equal(og["cells"], g[f"{tag}_occ_cells"])


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_cfg1_digests(pkg, seed):
    """BASELINE cfg 1 (1081 beams, 31x31x181): sha256 of grid / LUT / response volume + result doubles."""
    abi, synth = pkg.abi, pkg.synth
    g = np.load(os.path.join(G, "karto_cfg1.npz"))
    t = f"s{seed}"
    # the committed inputs are what the generator produces today (guards against a silent synth change)
    mc = synth.make_match_case(seed, dropout=0.01 if seed == 3 else 0.0)
    assert np.array_equal(mc.ranges, g[f"{t}_ranges"], equal_nan=True)
    pm = port.PortMatcher(abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser()))
    pm.set_scan(g[f"{t}_ranges"], g[f"{t}_pose"])
    pm.add_scans(g[f"{t}_base_ranges"], g[f"{t}_base_pose"])
    assert sha(pm.grid) == str(g[f"{t}_grid_sha"])
    assert int((pm.grid > 0).sum()) == int(g[f"{t}_grid_nonzero"])
    A, R = 22.5 * D, 0.25 * D
    assert sha(pm.compute_offsets(pm.sp[2], A, R)) == str(g[f"{t}_lut_sha"])
    rc, res = pm.correlate_scan(pm.sp, abi.Search(0.75, 0.75, 0.05, 0.05, A, R, 1, 0), want_sums=True)
    assert rc == 0
    assert sha(pm.last_sums) == str(g[f"{t}_sums_sha"])
    assert np.array_equal(pm.last_sums[:, :, 90], g[f"{t}_sums_center_plane"])
    assert np.array_equal(pm.last_sums.max(axis=(0, 1)), g[f"{t}_sums_max_per_angle"])
    r, mean, cov = unpack(g[f"{t}_corr"])
    pr = port.result_tuple(res)
    assert close(pr[0], r) and close(pr[1], mean) and close(pr[2], cov)
    rc, res = pm.match_scan(g[f"{t}_ranges"], g[f"{t}_pose"], g[f"{t}_base_ranges"], g[f"{t}_base_pose"])
    r, mean, cov = unpack(g[f"{t}_match"])
    pr = port.result_tuple(res)
    assert rc == 0 and close(pr[0], r) and close(pr[1], mean) and close(pr[2], cov)


def test_multibase_custom_laser(pkg):
    """12 base scans, sensor offset pose, custom laser (180 beams: the 'no +1' quirk), response expansion on."""
    abi, synth = pkg.abi, pkg.synth
    g = np.load(os.path.join(G, "karto_multibase.npz"))
    laser = synth.Laser(type=0, n_readings=180, min_angle=synth.deg2rad(-90), max_angle=synth.deg2rad(90),
                        angular_resolution=synth.deg2rad(1.0), min_range=0.05, max_range=25.0, range_threshold=8.0,
                        offset_pose=(0.12, -0.03, 0.05))
    al = abi.laser_from(laser)
    pm = port.PortMatcher(abi.matcher_params(0.8, 0.1, 0.1, 8.0, use_response_expansion=1), al)
    assert np.array_equal(pm.sensor_pose(g["odom"]), g["sensor_pose"])
    rc, res = pm.match_scan(g["ranges"][12], g["odom"], g["ranges"][:12], g["poses"][:12])
    assert rc == 0
    assert np.array_equal(pm.grid, g["grid"]) and np.array_equal(pm.grid_off, g["grid_offset"])
    r, mean, cov = unpack(g["match"])
    pr = port.result_tuple(res)
    assert close(pr[0], r) and close(pr[1], mean) and close(pr[2], cov)
    poses = g["poses"].copy()
    poses[12] = g["odom"]
    og = port.occupancy_grid(al, g["ranges"], poses, 0.1)
    assert [og["width"], og["height"], og["width_step"]] == list(g["occ_dims"])
    assert np.array_equal(og["passes"], g["occ_pass"]) and np.array_equal(og["hits"], g["occ_hit"])
    assert np.array_equal(og["cells"], g["occ_cells"])


def test_trace_lines():
    g = np.load(os.path.join(G, "karto_tracelines.npz"))
    pos = 0
    for (x0, y0, x1, y1), n in zip(g["segs"], g["lens"]):
        c = port.trace_line(64, 48, int(x0), int(y0), int(x1), int(y1))
        assert len(c) == n
        assert np.array_equal(c, g["cells"][pos:pos + n])
        pos += n


def test_bad_params(pkg):
    abi = pkg.abi
    L = port.lib()
    import ctypes as C
    g = abi.GridInfo()
    for kw in (dict(resolution=0.0), dict(search_size=-1.0), dict(smear_deviation=-0.1), dict(range_threshold=0.0),
               dict(smear_deviation=0.001), dict(smear_deviation=5.0)):
        args = dict(search_size=1.5, resolution=0.05, smear_deviation=0.03, range_threshold=9.25)
        args.update(kw)
        p = abi.matcher_params(**args)
        assert L.orc_matcher_layout(C.byref(p), C.byref(g)) == abi.B2S_ERR_BAD_PARAMS
