"""GPU parity tests for K1 (Karto correlative scan matcher) — run on the B200 box with `-m gpu`.

Every call goes through the C ABI (libb200slam.so via ctypes).  The checker is the CPU restatement
oracle/karto_oracle.c (itself pinned to the unmodified reference, tests/test_oracle_*.py) and the committed
golden vectors produced by the reference build.  Gates:
  grid bytes, lookup tables, integer response sums : bit-exact
  response / pose / covariance                     : |delta| <= 1e-4 required; we assert 1e-9
"""
import os

import numpy as np
import pytest

from oracle import port

pytestmark = pytest.mark.gpu
D = 0.01745329251994329577
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9  # the contract is 1e-4; the implementation is far tighter


@pytest.fixture(scope="module")
def M(pkg):
    m = pkg.load("matcher")
    assert m.device_count() > 0, "no CUDA device: the product path has no fallback"
    return m


def port_case(abi, params, laser, ranges, pose, base_ranges, base_poses):
    pm = port.PortMatcher(params, laser)
    pm.set_scan(ranges, pose)
    pm.add_scans(base_ranges, base_poses)
    return pm


def assert_result(gpu, b, res, tol=TOL):
    resp, pose, cov, status, ties = gpu
    pr = port.result_tuple(res)
    assert status[b] == res.status == 0
    assert abs(resp[b] - pr[0]) <= tol
    assert np.allclose(pose[b], pr[1], rtol=0, atol=tol)
    assert np.allclose(cov[b], pr[2], rtol=0, atol=tol)
    assert ties[b] == res.tie_count


def make_batch(synth, seeds, laser=None, **kw):
    cases = [synth.make_match_case(s, laser or synth.Laser(), dropout=0.02 if (s % 4 == 3) else 0.0, **kw)
             for s in seeds]
    return (cases, np.stack([c.ranges for c in cases]), np.stack([c.odom_pose for c in cases]),
            np.stack([c.base_ranges for c in cases])[:, None, :], np.stack([c.base_pose for c in cases])[:, None, :])


@pytest.mark.parametrize("kernel", [2, 3, 1])  # 2 = window kernel (hot path), 3 = same without empty-window dropping, 1 = generic
def test_cfg1_correlate_parity(pkg, M, kernel):
    """BASELINE cfg 1/2 shape: 1081 beams, 31x31x181 window, 0.05 m grid; 6 matches incl. NaN/inf dropouts."""
    abi, synth = pkg.abi, pkg.synth
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    cases, ranges, poses, bran, bpos = make_batch(synth, range(100, 106))
    B = len(cases)
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1)
    m.set_kernel(kernel)
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    A, R = 22.5 * D, 0.25 * D
    for pen in (1, 0):
        se = abi.Search(0.75, 0.75, 0.05, 0.05, A, R, pen, 0)
        sensor = np.stack([port.PortMatcher(params, laser).sensor_pose(p) for p in poses])
        gpu = m.correlate_scan(sensor, se)
        assert m.last_timing()["path"] == min(kernel, 2)
        st = m.last_stats()
        assert (st["empty_window_frac"] > 0.05) if kernel == 2 else (st["empty_window_frac"] == 0.0)
        for b in range(B):
            pm = port_case(abi, params, laser, ranges[b], poses[b], bran[b], bpos[b])
            if pen:
                # doubles: device cos/sin differ from glibc in the last ulp; integers derived from them are compared exactly below
                assert np.allclose(m.point_readings(b), pm.pts, rtol=0, atol=1e-12, equal_nan=True)
                g, off = m.grid(b)
                assert np.array_equal(off, pm.grid_off)
                assert np.array_equal(g, pm.grid)
                assert np.array_equal(m.compute_offsets(b, pm.sp[2], A, R), pm.compute_offsets(pm.sp[2], A, R))
            rc, res = pm.correlate_scan(pm.sp, se, want_sums=True)
            assert rc == 0
            assert np.array_equal(m.response_sums(b, (31, 31, 181)), pm.last_sums)
            assert_result(gpu, b, res)
    m.close()


def test_cfg1_golden(pkg, M):
    """The same call against numbers produced by the UNMODIFIED reference (tests/golden/karto_cfg1.npz)."""
    abi, synth = pkg.abi, pkg.synth
    g = np.load(os.path.join(G, "karto_cfg1.npz"))
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    B = 4
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1)
    ranges = np.stack([g[f"s{i}_ranges"] for i in range(B)])
    poses = np.stack([g[f"s{i}_pose"] for i in range(B)])
    m.set_scans(ranges, poses)
    m.add_scans(np.stack([g[f"s{i}_base_ranges"] for i in range(B)])[:, None, :],
                np.stack([g[f"s{i}_base_pose"] for i in range(B)])[:, None, :])
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    A, R = 22.5 * D, 0.25 * D
    resp, pose, cov, status, ties = m.correlate_scan(poses, abi.Search(0.75, 0.75, 0.05, 0.05, A, R, 1, 0))
    for b in range(B):
        assert sha(m.grid(b)[0]) == str(g[f"s{b}_grid_sha"])
        assert sha(m.compute_offsets(b, poses[b, 2], A, R)) == str(g[f"s{b}_lut_sha"])
        assert sha(m.response_sums(b, (31, 31, 181))) == str(g[f"s{b}_sums_sha"])
        v = g[f"s{b}_corr"]
        assert status[b] == 0 and abs(resp[b] - v[0]) <= TOL
        assert np.allclose(pose[b], v[1:4], rtol=0, atol=TOL) and np.allclose(cov[b].ravel(), v[4:], rtol=0, atol=TOL)
    resp, pose, cov, status, ties = m.match_scan()
    for b in range(B):
        v = g[f"s{b}_match"]
        assert status[b] == 0 and abs(resp[b] - v[0]) <= TOL
        assert np.allclose(pose[b], v[1:4], rtol=0, atol=TOL) and np.allclose(cov[b].ravel(), v[4:], rtol=0, atol=TOL)
    m.close()


def test_match_scan_parity(pkg, M):
    """ScanMatcher::MatchScan two-stage driver (coarse stride-2 sweep + fine 3x3 sweep) vs the restatement."""
    abi, synth = pkg.abi, pkg.synth
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    cases, ranges, poses, bran, bpos = make_batch(synth, range(200, 208))
    B = len(cases)
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1)
    for pen, refine in ((True, True), (False, True), (True, False)):
        gpu = m.match_scan_host(ranges, poses, bran, bpos, pen, refine)
        for b in range(B):
            pm = port.PortMatcher(params, laser)
            rc, res = pm.match_scan(ranges[b], poses[b], bran[b], bpos[b], pen, refine)
            assert rc == 0
            assert_result(gpu, b, res)
    m.close()


def test_multibase_custom_laser_golden(pkg, M):
    """12 base scans per match, sensor offset pose, custom 180-beam laser, response expansion enabled; golden from
    the reference build."""
    abi, synth = pkg.abi, pkg.synth
    g = np.load(os.path.join(G, "karto_multibase.npz"))
    laser = synth.Laser(type=0, n_readings=180, min_angle=synth.deg2rad(-90), max_angle=synth.deg2rad(90),
                        angular_resolution=synth.deg2rad(1.0), min_range=0.05, max_range=25.0, range_threshold=8.0,
                        offset_pose=(0.12, -0.03, 0.05))
    params = abi.matcher_params(0.8, 0.1, 0.1, 8.0, use_response_expansion=1)
    m = M.ScanMatcher(params, abi.laser_from(laser), max_batch=2, max_base_scans=12)
    # two identical matches in the batch: results must also be identical to each other
    ranges = np.stack([g["ranges"][12]] * 2)
    poses = np.stack([g["odom"]] * 2)
    resp, pose, cov, status, ties = m.match_scan_host(ranges, poses, np.stack([g["ranges"][:12]] * 2),
                                                      np.stack([g["poses"][:12]] * 2))
    for b in range(2):
        grid, off = m.grid(b)
        assert np.array_equal(grid, g["grid"]) and np.array_equal(off, g["grid_offset"])
        v = g["match"]
        assert status[b] == 0 and abs(resp[b] - v[0]) <= TOL
        assert np.allclose(pose[b], v[1:4], rtol=0, atol=TOL) and np.allclose(cov[b].ravel(), v[4:], rtol=0, atol=TOL)
    m.close()


@pytest.mark.parametrize("res,smear,search,rt", [(0.025, 0.03, 0.5, 6.0), (0.1, 0.3, 1.0, 6.0), (0.05, 0.1, 0.4, 6.0),
                                                 (0.025, 0.03, 1.5, 9.25)])
def test_other_geometries(pkg, M, res, smear, search, rt):
    """cfg-4 style 0.025 m grids (the last one is 652 KB: swept by the window kernel in row bands), the
    outdoor yaml's 0.1 m / smear 0.3 (13x13 smear kernel), and a mid case; MatchScan + a direct full-window
    CorrelateScan with the integer volume compared bit-exactly."""
    abi, synth = pkg.abi, pkg.synth
    laser_s = synth.Laser(range_threshold=rt)
    params, laser = abi.matcher_params(search, res, smear, rt), abi.laser_from(laser_s)
    cases, ranges, poses, bran, bpos = make_batch(synth, range(300, 303), laser_s, max_xy=0.1, max_th_deg=4)
    B = len(cases)
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1)
    gpu = m.match_scan_host(ranges, poses, bran, bpos)
    pms = []
    for b in range(B):
        pm = port.PortMatcher(params, laser)
        rc, resm = pm.match_scan(ranges[b], poses[b], bran[b], bpos[b])
        assert rc == 0
        assert np.array_equal(m.grid(b)[0], pm.grid)
        assert_result(gpu, b, resm)
        pm.set_scan(ranges[b], poses[b])
        pms.append(pm)
    half = 0.5 * (m.g.search_side - 1) * res
    se = abi.Search(half, half, res, res, 10 * D, 1 * D, 0, 0)
    gpu = m.correlate_scan(poses, se)
    side = m.g.search_side
    for b in range(B):
        rc, r = pms[b].correlate_scan(pms[b].sp, se, want_sums=True)
        assert rc == 0
        assert np.array_equal(m.response_sums(b, (side, side, 21)), pms[b].last_sums)
        assert_result(gpu, b, r)
    m.close()


def test_window_tiles_61(pkg, M):
    """A 61x61 search window (cfg-4's largest side) on a grid that still fits in shared memory: exercises the
    2x2 tiling of the window kernel with partial tiles; both kernels must give the same integers."""
    abi, synth = pkg.abi, pkg.synth
    laser_s = synth.Laser(range_threshold=3.0)
    params, laser = abi.matcher_params(3.0, 0.05, 0.03, 3.0), abi.laser_from(laser_s)
    cases, ranges, poses, bran, bpos = make_batch(synth, range(400, 402), laser_s)
    m = M.ScanMatcher(params, laser, max_batch=2, max_base_scans=1)
    assert m.g.search_side == 61
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    se = abi.Search(1.5, 1.5, 0.05, 0.05, 5 * D, 1 * D, 1, 0)
    out = {}
    for kernel in (2, 1):
        m.set_kernel(kernel)
        gpu = m.correlate_scan(poses, se)
        assert m.last_timing()["path"] == kernel
        out[kernel] = [m.response_sums(b, (61, 61, 11)) for b in range(2)]
    for b in range(2):
        pm = port_case(abi, params, laser, ranges[b], poses[b], bran[b], bpos[b])
        rc, r = pm.correlate_scan(pm.sp, se, want_sums=True)
        assert np.array_equal(out[2][b], pm.last_sums) and np.array_equal(out[1][b], pm.last_sums)
        assert_result(gpu, b, r)
    m.close()


def test_empty_grid_all_candidates_tie(pkg, M):
    """No base scans: all responses 0, every candidate ties (tree-reduction path of the tie average), covariance
    takes the MAX_VARIANCE branch (Mapper.cpp:545-552)."""
    abi, synth = pkg.abi, pkg.synth
    laser_s = synth.Laser(range_threshold=6.0)
    params, laser = abi.matcher_params(0.5, 0.05, 0.03, 6.0), abi.laser_from(laser_s)
    mc = synth.make_match_case(40, laser_s)
    m = M.ScanMatcher(params, laser, max_batch=1, max_base_scans=1)
    gpu = m.match_scan_host(mc.ranges[None], mc.odom_pose[None], np.zeros((1, 0, 1081)), np.zeros((1, 0, 3)))
    pm = port.PortMatcher(params, laser)
    rc, res = pm.match_scan(mc.ranges, mc.odom_pose, np.zeros((0, 1081)), np.zeros((0, 3)))
    assert_result(gpu, 0, res)
    m.close()


def test_status_codes(pkg, M):
    abi, synth = pkg.abi, pkg.synth
    laser = abi.laser_from(synth.Laser())
    for kw in (dict(resolution=0.0), dict(search_size=-1.0), dict(smear_deviation=-0.1), dict(range_threshold=0.0),
               dict(smear_deviation=0.001), dict(smear_deviation=5.0)):
        args = dict(search_size=1.5, resolution=0.05, smear_deviation=0.03, range_threshold=9.25)
        args.update(kw)
        with pytest.raises(M.B2SError) as e:
            M.ScanMatcher(abi.matcher_params(**args), laser, 1)
        assert e.value.status == abi.B2S_ERR_BAD_PARAMS
    params = abi.matcher_params(1.5, 0.05, 0.03, 9.25)
    m = M.ScanMatcher(params, laser, max_batch=2, max_base_scans=1)
    with pytest.raises(M.B2SError) as e:
        m.match_scan()
    assert e.value.status == abi.B2S_ERR_BAD_STATE
    cases, ranges, poses, bran, bpos = make_batch(synth, range(2))
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    # a search centre 50 m away: the candidate lattice leaves the grid -> karto::Exception in the reference
    centers = poses.copy()
    centers[1, 0] += 50.0
    resp, pose, cov, status, ties = m.correlate_scan(centers, abi.Search(0.75, 0.75, 0.05, 0.05, 5 * D, 1 * D, 1, 0))
    assert status[0] == abi.B2S_OK and status[1] == abi.B2S_ERR_OUT_OF_RANGE
    pm = port_case(abi, params, laser, ranges[1], poses[1], bran[1], bpos[1])
    rc, _ = pm.correlate_scan(centers[1], abi.Search(0.75, 0.75, 0.05, 0.05, 5 * D, 1 * D, 1, 0))
    assert rc == abi.B2S_ERR_OUT_OF_RANGE
    with pytest.raises(M.B2SError) as e:
        m.set_scans(np.zeros((3, 1081)), np.zeros((3, 3)))
    assert e.value.status == abi.B2S_ERR_TOO_LARGE
    m.close()


def window_checksum(grid, lut, base00, step, nx, ny, data_size):
    """Size-independent property: for every angle k, SUM_{x,y} sums[y,x,k] = SUM_i (sum of the ny x nx flat-index
    window of the grid at base00 + lut[k,i]).  Evaluated with a prefix sum of the flat grid (zero outside
    [0, data_size)), so it shares no code path with either kernel."""
    pad = ny * step + nx + 8
    P = np.concatenate([[0], np.cumsum(grid.astype(np.int64))])

    def pref(i):  # prefix sum with clamping = zeros outside the array
        return P[np.clip(i, 0, data_size)]

    valid = lut != np.iinfo(np.int32).max
    o = base00 + lut.astype(np.int64)
    tot = np.zeros(lut.shape[0], dtype=np.int64)
    for r in range(ny):
        a = o + r * step
        tot += np.where(valid, pref(a + nx) - pref(a), 0).sum(axis=1)
    return tot


def test_full_batch_1024_properties(pkg, M):
    """BASELINE cfg 2 at full size (B = 1024): every match's integer volume passes the window-sum checksum,
    matches fed identical inputs give identical outputs, and a sample is compared with the restatement."""
    abi, synth = pkg.abi, pkg.synth
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    uniq = 32
    cases, r0, p0, br0, bp0 = make_batch(synth, range(500, 500 + uniq))
    reps = 1024 // uniq
    ranges, poses = np.tile(r0, (reps, 1)), np.tile(p0, (reps, 1))
    bran, bpos = np.tile(br0, (reps, 1, 1)), np.tile(bp0, (reps, 1, 1))
    m = M.ScanMatcher(params, laser, max_batch=1024, max_base_scans=1)
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    A, R = 22.5 * D, 0.25 * D
    se = abi.Search(0.75, 0.75, 0.05, 0.05, A, R, 1, 0)
    resp, pose, cov, status, ties = m.correlate_scan(poses, se)
    assert m.last_timing()["path"] == 2
    assert (status == 0).all()
    # identical inputs -> identical outputs, wherever in the batch (and on whichever SM) they ran
    for k in range(1, reps):
        sl = slice(k * uniq, (k + 1) * uniq)
        assert np.array_equal(resp[sl], resp[:uniq]) and np.array_equal(pose[sl], pose[:uniq])
        assert np.array_equal(cov[sl], cov[:uniq])
    g = m.g
    for b in list(range(0, uniq, 4)) + [1023, 517]:
        sums = m.response_sums(b, (31, 31, 181))
        grid, off = m.grid(b)
        lut = m.compute_offsets(b, poses[b, 2], A, R)
        gx = int(pkg.abi.karto_round(((poses[b, 0] - 0.75) - off[0]) * (1.0 / 0.05))) + g.roi_x
        gy = int(pkg.abi.karto_round(((poses[b, 1] - 0.75) - off[1]) * (1.0 / 0.05))) + g.roi_y
        chk = window_checksum(grid, lut, gx + gy * g.width_step, g.width_step, 31, 31, g.data_size)
        assert np.array_equal(sums.sum(axis=(0, 1), dtype=np.int64), chk)
        assert np.array_equal(sums, m.response_sums(b % uniq, (31, 31, 181)))
    for b in (0, 7, 1000):
        pm = port_case(abi, params, laser, ranges[b], poses[b], bran[b], bpos[b])
        rc, res = pm.correlate_scan(pm.sp, se, want_sums=True)
        assert np.array_equal(m.response_sums(b, (31, 31, 181)), pm.last_sums)
        assert_result((resp, pose, cov, status, ties), b, res)
    # pose error vs ground truth is bounded by the search lattice (sanity, not parity)
    true = np.stack([c.true_pose for c in cases])
    err = np.abs(pose[:uniq, :2] - true[:, :2]).max()
    assert err < 0.1, err
    m.close()


def test_randomised_configurations(pkg, M):
    """Seeded fuzz over matcher geometry (resolution, smear, search size, beam count, sensor offset, window shape,
    batch size): every integer volume bit-exact, every result within TOL, on whichever kernel the handle picks."""
    abi, synth = pkg.abi, pkg.synth
    rng = np.random.default_rng(2024)
    kinds = {1: 0, 2: 0}
    for trial in range(14):
        res = float(rng.choice([0.03, 0.04, 0.05, 0.08, 0.1]))
        side_cells = int(rng.choice([7, 11, 15, 21, 33, 41]))
        search = (side_cells - 1) * res
        smear = float(rng.uniform(0.6, 3.0)) * res
        rt = float(rng.uniform(3.0, 7.0))
        n_beams = int(rng.choice([181, 360, 721, 1081]))
        span = float(rng.choice([180.0, 240.0, 270.0]))
        laser_s = synth.Laser(type=0, n_readings=n_beams, min_angle=synth.deg2rad(-span / 2), max_angle=synth.deg2rad(span / 2),
                              angular_resolution=synth.deg2rad(span / n_beams), min_range=0.05, max_range=25.0,
                              range_threshold=rt, offset_pose=(float(rng.uniform(-0.2, 0.2)), float(rng.uniform(-0.2, 0.2)),
                                                               float(rng.uniform(-0.3, 0.3))))
        params = abi.matcher_params(search, res, smear, rt, use_response_expansion=int(rng.integers(0, 2)))
        laser = abi.laser_from(laser_s)
        B, nb = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        world, poses, ranges = synth.make_trajectory(1000 + trial, nb + B, laser_s, step_xy=0.1, step_th_deg=3)
        if trial % 3 == 0:
            ranges[:, ::17] = np.nan
            ranges[:, 5::29] = np.inf
        cur_r = ranges[nb:]
        cur_p = poses[nb:] + rng.uniform(-0.05, 0.05, size=(B, 3))
        bran = np.stack([ranges[:nb]] * B)
        bpos = np.stack([poses[:nb]] * B)
        m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=nb)
        gpu = m.match_scan_host(cur_r, cur_p, bran, bpos, bool(trial % 2), True)
        pms = []
        for b in range(B):
            pm = port.PortMatcher(params, laser)
            rc, r0 = pm.match_scan(cur_r[b], cur_p[b], bran[b], bpos[b], bool(trial % 2), True)
            assert rc == 0
            assert np.array_equal(m.grid(b)[0], pm.grid), trial
            assert_result(gpu, b, r0)
            pm.set_scan(cur_r[b], cur_p[b])
            pms.append(pm)
        # a direct full-resolution sweep with an odd angle count
        half = 0.5 * (m.g.search_side - 1) * res
        na_off, na_res = float(rng.uniform(2, 12)) * D, float(rng.choice([0.5, 1.0, 1.5])) * D
        se = abi.Search(half, half, res, res, na_off, na_res, int(rng.integers(0, 2)), 0)
        centers = np.stack([pm.sp for pm in pms])
        gpu = m.correlate_scan(centers, se)
        kinds[m.last_timing()["path"]] += 1
        na = abi.n_steps(na_off, na_res)
        for b in range(B):
            rc, r = pms[b].correlate_scan(pms[b].sp, se, want_sums=True)
            assert rc == 0
            assert np.array_equal(m.response_sums(b, (m.g.search_side, m.g.search_side, na)), pms[b].last_sums), trial
            assert_result(gpu, b, r)
        m.close()
    assert kinds[2] >= 8  # most of these grids fit in shared memory and use the window kernel


@pytest.mark.parametrize("res,search,rt,na_deg,far", [(0.05, 1.5, 12.0, 10, 0), (0.025, 1.5, 9.25, 5, 0), (0.025, 0.75, 9.25, 8, 0),
                                                      (0.025, 1.5, 9.25, 4, 1), (0.05, 3.1, 12.0, 4, 1)])
def test_banded_window_kernel_large_grids(pkg, M, res, search, rt, na_deg, far):
    """Grids larger than shared memory (266 KB for a 12 m range threshold @0.05 m; 652 KB for cfg 4's 0.025 m grid,
    with a 61x61 or 31x31 window) are swept by the window kernel in row bands whose partial sums are combined with
    RED.ADD: bit-exact vs the restatement and vs the generic kernel, empty-window dropping on and off."""
    abi, synth = pkg.abi, pkg.synth
    laser_s = synth.Laser(range_threshold=rt)
    params, laser = abi.matcher_params(search, res, 0.03, rt), abi.laser_from(laser_s)
    cases, ranges, poses, bran, bpos = make_batch(synth, range(600, 603), laser_s, max_xy=0.2, max_th_deg=5)
    if far:
        # readings far beyond the range threshold: their window origins lie outside the grid on every side (flat index
        # below 0 / above the data, row-wrapped), some close enough that the upper row tiles still reach the grid
        ranges = ranges.copy()
        n = ranges.shape[1]
        ranges[:, ::5] = rt + 0.02 * (np.arange(0, n, 5) % 97)
        ranges[:, 3::11] = 14.0 + (np.arange(3, n, 11) % 13)
    B = len(cases)
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1)
    assert m.g.data_size > 230_000
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    side = m.g.search_side
    half = 0.5 * (side - 1) * res
    se = abi.Search(half, half, res, res, na_deg * D, 1 * D, 1, 0)
    na = abi.n_steps(na_deg * D, 1 * D)
    vols = {}
    for kernel in (2, 3, 1):
        m.set_kernel(kernel)
        gpu = m.correlate_scan(poses, se)
        assert m.last_timing()["path"] == min(kernel, 2)
        vols[kernel] = [m.response_sums(b, (side, side, na)) for b in range(B)]
    for b in range(B):
        pm = port_case(abi, params, laser, ranges[b], poses[b], bran[b], bpos[b])
        rc, r = pm.correlate_scan(pm.sp, se, want_sums=True)
        assert rc == 0
        for kernel in (2, 3, 1):
            assert np.array_equal(vols[kernel][b], pm.last_sums), (kernel, b)
        assert_result(gpu, b, r)
    m.close()


@pytest.mark.parametrize("res,search,rt", [(0.05, 1.5, 9.25), (0.05, 3.7, 6.0), (0.1, 15.0, 5.0), (0.1, 15.0, 50.0)])
def test_stride2_coarse_lattice(pkg, M, res, search, rt):
    """The coarse stage of MatchScan searches every other cell (Mapper.cpp:233-234).  The window kernel handles that
    lattice with 16 candidates per 32-byte tile row: 16x16 (one tile), 38x38 (3x2 tiles) and the outdoor yaml's
    76x76 loop-closure window (151-cell side @0.1 m, smear 0.3) on a small grid (rt = 5 m) and on cfg 5's real 1.36 MB
    grid (rt = 50 m: 1165 x 1168 bytes, swept as (row band, row tile) units).
    Integer volumes bit-exact vs the restatement on every kernel."""
    abi, synth = pkg.abi, pkg.synth
    laser_s = synth.Laser(range_threshold=rt)
    smear = 0.3 if res == 0.1 else 0.03
    params, laser = abi.matcher_params(search, res, smear, rt), abi.laser_from(laser_s)
    cases, ranges, poses, bran, bpos = make_batch(synth, range(700, 702), laser_s, max_xy=0.2, max_th_deg=5)
    B = len(cases)
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1)
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    side = m.g.search_side
    half = 0.5 * (side - 1) * res
    se = abi.Search(half, half, 2 * res, 2 * res, 20 * D, 2 * D, 1, 0)
    nxy, na = abi.n_steps(half, 2 * res), abi.n_steps(20 * D, 2 * D)
    vols = {}
    for kernel in (2, 3, 1):
        m.set_kernel(kernel)
        gpu = m.correlate_scan(poses, se)
        assert m.last_timing()["path"] == min(kernel, 2)
        vols[kernel] = [m.response_sums(b, (nxy, nxy, na)) for b in range(B)]
    for b in range(B):
        pm = port_case(abi, params, laser, ranges[b], poses[b], bran[b], bpos[b])
        rc, r = pm.correlate_scan(pm.sp, se, want_sums=True)
        assert rc == 0
        for kernel in (2, 3, 1):
            assert np.array_equal(vols[kernel][b], pm.last_sums), (kernel, b)
        assert_result(gpu, b, r)
    m.close()


@pytest.mark.parametrize("world,kernel", [(2, 2), (3, 2), (2, 1), (5, 2)])
def test_angle_split_sweep_matches_whole_sweep(pkg, M, world, kernel):
    """the angle range of ONE sweep split over `world` ranks, emulated on one handle (each rank's
    begin/ties phases run in turn; numpy max/sum stand for the all-reduces).  Result must equal the whole sweep and
    the CPU restatement; empty-grid ties (every candidate ties) exercise the SUM path."""
    abi, synth, par = pkg.abi, pkg.synth, pkg.load("parallel")
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    cases, ranges, poses, bran, bpos = make_batch(synth, range(300, 304))
    B = len(cases)
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1)
    m.set_kernel(kernel)
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    A, R = 22.5 * D, 0.25 * D
    na = abi.n_steps(A, R)
    sensor = np.stack([port.PortMatcher(params, laser).sensor_pose(p) for p in poses])
    for pen in (1, 0):
        se = abi.Search(0.75, 0.75, 0.05, 0.05, A, R, pen, 0)
        whole = m.correlate_scan(sensor, se)
        bounds = [par.shard_bounds(na, world, r) for r in range(world)]
        parts = [m.split_begin(sensor, se, lo, hi - lo) for lo, hi in bounds]
        best = np.max([p[0] for p in parts], axis=0)
        probs = np.max([p[1] for p in parts], axis=0)
        assert all((p[2] == 0).all() for p in parts)
        ties = np.zeros((B, 5))
        for lo, hi in bounds:
            m.split_begin(sensor, se, lo, hi - lo)  # this "rank"'s sweep volume is resident again
            ties += m.split_ties(best)
        got = m.split_finish(best, ties, probs)
        for a, b_ in zip(got, whole):
            assert np.allclose(a, b_, rtol=0, atol=TOL), (world, pen)
        for b in range(B):
            pm = port_case(abi, params, laser, ranges[b], poses[b], bran[b], bpos[b])
            rc, res = pm.correlate_scan(pm.sp, se)
            assert rc == 0
            assert_result(got, b, res)
    # world > number of angles: some ranks sweep nothing
    se = abi.Search(0.75, 0.75, 0.05, 0.05, 1.0 * D, 1.0 * D, 1, 0)
    na = abi.n_steps(1.0 * D, 1.0 * D)
    assert na == 3
    whole = m.correlate_scan(sensor, se)
    bounds = [par.shard_bounds(na, 5, r) for r in range(5)]
    parts = [m.split_begin(sensor, se, lo, hi - lo) for lo, hi in bounds]
    best = np.max([p[0] for p in parts], axis=0)
    probs = np.max([p[1] for p in parts], axis=0)
    ties = np.zeros((B, 5))
    for lo, hi in bounds:
        m.split_begin(sensor, se, lo, hi - lo)
        ties += m.split_ties(best)
    got = m.split_finish(best, ties, probs)
    for a, b_ in zip(got, whole):
        assert np.allclose(a, b_, rtol=0, atol=TOL)
    # the single-process form of the orchestration helper (world 1, no process group)
    got1 = par.correlate_scan_angle_split(m, sensor, se, na)
    for a, b_ in zip(got1, whole):
        assert np.allclose(a, b_, rtol=0, atol=TOL)
    m.close()


def test_correlate_begin_end_pipelined_over_two_handles(pkg):
    """b2s_matcher_correlate_scan_begin/_end on two handles, interleaved as a 2-deep pipeline: identical results to the
    synchronous call, and _end without _begin is a BAD_STATE error."""
    abi, synth, M = pkg.abi, pkg.synth, pkg.load("matcher")
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    se = abi.Search(0.75, 0.75, 0.05, 0.05, 22.5 * D, 0.25 * D, 1, 0)
    sets = []
    for k in range(4):
        cases = [synth.make_match_case(7000 + 10 * k + i) for i in range(6)]
        sets.append((np.stack([c.ranges for c in cases]), np.stack([c.odom_pose for c in cases]),
                     np.stack([c.base_ranges for c in cases])[:, None, :], np.stack([c.base_pose for c in cases])[:, None, :]))
    ms = [M.ScanMatcher(params, laser, max_batch=6, max_base_scans=1) for _ in range(2)]
    with pytest.raises(M.B2SError) as e:
        ms[0].set_scans(sets[0][0], sets[0][1]); ms[0].add_scans(sets[0][2], sets[0][3]); ms[0].correlate_scan_end()
    assert e.value.status == abi.B2S_ERR_BAD_STATE
    want = []
    for r, p, br, bp in sets:
        ms[0].set_scans(r, p); ms[0].add_scans(br, bp)
        want.append(ms[0].correlate_scan(p, se))

    def begin(i):
        r, p, br, bp = sets[i]
        ms[i % 2].set_scans(r, p); ms[i % 2].add_scans(br, bp); ms[i % 2].correlate_scan_begin(p, se)

    got = []
    begin(0)
    for i in range(4):
        if i + 1 < 4:
            begin(i + 1)
        got.append(ms[i % 2].correlate_scan_end())
    for w, g in zip(want, got):
        for a, b in zip(w, g):
            assert np.array_equal(a, b)
    for m in ms:
        m.close()


def test_scan_pool_base_sets(pkg):
    """b2s_matcher_pool_append + _add_scans_pool: base scans referenced by pool row give the same correlation grids and
    the same MatchScan results as uploading their readings; rows out of range are refused."""
    abi, synth, M = pkg.abi, pkg.synth, pkg.load("matcher")
    laser = synth.Laser()
    world, poses, ranges = synth.make_trajectory(12, 9, laser, step_xy=0.1, step_th_deg=2)
    params = abi.matcher_params(1.5, 0.05, 0.03, 9.25)
    chains = [[0, 1, 2, 3], [2, 3, 4, 4], [5, 6, 7, 7]]          # ragged chains padded by repetition
    cur = [4, 5, 8]
    cr, cp = ranges[cur], poses[cur] + np.array([0.05, -0.04, 0.02])
    a = M.ScanMatcher(params, abi.laser_from(laser), max_batch=3, max_base_scans=4)
    a.set_scans(cr, cp)
    a.add_scans(np.stack([ranges[c] for c in chains]), np.stack([poses[c] for c in chains]))
    want = a.match_scan()
    b = M.ScanMatcher(params, abi.laser_from(laser), max_batch=3, max_base_scans=4)
    rows = [b.pool_append(ranges[i]) for i in range(9)]
    assert rows == list(range(9))
    b.set_scans(cr, cp)
    b.add_scans_pool(chains, np.stack([poses[c] for c in chains]))
    got = b.match_scan()
    for k in range(3):
        assert np.array_equal(a.grid(k)[0], b.grid(k)[0])
    for x, y in zip(want, got):
        assert np.array_equal(x, y)
    with pytest.raises(M.B2SError) as e:
        b.add_scans_pool([[0, 1, 2, 9]] * 3, np.stack([poses[c] for c in chains]))
    assert e.value.status == abi.B2S_ERR_OUT_OF_RANGE
    a.close(), b.close()


def test_cfg4_largest_window_full_size(pkg, M):
    """BASELINE cfg 4's largest shape at FULL size: one match, 61 x 61 x 361 candidates (+-0.75 m @0.025 m, +-45 deg @0.25
    deg) on the 807 x 807 grid: all 1 343 281 integer response sums bit-exact vs the restatement, response / pose /
    covariance within 1e-9 (the oracle takes a few seconds for its 1.45 G lookups)."""
    abi, synth = pkg.abi, pkg.synth
    l4 = synth.Laser(range_threshold=9.25)
    params, laser = abi.matcher_params(1.5, 0.025, 0.03, 9.25), abi.laser_from(l4)
    c = synth.make_match_case(4_100_001, l4)
    m = M.ScanMatcher(params, laser, max_batch=1, max_base_scans=1)
    m.set_scans(c.ranges[None], c.odom_pose[None])
    m.add_scans(c.base_ranges[None, None], c.base_pose[None, None])
    se = abi.Search(0.75, 0.75, 0.025, 0.025, 45 * D, 0.25 * D, 1, 0)
    assert (abi.n_steps(0.75, 0.025), abi.n_steps(45 * D, 0.25 * D)) == (61, 361)
    gpu = m.correlate_scan(c.odom_pose[None], se)
    assert m.last_timing()["path"] == 2
    pm = port_case(abi, params, laser, c.ranges, c.odom_pose, c.base_ranges[None], c.base_pose[None])
    rc, r = pm.correlate_scan(pm.sp, se, want_sums=True)
    assert rc == 0
    assert np.array_equal(m.response_sums(0, (61, 61, 361)), pm.last_sums)
    gpu = m.correlate_scan(pm.sp[None], se)
    assert_result(gpu, 0, r)
    m.close()
