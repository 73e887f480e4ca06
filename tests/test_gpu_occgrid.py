"""GPU parity tests for K2c (karto::OccupancyGrid::CreateFromScans) through the C ABI.  Gate: pass/hit counters,
cell states and grid dimensions bit-exact vs the restatement and the reference-made golden vectors."""
import os

import numpy as np
import pytest

from oracle import port

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def O(pkg):
    m = pkg.load("matcher")
    assert m.device_count() > 0
    return pkg.load("occgrid")


def same(a, b):
    assert (a["width"], a["height"], a["width_step"]) == (b["width"], b["height"], b["width_step"])
    assert np.array_equal(a["offset"], b["offset"])
    assert np.array_equal(a["passes"], b["passes"])
    assert np.array_equal(a["hits"], b["hits"])
    assert np.array_equal(a["cells"], b["cells"])


def test_trajectory_vs_oracle(pkg, O):
    abi, synth = pkg.abi, pkg.synth
    laser = synth.Laser()
    world, poses, ranges = synth.make_trajectory(3, 70, laser, step_xy=0.3, step_th_deg=8)
    ranges[4, ::50] = np.nan
    ranges[5, ::70] = np.inf
    ranges[6, ::90] = 0.05
    ranges[7, ::33] = 40.0
    al = abi.laser_from(laser)
    g = O.OccupancyGrid(al, ranges, poses, 0.05)
    a = g.arrays()
    b = port.occupancy_grid(al, ranges, poses, 0.05)
    same(a, b)
    assert a["cell_visits"] == b["cell_visits"]
    ros = g.ros_map()
    c = a["cells"][:, :a["width"]]
    assert ((ros == -1) == (c == 0)).all() and ((ros == 100) == (c == 100)).all() and ((ros == 0) == (c == 255)).all()
    g.close()


def test_golden_small_and_multibase(pkg, O):
    abi, synth = pkg.abi, pkg.synth
    g = np.load(os.path.join(G, "karto_small.npz"))
    laser = synth.Laser(type=2, n_readings=361, min_angle=synth.deg2rad(-90), max_angle=synth.deg2rad(90),
                        angular_resolution=synth.deg2rad(0.5), min_range=0.0, max_range=80.0, range_threshold=6.0)
    for tag in ("clean", "dropout"):
        og = O.OccupancyGrid(abi.laser_from(laser), np.stack([g[f"{tag}_base_ranges"], g[f"{tag}_ranges"]]),
                             np.stack([g[f"{tag}_base_pose"], g[f"{tag}_pose"]]), 0.05)
        a = og.arrays()
        assert [a["width"], a["height"], a["width_step"]] == list(g[f"{tag}_occ_dims"])
        assert np.array_equal(a["offset"], g[f"{tag}_occ_offset"])
        assert np.array_equal(a["passes"], g[f"{tag}_occ_pass"]) and np.array_equal(a["hits"], g[f"{tag}_occ_hit"])
        assert np.array_equal(a["cells"], g[f"{tag}_occ_cells"])
        og.close()
    g = np.load(os.path.join(G, "karto_multibase.npz"))
    laser = synth.Laser(type=0, n_readings=180, min_angle=synth.deg2rad(-90), max_angle=synth.deg2rad(90),
                        angular_resolution=synth.deg2rad(1.0), min_range=0.05, max_range=25.0, range_threshold=8.0,
                        offset_pose=(0.12, -0.03, 0.05))
    poses = g["poses"].copy()
    poses[12] = g["odom"]
    og = O.OccupancyGrid(abi.laser_from(laser), g["ranges"], poses, 0.1)
    a = og.arrays()
    assert [a["width"], a["height"], a["width_step"]] == list(g["occ_dims"])
    assert np.array_equal(a["passes"], g["occ_pass"]) and np.array_equal(a["hits"], g["occ_hit"])
    assert np.array_equal(a["cells"], g["occ_cells"])
    og.close()


def test_empty_and_bad(pkg, O):
    abi, synth = pkg.abi, pkg.synth
    M = pkg.load("matcher")
    al = abi.laser_from(synth.Laser())
    g = O.OccupancyGrid(al, np.zeros((0, 1081)), np.zeros((0, 3)), 0.05)
    assert g.is_null  # CreateFromScans returns NULL for no scans (Karto.h:5661-5664)
    with pytest.raises(M.B2SError) as e:
        O.OccupancyGrid(al, np.ones((1, 1081)), np.zeros((1, 3)), 0.0)
    assert e.value.status == abi.B2S_ERR_BAD_PARAMS


def test_large_map_properties(pkg, O):
    """2000 scans x 1081 beams (a full map rebuild): counter conservation laws that do not need the oracle —
    SUM(hit) = number of valid end points inside the map, SUM(pass) = cell_visits - SUM(hit), hit <= pass — plus a
    bit-exact oracle comparison on a 100-scan prefix."""
    abi, synth = pkg.abi, pkg.synth
    laser = synth.Laser()
    world, poses, ranges = synth.make_trajectory(9, 200, laser, step_xy=0.25, step_th_deg=6)
    poses, ranges = np.tile(poses, (10, 1)), np.tile(ranges, (10, 1))
    al = abi.laser_from(laser)
    g = O.OccupancyGrid(al, ranges, poses, 0.05)
    a = g.arrays()
    assert (a["hits"] <= a["passes"]).all()
    assert int(a["passes"].sum()) + int(a["hits"].sum()) == a["cell_visits"]
    b = port.occupancy_grid(al, ranges[:200], poses[:200], 0.05)
    assert np.array_equal(a["passes"], b["passes"] * 10) and np.array_equal(a["hits"], b["hits"] * 10)
    t = g.last_timing()
    assert t["raytrace_ms"] > 0
    g.close()


@pytest.mark.parametrize("world", [2, 3, 5])
def test_sharded_scan_list_equals_whole(pkg, O, world):
    """the scan list sharded over `world` ranks (emulated on one GPU): per-shard boxes reduced by
    min/max, per-shard counters summed, thresholded once — identical to CreateFromScans over the whole list, incl.
    an EMPTY shard (world 5 > 4 scans in the second case) and the device-pointer alias used for the NCCL all-reduce."""
    import torch
    abi, synth, par = pkg.abi, pkg.synth, pkg.load("parallel")
    laser = synth.Laser()
    al = abi.laser_from(laser)
    for n_scans in (31, 4):
        _, poses, ranges = synth.make_trajectory(8, n_scans, laser, step_xy=0.3, step_th_deg=8)
        whole = O.OccupancyGrid(al, ranges, poses, 0.05)
        bounds = [par.shard_bounds(n_scans, world, r) for r in range(world)]
        boxes = np.stack([O.scans_bbox(al, ranges[lo:hi], poses[lo:hi]) for lo, hi in bounds])
        bbox = np.concatenate([boxes[:, :2].min(axis=0), boxes[:, 2:].max(axis=0)])
        shards = [O.OccupancyGrid(al, ranges[lo:hi], poses[lo:hi], 0.05, bbox=bbox) for lo, hi in bounds]
        arrs = [s.arrays() for s in shards]
        # device alias: torch sees the library's counters without a copy
        dpass, dhit = shards[0].device_counters()
        tp = torch.as_tensor(dpass, device="cuda")
        assert np.array_equal(tp.cpu().numpy().view(np.uint32), arrs[0]["passes"].reshape(-1))
        tp += torch.as_tensor(np.sum([a["passes"] for a in arrs[1:]], axis=0).reshape(-1).astype(np.int32), device="cuda")
        th = torch.as_tensor(dhit, device="cuda")
        th += torch.as_tensor(np.sum([a["hits"] for a in arrs[1:]], axis=0).reshape(-1).astype(np.int32), device="cuda")
        torch.cuda.synchronize()
        shards[0].update()
        same(shards[0].arrays(), whole.arrays())
        # host path (what the gloo tests use)
        shards[1].set_counters(np.sum([a["passes"] for a in arrs], axis=0), np.sum([a["hits"] for a in arrs], axis=0))
        shards[1].update()
        same(shards[1].arrays(), whole.arrays())
        assert sum(a["cell_visits"] for a in arrs) == whole.arrays()["cell_visits"]
        for s in shards:
            s.close()
        whole.close()
    # single-process form of the orchestration helper
    g1 = par.occupancy_grid_sharded(O, al, ranges, poses, 0.05)
    same(g1.arrays(), O.OccupancyGrid(al, ranges, poses, 0.05).arrays())

