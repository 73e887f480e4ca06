"""World-size-2 (and 3, ragged) gloo tests of the batch sharding / result gathering used by the N>1 path.
CPU only: the per-match work is a stand-in (a deterministic function of the match index)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import importlib
    import torch.distributed as dist
    par = importlib.import_module("creating-2d-laser-slam-from-scratch_b200.parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = par.shard_bounds(n_total, world, rank)
    idx = np.arange(lo, hi, dtype=np.float64)
    local = np.stack([idx, idx * idx, np.sin(idx)], axis=1)  # stand-in for (response, pose...) rows
    full = par.gather_results(local, n_total)
    t = par.max_over_ranks(10.0 + rank)
    q.put((rank, full, t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 1024), (2, 7), (3, 10)])
def test_shard_and_gather(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(n_total, dtype=np.float64)
    expect = np.stack([idx, idx * idx, np.sin(idx)], axis=1)
    for rank, full, t in outs:
        assert np.array_equal(full, expect)
        assert t == 10.0 + world - 1


def test_shard_bounds_cover_exactly(pkg):
    par = pkg.load("parallel")
    for n in (0, 1, 5, 1024, 1027):
        for w in (1, 2, 3, 8):
            b = [par.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


class StandInMatcher:
    """CPU stand-in with the three split-sweep phases of ScanMatcher (matcher.py) over a seeded response volume
    R[B, na, cells]; only the collective plumbing of correlate_scan_angle_split is under test here."""

    def __init__(self, B, na, cells, seed=5):
        rng = np.random.default_rng(seed)
        self.R = rng.integers(0, 6, size=(B, na, cells)).astype(np.float64)  # small range => many exact ties
        self.batch = B

    def split_begin(self, centers, search, k_first, k_count):
        self.sub = self.R[:, k_first:k_first + k_count]
        self.k_first = k_first
        if k_count == 0:
            return np.full(self.batch, -1.0), np.zeros((self.batch, self.R.shape[2])), np.zeros(self.batch, np.int32)
        return self.sub.max(axis=(1, 2)), self.sub.max(axis=1), np.zeros(self.batch, np.int32)

    def split_ties(self, best):
        out = np.zeros((self.batch, 5))
        for b in range(self.batch):
            k, c = np.nonzero(self.sub[b] == best[b])
            out[b] = [c.sum(), (c * c).sum(), (k + self.k_first).sum(), 0.0, len(k)]
        return out

    def split_finish(self, best, ties, probs):
        return best.copy(), ties[:, :3] / ties[:, 4:5], probs.sum(axis=1), np.zeros(self.batch, np.int32), ties[:, 4].astype(np.int32)


def split_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import importlib
    import torch.distributed as dist
    par = importlib.import_module("creating-2d-laser-slam-from-scratch_b200.parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = StandInMatcher(4, 7, 25)
    out = par.correlate_scan_angle_split(m, None, None, 7)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_angle_split_collectives(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=split_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = StandInMatcher(4, 7, 25)
    best, probs, _ = m.split_begin(None, None, 0, 7)
    expect = m.split_finish(best, m.split_ties(best), probs)
    for rank, out in outs:
        for a, b in zip(out, expect):
            assert np.array_equal(a, b)


class _StandInGrid:
    """numpy stand-in for occgrid.OccupancyGrid (shard form): every scan adds 3 passes + 1 hit at its pose's cell."""

    class _Info:
        data_size = 0

    def __init__(self, laser, ranges, poses, resolution, device=0, bbox=None):
        w = int(round((bbox[2] - bbox[0]) / resolution)) + 1
        h = int(round((bbox[3] - bbox[1]) / resolution)) + 1
        self.p, self.h = np.zeros((h, w), np.uint32), np.zeros((h, w), np.uint32)
        for x, y, _ in np.asarray(poses).reshape(-1, 3):
            cx, cy = int(round((x - bbox[0]) / resolution)), int(round((y - bbox[1]) / resolution))
            self.p[cy, cx] += 3
            self.h[cy, cx] += 1
        self.info = self._Info()
        self.info.data_size = w * h
        self.cells = None

    def arrays(self):
        return dict(passes=self.p, hits=self.h, cells=self.cells)

    def set_counters(self, p, h):
        self.p, self.h = np.asarray(p, np.uint32).reshape(self.p.shape), np.asarray(h, np.uint32).reshape(self.h.shape)

    def update(self):
        self.cells = np.where(self.p > 2, 100, 0).astype(np.uint8)


class _StandInOccMod:
    OccupancyGrid = _StandInGrid

    @staticmethod
    def scans_bbox(laser, ranges, poses, device=0):
        p = np.asarray(poses).reshape(-1, 3)
        big = 999999999999999999.99999
        return np.array([p[:, 0].min(), p[:, 1].min(), p[:, 0].max(), p[:, 1].max()]) if len(p) else np.array([big, big, -big, -big])


def occ_worker(rank, world, port, n_scans, q):
    sys.path.insert(0, ROOT)
    import importlib
    import torch.distributed as dist
    par = importlib.import_module("creating-2d-laser-slam-from-scratch_b200.parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    poses = np.random.default_rng(3).uniform(-5, 5, size=(n_scans, 3))
    lo, hi = par.shard_bounds(n_scans, world, rank)
    g = par.occupancy_grid_sharded(_StandInOccMod, None, None, poses[lo:hi], 0.5)
    q.put((rank, g.p, g.h, g.cells))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_scans", [(2, 40), (3, 2)])  # (3, 2): one rank has an empty shard
def test_sharded_occupancy_grid_collectives(world, n_scans):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=occ_worker, args=(r, world, port, n_scans, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    poses = np.random.default_rng(3).uniform(-5, 5, size=(n_scans, 3))
    whole = _StandInGrid(None, None, poses, 0.5, bbox=_StandInOccMod.scans_bbox(None, None, poses))
    whole.update()
    for rank, p, h, cells in outs:
        assert np.array_equal(p, whole.p) and np.array_equal(h, whole.h) and np.array_equal(cells, whole.cells)
