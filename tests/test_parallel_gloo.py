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
