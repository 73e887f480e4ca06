"""Real multi-GPU check of the angle-split sweep; not collected by pytest.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_checks.py

Every rank holds the same scans + grids, sweeps its share of the angle steps, and the best response / per-cell maxima /
tie sums are all-reduced over NCCL.  Rank 0 also runs the whole sweep alone and prints one JSON line with the largest
difference and device timings (max over ranks)."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
D = 0.01745329251994329577


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    abi, synth, M, par = pkg.abi, pkg.synth, pkg.load("matcher"), pkg.load("parallel")
    B = int(os.environ.get("B2S_SPLIT_BATCH", 16))
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    cases = [synth.make_match_case(700 + s, synth.Laser()) for s in range(min(B, 8))]
    pick = [cases[i % len(cases)] for i in range(B)]
    ranges, poses = np.stack([c.ranges for c in pick]), np.stack([c.odom_pose for c in pick])
    bran, bpos = np.stack([c.base_ranges for c in pick])[:, None, :], np.stack([c.base_pose for c in pick])[:, None, :]
    m = M.ScanMatcher(params, laser, max_batch=B, max_base_scans=1, device=local)
    m.set_scans(ranges, poses)
    m.add_scans(bran, bpos)
    A, R = 22.5 * D, 0.25 * D
    se = abi.Search(0.75, 0.75, 0.05, 0.05, A, R, 1, 0)
    na = abi.n_steps(A, R)
    centers = poses.copy()
    worst, t_split, t_whole = 0.0, [], []
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = par.correlate_scan_angle_split(m, centers, se, na)
        torch.cuda.synchronize()
        t_split.append(par.max_over_ranks(time.perf_counter() - t0))
        t0 = time.perf_counter()
        whole = m.correlate_scan(centers, se)
        t_whole.append(par.max_over_ranks(time.perf_counter() - t0))
        for a, b in zip(got, whole):
            worst = max(worst, float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)))))
        assert (whole[3] == 0).all()
    worst = par.max_over_ranks(worst)
    if rank == 0:
        print(json.dumps({"check": "angle_split_vs_whole_sweep", "n_gpus": world, "batch": B, "angles": na,
                          "max_abs_diff": worst, "ok": bool(worst <= 1e-9),
                          "host_ms_split_incl_allreduce": round(1e3 * min(t_split[1:]), 3),
                          "host_ms_whole_one_gpu": round(1e3 * min(t_whole[1:]), 3)}))
    # ---- the same split with the collectives inside the library (ncclAllReduce on device buffers, the caller's communicator)
    comm = par.NcclComm()
    worst_lib, t_lib, tim = 0.0, [], {}
    big = os.environ.get("B2S_SPLIT_BIG", "0") == "1"
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        got, tim = par.correlate_scan_split_in_library(m, centers, se, comm)
        t_lib.append(par.max_over_ranks(time.perf_counter() - t0))
        for a, b in zip(got, whole):
            worst_lib = max(worst_lib, float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)))))
    worst_lib = par.max_over_ranks(worst_lib)
    if rank == 0:
        print(json.dumps({"check": "in_library_nccl_split_vs_whole_sweep", "n_gpus": world, "batch": B, "angles": na,
                          "max_abs_diff": worst_lib, "ok": bool(worst_lib <= 1e-9),
                          "host_ms_split_in_library": round(1e3 * min(t_lib[1:]), 3),
                          "host_ms_split_python_staged": round(1e3 * min(t_split[1:]), 3),
                          "host_ms_whole_one_gpu": round(1e3 * min(t_whole[1:]), 3), "collective_ms": tim}))
    worst = max(worst, worst_lib)
    if big:  # BASELINE cfg 4's largest window: ONE match, 61 x 61 x 361 on a 0.025 m grid, angles split over the ranks
        l4 = synth.Laser(range_threshold=9.25)
        p4 = abi.matcher_params(1.5, 0.025, 0.03, 9.25)
        c4 = synth.make_match_case(4_100_000, l4)
        m4 = M.ScanMatcher(p4, abi.laser_from(l4), max_batch=1, max_base_scans=1, device=local)
        m4.set_scans(c4.ranges[None], c4.odom_pose[None])
        m4.add_scans(c4.base_ranges[None, None], c4.base_pose[None, None])
        se4 = abi.Search(0.75, 0.75, 0.025, 0.025, 45 * D, 0.25 * D, 1, 0)
        whole4 = m4.correlate_scan(c4.odom_pose[None], se4)
        t4, t4w, tim4, w4 = [], [], {}, 0.0
        for it in range(4):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            got4, tim4 = par.correlate_scan_split_in_library(m4, c4.odom_pose[None], se4, comm)
            t4.append(par.max_over_ranks(time.perf_counter() - t0))
            t0 = time.perf_counter()
            m4.correlate_scan(c4.odom_pose[None], se4)
            t4w.append(par.max_over_ranks(time.perf_counter() - t0))
            for a, b in zip(got4, whole4):
                w4 = max(w4, float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)))))
        w4 = par.max_over_ranks(w4)
        if rank == 0:
            print(json.dumps({"check": "cfg4_largest_window_one_match_split", "n_gpus": world, "window": [61, 61, 361],
                              "max_abs_diff": w4, "ok": bool(w4 <= 1e-9), "host_ms_split_in_library": round(1e3 * min(t4[1:]), 3),
                              "host_ms_whole_one_gpu": round(1e3 * min(t4w[1:]), 3), "collective_ms": tim4}))
        worst = max(worst, w4)
        m4.close()
    # ---- K2c: the scan list sharded over the ranks, counters all-reduced in place over NCCL ----
    O = pkg.load("occgrid")
    n_scans = int(os.environ.get("B2S_SPLIT_SCANS", 400))
    _, tposes, tranges = synth.make_trajectory(21, n_scans, synth.Laser(), step_xy=0.2, step_th_deg=5)
    lo, hi = par.shard_bounds(n_scans, world, rank)
    t_sh, t_wh = [], []
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        gs = par.occupancy_grid_sharded(O, laser, tranges[lo:hi], tposes[lo:hi], 0.05, device=local, nccl_comm=comm)
        t_sh.append(par.max_over_ranks(time.perf_counter() - t0))
        t0 = time.perf_counter()
        gw = O.OccupancyGrid(laser, tranges, tposes, 0.05, device=local)
        t_wh.append(par.max_over_ranks(time.perf_counter() - t0))
        a, b = gs.arrays(), gw.arrays()
        same = all(np.array_equal(a[k], b[k]) for k in ("passes", "hits", "cells", "offset")) and \
            (a["width"], a["height"]) == (b["width"], b["height"])
        gs.close(); gw.close()
    same_all = par.max_over_ranks(0.0 if same else 1.0) == 0.0
    if rank == 0:
        print(json.dumps({"check": "occupancy_grid_sharded_vs_whole", "n_gpus": world, "scans": n_scans,
                          "bit_identical_on_all_ranks": bool(same_all),
                          "host_ms_sharded_incl_allreduce": round(1e3 * min(t_sh[1:]), 3),
                          "host_ms_whole_one_gpu": round(1e3 * min(t_wh[1:]), 3)}))
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if not same_all:
        sys.exit(2)
    if worst > 1e-9:
        sys.exit(1)


if __name__ == "__main__":
    main()
