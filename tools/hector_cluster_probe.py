"""Probe (not a test): fast-mode Hector stream with the match spread over a thread-block cluster vs on one CTA
(B2S_HS_CLUSTER=0), with the matching CTA's cycle counters.  usage: python tools/hector_cluster_probe.py [n_scans]"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
synth, H = pkg.synth, pkg.load("hector")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
laser = synth.Laser()
_, poses, ranges = synth.make_trajectory(22, n, laser, step_xy=0.05, step_th_deg=1.0)
pts = [H.scan_to_data_container(ranges[i], laser, 0.05, max_dist=30.0, min_dist=0.2) for i in range(n)]
kw = dict(resolution=0.05, size_x=1000, size_y=1000, start=(0.5, 0.5), levels=3, update_free=0.4, update_occupied=0.9,
          min_dist=0.4, min_angle=0.9)
first = poses[0].astype(np.float32)
out, traces = {"scans": n}, {}
CONFIGS = [("fast_cluster", {}, False), ("fast_cluster8", {"B2S_HS_CLUSTER": "8"}, False), ("fast_cluster2", {"B2S_HS_CLUSTER": "2"}, False), ("fast_cluster_master_share", {"B2S_HS_MASTER_SHARE": "1"}, False), ("fast_cluster_148ctas", {"B2S_HS_STREAM_CTAS": "148"}, False),
           ("fast_one_cta", {"B2S_HS_CLUSTER": "0"}, False), ("exact", {}, True)]
if len(sys.argv) > 2:
    CONFIGS = [c for c in CONFIGS if c[0] in sys.argv[2].split(",")]
for tag, env, exact in CONFIGS:
    for k in ("B2S_HS_CLUSTER", "B2S_HS_STREAM_CTAS", "B2S_HS_MASTER_SHARE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    hs = H.HectorSlam(exact=exact, **kw)
    hs.process_stream(pts[:8], (0, 0), first_hint=first)
    hs.reset()
    best = None
    for rep in range(2):
        hs.reset()
        t0 = time.perf_counter()
        p, u, _ = hs.process_stream(pts, (0, 0), first_hint=first)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    traces[tag] = p
    st = hs.stats()
    out[tag] = {"cluster_size": hs.cluster_size(), "scans_per_s": n / best, "us_per_scan": 1e6 * best / n,
                "updates": int(u.sum()), "match_ms": st["match_ms"], "update_ms": st["update_ms"],
                "xy_err": float(np.abs(p[-1][:2] - poses[-1][:2]).max()), "profile": hs.profile()}
    est = first
    hs.reset()
    t0 = time.perf_counter()
    for i in range(min(n, 1000)):
        est, _ = hs.update(pts[i], (0, 0), est)
    dt = time.perf_counter() - t0
    out[tag]["per_scan_calls_per_s"] = min(n, 1000) / dt
    hs.close()
if "fast_one_cta" in traces and "fast_cluster" in traces:
    out["max_pose_diff_cluster_vs_one_cta"] = float(np.abs(traces["fast_cluster"] - traces["fast_one_cta"]).max())
if "exact" in traces and "fast_cluster" in traces:
    out["max_pose_diff_cluster_vs_exact"] = float(np.abs(traces["fast_cluster"] - traces["exact"]).max())
print(json.dumps(out))
