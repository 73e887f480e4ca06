"""Probe (not a test): cfg-3 Hector stream rates of the device-resident processor — per-scan calls vs the one-call stream,
exact vs fast mode — plus the batched multi-map mapping step.  usage: python tools/hector_stream_probe.py [n_scans] [batch]"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
synth, H = pkg.synth, pkg.load("hector")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
laser = synth.Laser()
_, poses, ranges = synth.make_trajectory(22, n, laser, step_xy=0.05, step_th_deg=1.0)
pts = [H.scan_to_data_container(ranges[i], laser, 0.05, max_dist=30.0, min_dist=0.2) for i in range(n)]
kw = dict(resolution=0.05, size_x=1000, size_y=1000, start=(0.5, 0.5), levels=3, update_free=0.4, update_occupied=0.9,
          min_dist=0.4, min_angle=0.9)
out = {"scans": n}
for exact, l2 in ((True, "0"), (False, "0")):
    os.environ["B2S_HS_L2_LOADS"] = l2
    tag = ("exact" if exact else "fast") + ("_l2loads" if l2 == "1" else "")
    hs = H.HectorSlam(exact=exact, **kw)
    est = poses[0].astype(np.float32)
    est, _ = hs.update(pts[0], (0, 0), est)
    t0 = time.perf_counter()
    for i in range(1, n):
        est, _ = hs.update(pts[i], (0, 0), est)
    st = hs.stats()
    dt = time.perf_counter() - t0
    out[f"per_scan_{tag}"] = {"scans_per_s": (n - 1) / dt, "us_per_scan": 1e6 * dt / (n - 1), "updates": st["updated"],
                              "match_ms": st["match_ms"], "update_ms": st["update_ms"],
                              "xy_err": float(np.abs(est[:2] - poses[-1][:2]).max())}
    hs.close()
    hs = H.HectorSlam(exact=exact, **kw)
    hs.process_stream(pts[:8], (0, 0), first_hint=poses[0].astype(np.float32))
    hs.reset()
    t0 = time.perf_counter()
    p, u, _ = hs.process_stream(pts, (0, 0), first_hint=poses[0].astype(np.float32))
    dt = time.perf_counter() - t0
    st = hs.stats()
    out[f"profile_{tag}"] = hs.profile()
    out[f"stream_{tag}"] = {"scans_per_s": n / dt, "us_per_scan": 1e6 * dt / n, "updates": int(u.sum()),
                            "cell_visits": st["cell_visits"], "match_ms": st["match_ms"], "update_ms": st["update_ms"],
                            "xy_err": float(np.abs(p[-1][:2] - poses[-1][:2]).max())}
    hs.close()
os.environ["B2S_HS_L2_LOADS"] = "0"
# batched mapping (map_without_matching): B independent 1000^2 maps, one scan each per step, scans resident on the device
import torch
cap = max(len(p) for p in pts)
hb = H.HectorSlam(batch=B, max_points=cap, **dict(kw, min_dist=0.0, min_angle=0.0))
steps = 40
dp = torch.zeros((steps, B, cap, 2), dtype=torch.float32, device="cuda")
dn = torch.zeros((steps, B), dtype=torch.int32, device="cuda")
dh = torch.zeros((steps, B, 3), dtype=torch.float32, device="cuda")
for s in range(steps):
    for b in range(B):
        i = (s * 7 + b * 13) % n
        dp[s, b, :len(pts[i])] = torch.from_numpy(pts[i]).cuda()
        dn[s, b] = len(pts[i])
        dh[s, b] = torch.from_numpy(poses[i].astype(np.float32)).cuda()
torch.cuda.synchronize()
for s in range(3):
    hb.update_batch_device(dp[s].data_ptr(), dn[s].data_ptr(), cap, (0, 0), dh[s].data_ptr(), True)
hb.sync()
v0 = hb.stats()["cell_visits"]
t0 = time.perf_counter()
for s in range(3, steps):
    hb.update_batch_device(dp[s].data_ptr(), dn[s].data_ptr(), cap, (0, 0), dh[s].data_ptr(), True)
hb.sync()
dt = time.perf_counter() - t0
v = hb.stats()["cell_visits"] - v0
out["batched_mapping"] = {"batch": B, "steps": steps - 3, "scans_per_s": B * (steps - 3) / dt, "cell_visits": v,
                          "cells_per_s": v / dt, "A2_GBps": 2 * 8 * v / dt / 1e9, "ms_per_step": 1e3 * dt / (steps - 3)}
print(json.dumps(out, indent=1))
# mapping-only stream (one map, every scan updates) at two lengths: fixed or per-scan cost?
for n_m in (100, 1000):
    hs = H.HectorSlam(**dict(kw, min_dist=0.0, min_angle=0.0))
    hs.process_stream(pts[:4], (0, 0), pose_hints=poses[:4].astype(np.float32), map_without_matching=True)
    hs.reset()
    hs.sync()
    ph = poses[:n_m].astype(np.float32)
    t0 = time.perf_counter()
    hs.process_stream(pts[:n_m], (0, 0), pose_hints=ph, map_without_matching=True)
    dt = time.perf_counter() - t0
    st = hs.stats()
    print(json.dumps({"mapping_only_stream": n_m, "ms": dt * 1e3, "us_per_scan": 1e6 * dt / n_m, "update_ms": st["update_ms"],
                      "match_ms": st["match_ms"], "visits": st["cell_visits"], "profile": hs.profile()}))
    hs.close()
