"""Workload for an ncu capture of the one-call Hector stream kernel (k_hs_stream): a warm-up call, then ONE call over
n scans.  usage: python tools/hector_stream_ncu.py [n_scans] [exact|fast]"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
synth, H = pkg.synth, pkg.load("hector")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
exact = (sys.argv[2] if len(sys.argv) > 2 else "fast") == "exact"
laser = synth.Laser()
_, poses, ranges = synth.make_trajectory(22, n, laser, step_xy=0.05, step_th_deg=1.0)
pts = [H.scan_to_data_container(ranges[i], laser, 0.05, max_dist=30.0, min_dist=0.2) for i in range(n)]
kw = dict(resolution=0.05, size_x=1000, size_y=1000, start=(0.5, 0.5), levels=3, update_free=0.4, update_occupied=0.9,
          min_dist=0.4, min_angle=0.9)
hs = H.HectorSlam(exact=exact, **kw)
first = poses[0].astype(np.float32)
hs.process_stream(pts[:8], (0, 0), first_hint=first)
hs.reset()
p, u, _ = hs.process_stream(pts, (0, 0), first_hint=first)
print("cluster", hs.cluster_size(), "updates", int(u.sum()), "xy err", float(np.abs(p[-1][:2] - poses[-1][:2]).max()))
