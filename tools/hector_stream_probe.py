"""cfg-3 style Hector SLAM stream for profiling: python tools/hector_stream_probe.py [scans]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
synth, H = pkg.synth, pkg.load("hector")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
laser = synth.Laser()
_, poses, ranges = synth.make_trajectory(22, n, laser, step_xy=0.05, step_th_deg=1.0)
pts = [H.scan_to_data_container(ranges[i], laser, 0.05, max_dist=30.0, min_dist=0.2) for i in range(n)]
hs = H.HectorSlam(resolution=0.05, size_x=1000, size_y=1000, levels=3, update_free=0.4, update_occupied=0.9, min_dist=0.4, min_angle=0.9)
est, _ = hs.update(pts[0], (0, 0), poses[0].astype(np.float32))
t0 = time.perf_counter()
for i in range(1, n):
    est, _ = hs.update(pts[i], (0, 0), est)
dt = time.perf_counter() - t0
print(f"{(n - 1) / dt:.0f} scans/s", hs.stats(), np.abs(est[:2] - poses[-1][:2]).max())
