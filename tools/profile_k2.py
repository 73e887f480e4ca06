"""Launch each K2 / Hector kernel a couple of times on bench-shaped inputs (driver for `ncu -k regex:...`; not a test).
usage: python tools/profile_k2.py [maps]"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
abi, synth = pkg.abi, pkg.synth
O, H, GM = pkg.load("occgrid"), pkg.load("hector"), pkg.load("gmapping")
import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
laser = synth.Laser()
al = abi.laser_from(laser)
# K2c: 2000-scan rebuild (bench shape)
_, poses, ranges = synth.make_trajectory(21, 200, laser, step_xy=0.25, step_th_deg=6)
P, R = np.tile(poses, (10, 1)), np.tile(ranges, (10, 1))
for _ in range(2):
    g = O.OccupancyGrid(al, R, P, 0.05)
    g.close()
# K2b
ang = (np.float32(laser.min_angle) + np.arange(1081, dtype=np.float32) * np.float32(laser.angular_resolution)).astype(np.float64)
gm = GM.GMap(-50, -50, 50, 50, 0.05)
for i in range(3):
    gm.compute_map(ranges[i].astype(np.float32).astype(np.float64), ang, (float(poses[i, 0]), float(poses[i, 1])))
gm.close()
# K2a: stream + batched maps
_, sp, sr = synth.make_trajectory(22, 400, laser, step_xy=0.05, step_th_deg=1.0)
pts = [H.scan_to_data_container(sr[i], laser, 0.05, max_dist=30.0, min_dist=0.2) for i in range(400)]
kw = dict(resolution=0.05, size_x=1000, size_y=1000, start=(0.5, 0.5), levels=3, update_free=0.4, update_occupied=0.9,
          min_dist=0.4, min_angle=0.9)
for exact in (True, False):
    hs = H.HectorSlam(exact=exact, **kw)
    hs.process_stream(pts[:60], (0, 0), first_hint=sp[0].astype(np.float32))
    hs.close()
hs = H.HectorSlam(**dict(kw, min_dist=0.0, min_angle=0.0))
hs.process_stream(pts[:60], (0, 0), pose_hints=sp[:60].astype(np.float32), map_without_matching=True)
hs.close()
cap = max(len(p) for p in pts)
hb = H.HectorSlam(batch=B, max_points=cap, **dict(kw, min_dist=0.0, min_angle=0.0))
steps = 4
hp = np.zeros((steps, B, cap, 2), np.float32); hn = np.zeros((steps, B), np.int32); hh = np.zeros((steps, B, 3), np.float32)
for s in range(steps):
    for b in range(B):
        i = (s * 7 + b * 13) % 400
        hp[s, b, :len(pts[i])] = pts[i]; hn[s, b] = len(pts[i]); hh[s, b] = sp[i]
dp, dn, dh = torch.from_numpy(hp).cuda(), torch.from_numpy(hn).cuda(), torch.from_numpy(hh).cuda()
torch.cuda.synchronize()
for s in range(steps):
    hb.update_batch_device(dp[s].data_ptr(), dn[s].data_ptr(), cap, (0, 0), dh[s].data_ptr(), True)
hb.sync()
hb.close()
# the hall case (maps stream through HBM)
_, hp2, hr2 = synth.make_trajectory(29, 48, laser, half_w=22.0, half_h=22.0, step_xy=0.3, step_th_deg=8, n_boxes=24)
pts2 = [H.scan_to_data_container(hr2[i], laser, 0.05, max_dist=30.0, min_dist=0.2) for i in range(len(hp2))]
cap2 = max(len(p) for p in pts2)
hb = H.HectorSlam(batch=B, max_points=cap2, **dict(kw, min_dist=0.0, min_angle=0.0))
hq = np.zeros((steps, B, cap2, 2), np.float32); hm = np.zeros((steps, B), np.int32); hg = np.zeros((steps, B, 3), np.float32)
for s in range(steps):
    for b in range(B):
        i = (s * 5 + b * 11) % len(pts2)
        hq[s, b, :len(pts2[i])] = pts2[i]; hm[s, b] = len(pts2[i]); hg[s, b] = hp2[i]
dq, dm, dg = torch.from_numpy(hq).cuda(), torch.from_numpy(hm).cuda(), torch.from_numpy(hg).cuda()
torch.cuda.synchronize()
for s in range(steps):
    hb.update_batch_device(dq[s].data_ptr(), dm[s].data_ptr(), cap2, (0, 0), dg[s].data_ptr(), True)
hb.sync()
hb.close()
hb = H.HectorSlam(batch=B, max_points=cap, **kw)  # B SLAM processors: k_hs_match at B CTAs
for s in range(2):
    hb.update_batch_device(dp[s].data_ptr(), dn[s].data_ptr(), cap, (0, 0), dh[s].data_ptr(), False)
hb.sync()
hb.close()
print("profile_k2 done")
