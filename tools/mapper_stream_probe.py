"""Small lesson6-front-end stream (shipped indoor yaml) for profiling: python tools/mapper_stream_probe.py [key_frames]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
synth, abi, MP = pkg.synth, pkg.abi, pkg.load("mapper")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
lm = synth.Laser(range_threshold=12.0)
_, tru, odo, rng = synth.make_loop_trajectory(17, n, lm, radius=2.0, step=0.25)
yaml = dict(scan_buffer_size=110, scan_buffer_maximum_scan_distance=100.0, link_match_minimum_response_fine=0.1,
            link_scan_maximum_distance=1.5, loop_search_maximum_distance=10.0, loop_match_minimum_chain_size=5,
            loop_match_maximum_variance_coarse=9.0, loop_match_minimum_response_coarse=0.35,
            loop_match_minimum_response_fine=0.45, minimum_travel_heading=0.174, loop_search_size=10.0,
            both_distance_variance_penalty=0.25, both_angle_variance_penalty=0.01, both_fine_search_angle_offset=0.00349,
            both_coarse_search_angle_offset=0.349, both_coarse_angle_resolution=0.0349, both_use_response_expansion=1)
m = MP.Mapper(MP.default_params(12.0, **yaml), abi.laser_from(lm))
m.process(rng[0], odo[0], 0.0)
t0 = time.perf_counter()
for i in range(1, n):
    m.process(rng[i], odo[i], 0.1 * i)
dt = time.perf_counter() - t0
print(f"{(n - 1) / dt:.1f} key frames/s, {1e3 * dt / (n - 1):.2f} ms each", m.stats())
