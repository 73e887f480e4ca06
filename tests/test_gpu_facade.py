"""The C++ façade (include/b200slam/karto_facade.hpp) compiled into a standalone binary and run on the GPU box:
reference call sequence Create -> MatchScan -> CorrelateScan -> OccupancyGrid::CreateFromScans, checked against
the restatement."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "creating-2d-laser-slam-from-scratch_b200")
D = 0.01745329251994329577


def build_demo(tmp):
    exe = os.path.join(tmp, "facade_demo")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "facade_demo.cpp"), "-o", exe, "-L" + PKG, "-lb200slam",
                           "-Wl,-rpath," + PKG])
    return exe


def test_facade_compiles(tmp_path, pkg):
    pkg.load("matcher")  # the library must exist
    assert os.path.exists(build_demo(str(tmp_path)))


@pytest.mark.gpu
def test_facade_runs_reference_call_sequence(tmp_path, pkg):
    abi, synth = pkg.abi, pkg.synth
    exe = build_demo(str(tmp_path))
    world, poses, ranges = synth.make_trajectory(8, 6, synth.Laser(), step_xy=0.15, step_th_deg=3)
    odom = poses[5] + np.array([0.1, -0.08, 0.03])
    poses_in = poses.copy()
    poses_in[5] = odom
    text = "6\n" + "\n".join(" ".join(repr(float(v)) for v in np.concatenate([poses_in[i], ranges[i]])) for i in range(6))
    out = subprocess.run([exe], input=text, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    lines = {l.split()[0]: [float(x) for x in l.split()[1:]] for l in out.stdout.strip().splitlines()}
    params, laser = abi.matcher_params(1.5, 0.05, 0.03, 9.25), abi.laser_from(synth.Laser())
    pm = port.PortMatcher(params, laser)
    rc, res = pm.match_scan(ranges[5], odom, ranges[:5], poses[:5])
    assert rc == 0
    assert abs(lines["match"][0] - res.response) < 1e-9 and np.allclose(lines["match"][1:4], res.pose[:], atol=1e-9)
    assert np.allclose(lines["match"][4:7], [res.cov[0], res.cov[4], res.cov[8]], atol=1e-9)
    pm.set_scan(ranges[5], odom)
    rc, r2 = pm.correlate_scan(pm.sp, abi.Search(0.75, 0.75, 0.05, 0.05, 22.5 * D, 0.25 * D, 1, 0))
    assert abs(lines["corr"][0] - r2.response) < 1e-9 and np.allclose(lines["corr"][1:4], r2.pose[:], atol=1e-9)
    og = port.occupancy_grid(laser, ranges, poses_in, 0.05)
    c = og["cells"][:, :og["width"]]
    assert lines["grid"] == [og["width"], og["height"], int((c == 100).sum()), int((c == 255).sum())]
    # the front-end classes: every scan becomes a key frame; the last corrected pose stays near the true pose, and the
    # Hector processor (first scan mapped at its true pose, then self-driven) tracks the trajectory
    assert lines["mapper"][:2] == [6, 6] and np.abs(np.array(lines["mapper"][2:4]) - poses[5][:2]).max() < 0.1
    assert np.abs(np.array(lines["hector"][:2]) - poses[5][:2]).max() < 0.1
