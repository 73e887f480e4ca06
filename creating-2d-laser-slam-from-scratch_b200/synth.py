"""Synthetic 2-D worlds and laser scans (SURVEY.md §8(d) "Synthetic inputs").

Workload generator shared by tests/, bench.py and the golden-fixture script.  It only
produces INPUT data (range readings + poses); it never computes a result that is compared,
so it is neither checker nor product.  Everything is seeded (numpy PCG64) and double precision.

Sensor model = the reference's `LaserRangeFinder_Hokuyo_UTM_30LX` preset
(/root/reference/lesson6/lib/open_karto/include/open_karto/Karto.h:4048-4065):
1081 beams, -135 deg .. +135 deg, 0.25 deg, range [0.1, 30] m.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

KT_PI_180 = 0.01745329251994329577  # Math.h:35


def deg2rad(d: float) -> float:
    """math::DegreesToRadians (Math.h:56-59): degrees * KT_PI_180."""
    return d * KT_PI_180


@dataclass
class Laser:
    """Plain-struct stand-in for karto::LaserRangeFinder (Karto.h:3700-4184)."""
    type: int = 4  # LaserRangeFinder_Hokuyo_UTM_30LX
    n_readings: int = 1081
    min_angle: float = deg2rad(-135)
    max_angle: float = deg2rad(135)
    angular_resolution: float = deg2rad(0.25)
    min_range: float = 0.1
    max_range: float = 30.0
    range_threshold: float = 9.25
    offset_pose: tuple = (0.0, 0.0, 0.0)


@dataclass
class World:
    segments: np.ndarray  # [S,4] x0,y0,x1,y1
    half_w: float = 8.0
    half_h: float = 6.0


def _box(cx, cy, w, h, ang):
    c, s = np.cos(ang), np.sin(ang)
    pts = np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) * 0.5
    pts = pts @ np.array([[c, s], [-s, c]]) + np.array([cx, cy])
    return [[*pts[i], *pts[(i + 1) % 4]] for i in range(4)]


def make_world(seed: int = 0, half_w: float = 8.0, half_h: float = 6.0, n_boxes: int = 6) -> World:
    """A closed rectangular room (2*half_w x 2*half_h metres) with seeded rotated boxes inside."""
    rng = np.random.default_rng(seed)
    segs = _box(0.0, 0.0, 2 * half_w, 2 * half_h, 0.0)
    for _ in range(n_boxes):
        while True:
            cx = rng.uniform(-half_w + 1.0, half_w - 1.0)
            cy = rng.uniform(-half_h + 1.0, half_h - 1.0)
            if cx * cx + cy * cy > 2.5 ** 2:  # keep the middle free for the robot
                break
        segs += _box(cx, cy, rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.6), rng.uniform(0, np.pi))
    return World(np.asarray(segs, dtype=np.float64), half_w, half_h)


def cast_scan(world: World, pose, laser: Laser = Laser(), rng: np.random.Generator | None = None,
              sigma: float = 0.01, dropout: float = 0.0) -> np.ndarray:
    """Exact ray cast from `pose` (x, y, heading) + N(0, sigma) noise; `dropout` = fraction of
    readings replaced by NaN / +inf (alternating) to exercise INVALID_SCAN (Karto.h:6478-6483)."""
    n = laser.n_readings
    ang = pose[2] + laser.min_angle + np.arange(n) * laser.angular_resolution
    d = np.stack([np.cos(ang), np.sin(ang)], axis=1)  # [N,2]
    a = world.segments[:, 0:2]
    e = world.segments[:, 2:4] - a  # [S,2]
    ao = a - np.asarray(pose[:2])[None, :]  # [S,2]
    den = d[:, None, 0] * e[None, :, 1] - d[:, None, 1] * e[None, :, 0]  # cross(d,e) [N,S]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (ao[None, :, 0] * e[None, :, 1] - ao[None, :, 1] * e[None, :, 0]) / den
        u = (ao[None, :, 0] * d[:, None, 1] - ao[None, :, 1] * d[:, None, 0]) / den
    ok = (np.abs(den) > 1e-12) & (t > 1e-9) & (u >= 0.0) & (u <= 1.0)
    t = np.where(ok, t, np.inf)
    r = t.min(axis=1)
    r = np.where(np.isfinite(r), r, laser.max_range + 1.0)
    if rng is not None and sigma > 0:
        r = r + rng.normal(0.0, sigma, size=n)
    if rng is not None and dropout > 0:
        m = rng.random(n) < dropout
        idx = np.nonzero(m)[0]
        r[idx[0::2]] = np.nan
        r[idx[1::2]] = np.inf
    return np.ascontiguousarray(r, dtype=np.float64)


@dataclass
class MatchCase:
    """One BASELINE cfg-1/2 unit: a current scan, its odometry guess, and one base scan."""
    world: World
    laser: Laser
    base_pose: np.ndarray
    base_ranges: np.ndarray
    true_pose: np.ndarray
    odom_pose: np.ndarray
    ranges: np.ndarray
    seed: int = 0
    extra: dict = field(default_factory=dict)


def make_match_case(seed: int, laser: Laser = Laser(), dropout: float = 0.0, n_boxes: int = 6,
                    max_xy: float = 0.3, max_th_deg: float = 10.0) -> MatchCase:
    """cfg-1 recipe: base scan at a seeded pose near the room centre, current scan at
    truth = base + small motion, odometry guess = truth + U(+-max_xy, +-max_xy, +-max_th_deg)."""
    rng = np.random.default_rng(1_000_003 * (seed + 1))
    world = make_world(seed, n_boxes=n_boxes)
    base_pose = np.array([rng.uniform(-1.0, 1.0), rng.uniform(-1.0, 1.0), rng.uniform(-np.pi, np.pi)])
    true_pose = base_pose + np.array([rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), rng.uniform(-0.1, 0.1)])
    odom_pose = true_pose + np.array([rng.uniform(-max_xy, max_xy), rng.uniform(-max_xy, max_xy),
                                      deg2rad(rng.uniform(-max_th_deg, max_th_deg))])
    # keep headings inside (-pi, pi] like a tf-derived yaw would be
    for p in (base_pose, true_pose, odom_pose):
        p[2] = (p[2] + np.pi) % (2 * np.pi) - np.pi
    base_ranges = cast_scan(world, base_pose, laser, rng, 0.01, dropout)
    ranges = cast_scan(world, true_pose, laser, rng, 0.01, dropout)
    return MatchCase(world, laser, base_pose, base_ranges, true_pose, odom_pose, ranges, seed)


def make_trajectory(seed: int, n_poses: int, laser: Laser = Laser(), half_w: float = 8.0, half_h: float = 6.0,
                    step_xy: float = 0.05, step_th_deg: float = 1.0, n_boxes: int = 6):
    """Smooth random walk (cfg-3 recipe): <= step_xy metres / <= step_th_deg per step, kept inside the room.
    Returns (world, poses[n,3], ranges[n,N])."""
    rng = np.random.default_rng(7_000_001 * (seed + 1))
    world = make_world(seed, half_w, half_h, n_boxes)
    poses = np.zeros((n_poses, 3))
    ranges = np.zeros((n_poses, laser.n_readings))
    p = np.array([0.0, 0.0, rng.uniform(-np.pi, np.pi)])
    v = np.zeros(3)
    for i in range(n_poses):
        poses[i] = p
        ranges[i] = cast_scan(world, p, laser, rng, 0.01, 0.0)
        v = 0.8 * v + 0.2 * np.array([rng.uniform(-step_xy, step_xy), rng.uniform(-step_xy, step_xy),
                                      deg2rad(rng.uniform(-step_th_deg, step_th_deg))])
        q = p + v
        if abs(q[0]) > 2.0 or abs(q[1]) > 2.0:  # stay in the box-free middle of the room
            v[:2] = -v[:2]
            q = p + v
        q[2] = (q[2] + np.pi) % (2 * np.pi) - np.pi
        p = q
    return world, poses, ranges


def make_loop_trajectory(seed: int, n_poses: int, laser: Laser = Laser(), radius: float = 2.0, step: float = 0.25,
                         drift=(0.004, 0.003, 0.0015), half_w: float = 8.0, half_h: float = 6.0, n_boxes: int = 6):
    """A robot circling the box-free middle of the room (heading tangent to the circle), lap after lap, so that the
    front end sees running-window matches, near-chain links and loop-closure candidates.  Odometry = the true motion
    composed with a small seeded drift per step.  Returns (world, true_poses[n,3], odom_poses[n,3], ranges[n,N])."""
    rng = np.random.default_rng(9_000_011 * (seed + 1))
    world = make_world(seed, half_w, half_h, n_boxes)
    true = np.zeros((n_poses, 3))
    odom = np.zeros((n_poses, 3))
    ranges = np.zeros((n_poses, laser.n_readings))
    dphi = step / radius
    for i in range(n_poses):
        phi = i * dphi
        true[i] = (radius * np.cos(phi), radius * np.sin(phi), (phi + np.pi / 2 + np.pi) % (2 * np.pi) - np.pi)
        ranges[i] = cast_scan(world, true[i], laser, rng, 0.01, 0.0)
        if i == 0:
            odom[i] = true[i]
        else:  # relative true motion in the previous true frame, perturbed, re-composed onto the previous odometry
            c, s = np.cos(true[i - 1, 2]), np.sin(true[i - 1, 2])
            d = true[i, :2] - true[i - 1, :2]
            local = np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1]]) + rng.normal(0, drift[:2])
            dth = (true[i, 2] - true[i - 1, 2] + np.pi) % (2 * np.pi) - np.pi + rng.normal(0, drift[2])
            co, so = np.cos(odom[i - 1, 2]), np.sin(odom[i - 1, 2])
            odom[i, 0] = odom[i - 1, 0] + co * local[0] - so * local[1]
            odom[i, 1] = odom[i - 1, 1] + so * local[0] + co * local[1]
            odom[i, 2] = (odom[i - 1, 2] + dth + np.pi) % (2 * np.pi) - np.pi
    return world, true, odom, ranges
