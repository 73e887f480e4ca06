"""Shared workload of the lesson6 front-end (mapper) tests: a robot circling a room, lap after lap, with drifting
odometry; small matcher windows keep the CPU reference fast.  Used by the CPU host-logic tests, the GPU tests and
tests/golden/make_golden.py."""
import numpy as np

KW = dict(scan_buffer_size=20, link_match_minimum_response_fine=0.1, link_scan_maximum_distance=1.5,
          loop_search_maximum_distance=3.0, loop_match_minimum_chain_size=5, loop_match_maximum_variance_coarse=9.0,
          loop_match_minimum_response_coarse=0.35, loop_match_minimum_response_fine=0.45,
          sequential_search_size=0.5, sequential_resolution=0.05, loop_search_size=4.0)


def workload(pkg, seed, n, drift=(0.004, 0.003, 0.0015)):
    laser = pkg.synth.Laser(range_threshold=9.25)
    world, true, odom, ranges = pkg.synth.make_loop_trajectory(seed, n, laser, radius=2.0, step=0.25, drift=drift)
    prm = pkg.load("mapper").default_params(9.25, **KW)
    return laser, prm, true, odom, ranges


def run(mapper, odom, ranges, dt=0.1):
    flags, firsts = [], []
    for i in range(len(ranges)):
        ok, c = mapper.process(ranges[i], odom[i], dt * i)
        flags.append(ok)
        firsts.append(c)
    return np.array(flags), np.array(firsts)


FLEET_DRIFT = (0.002, 0.0015, 0.0008)


def fleet_events(n_per_robot=50, starts=(0, 50, 100), stagger=8):
    """Three robots on the same lap trajectory, one lap apart, entering `stagger` steps after each other:
    [(robot, frame)] in processing order."""
    ev = []
    for t in range(n_per_robot + stagger * (len(starts) - 1)):
        for r, s0 in enumerate(starts):
            k = t - stagger * r
            if 0 <= k < n_per_robot:
                ev.append((r, s0 + k))
    return ev


def run_fleet(mapper, sensors, odom, ranges, dt=0.1):
    """sensors[robot] = what the mapper's process() takes as `sensor`."""
    flags, firsts = [], []
    for k, (rob, i) in enumerate(fleet_events()):
        ok, c = mapper.process(ranges[i], odom[i], dt * k, sensor=sensors[rob])
        flags.append(ok)
        firsts.append(c)
    return np.array(flags), np.array(firsts)
