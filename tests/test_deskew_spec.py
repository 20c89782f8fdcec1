"""CPU check of the per-point deskew extension's two statements against each other: the numpy
quaternion-slerp statement (oracle/deskew_port.py) and the independent 50-digit SO(3)-geodesic
evaluator (tests/mp_deskew.py) must describe the same function (N4 has no reference behaviour to
pin to; the GPU path is checked against both in tests/test_gpu_deskew.py)."""
import numpy as np

from oracle import deskew_port as D
from veloslam_b200 import synth

import mp_deskew as M


def test_geodesic_and_slerp_statements_agree():
    poses = synth.ins_trajectory(200, yaw_amp_deg=80.0, yaw_period_s=6.0, speed=20.0)
    rng = np.random.default_rng(1)
    n = 24
    xyz = rng.uniform(-80, 80, (n, 3))
    t = synth.T0_US + np.sort(rng.integers(-30_000, 2_050_000, n)).astype(np.int64)  # incl. extrapolation
    off = rng.integers(0, 600, n)
    org = t - rng.integers(0, 90_000, n)
    want = D.deskew_points(xyz, np.arange(n), off, t, org, *poses)
    tl = M.Timeline(*poses)
    for i in range(n):
        w = M.deskew_point(tl, xyz[i], int(t[i]), int(off[i]), int(org[i]))
        assert np.abs(np.array(w) - want[i]).max() < 1e-9


def test_yaw_wrap_takes_the_shorter_arc_in_both_statements():
    pt = synth.T0_US + 10_000 * np.arange(4, dtype=np.int64)
    trv = np.zeros((4, 9))
    trv[:, 5] = [170.0, -170.0, 170.0, -170.0]      # 20 degrees through +-180, not 340 the long way
    tl = M.Timeline(pt, trv)
    p = np.array([[10.0, 0.0, 0.0]])
    t = np.array([pt[0] + 5000])
    got = M.deskew_point(tl, p[0], int(t[0]), 0, int(pt[0]))
    want = D.deskew_points(p, [0], [0], t, np.array([pt[0]]), pt, trv)[0]
    assert np.abs(np.array(got) - want).max() < 1e-9
    # midway between yaw 170 and -170 is yaw 180: relative to the origin pose (yaw 170) a 10 degree turn
    assert abs(np.hypot(got[0], got[1]) - 10.0) < 1e-9 and abs(np.degrees(np.arctan2(got[1], got[0])) - 10.0) < 1e-6
