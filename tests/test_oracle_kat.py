"""Known-answer tests pinning the CPU oracle to values derived by hand from the reference
source.  The reference ships no tests or golden vectors for this
path, so these hand-derived answers (and, where built, oracle/_ref -- see
test_oracle_vs_ref.py) are what the oracle is pinned to.
"""
import math

import numpy as np
import pytest

from oracle.oracle import Oracle
from veloslam_b200 import synth

from helpers import make_packet, single_return_packet

T0 = synth.T0_US


def decode_one(calib, pk, t=T0):
    o = Oracle()
    o.set_calibration(calib)
    o.trace_enable()
    o.process_packets(synth.as_bytes(pk), np.array([t], dtype=np.int64))
    return o, o.trace()


# --- decode + calibration (HDLParser.cxx:587-639) ------------------------------------------
@pytest.mark.parametrize("az,expect", [(0, (0.0, 10.0, 0.0)), (9000, (10.0, 0.0, 0.0)),
                                       (18000, (0.0, -10.0, 0.0)), (27000, (-10.0, 0.0, 0.0))])
def test_identity_calibration_axes(az, expect):
    _, tr = decode_one(synth.calib_identity(64), single_return_packet(az, 3, 5000))
    assert len(tr["x"]) == 1
    got = (tr["x"][0], tr["y"][0], tr["z"][0])
    assert np.allclose(got, expect, atol=2e-6)
    assert tr["azimuth"][0] == az and tr["distance"][0] == 5000 and tr["laser"][0] == 3
    assert tr["intensity"][0] == 7


def test_vertical_correction_30deg():
    c = synth.calib_identity(64)
    c.vert_deg[5] = 30.0
    _, tr = decode_one(c, single_return_packet(0, 5, 5000))
    assert np.allclose((tr["x"][0], tr["y"][0], tr["z"][0]), (0.0, 8.660254, 5.0), atol=1e-6)


def test_distance_correction_adds_a_metre():
    c = synth.calib_identity(64)
    c.dist_cm[2] = 100.0
    _, tr = decode_one(c, single_return_packet(0, 2, 5000))
    assert np.allclose((tr["x"][0], tr["y"][0], tr["z"][0]), (0.0, 11.0, 0.0), atol=1e-6)


def test_horizontal_and_vertical_offsets():
    c = synth.calib_identity(64)
    c.hoff_cm[0] = 10.0
    c.voff_cm[0] = 20.0
    _, tr = decode_one(c, single_return_packet(0, 0, 5000))
    assert np.allclose((tr["x"][0], tr["y"][0], tr["z"][0]), (-0.1, 10.0, 0.2), atol=1e-6)


def test_rot_correction_uses_libm_branch():
    c = synth.calib_identity(64)
    c.rot_deg[1] = 90.0          # az/100 - 90 deg
    _, tr = decode_one(c, single_return_packet(9000, 1, 5000))
    assert np.allclose((tr["x"][0], tr["y"][0]), (0.0, 10.0), atol=1e-6)


def test_zero_distance_emits_nothing():
    _, tr = decode_one(synth.calib_identity(64), single_return_packet(100, 4, 0))
    assert len(tr["x"]) == 0


def test_upper_block_offsets_laser_by_32():
    _, tr = decode_one(synth.calib_identity(64), single_return_packet(0, 7, 1000, upper=True))
    assert tr["laser"][0] == 39


def test_wrong_size_packet_is_dropped():
    o = Oracle()
    o.set_calibration(synth.calib_identity(64))
    o.trace_enable()
    pk = synth.as_bytes(single_return_packet(0, 7, 1000))
    o.process_packet(pk[0, :1205], T0)
    assert len(o.trace()["x"]) == 0 and o.state()["last_azimuth"] == -1


def test_laser_selection_and_points_skip():
    pk, t = synth.hdl64_packets(1, zero_frac=0.0)
    o = Oracle()
    o.set_calibration(synth.calib_identity(64))
    sel = np.ones(64, np.int32)
    sel[[3, 40]] = 0
    o.set_laser_selection(sel)
    o.set_points_skip(2)                 # blocks 0,3,6,9 only
    o.trace_enable()
    o.process_packets(synth.as_bytes(pk), t)
    tr = o.trace()
    assert set(np.unique(tr["block"])) == {0, 3, 6, 9}
    assert 3 not in tr["laser"] and 40 not in tr["laser"]
    assert len(tr["x"]) == 4 * 31


def test_crop_flag_semantics_are_inverted():
    """cropInside=false + crop on removes points INSIDE the box (HDLParser.cxx:629-639)."""
    c = synth.calib_identity(64)
    region = [-1.0, 1.0, 9.0, 11.0, -1.0, 1.0]       # contains (0, 10, 0)
    for crop_inside, n_expected in ((0, 0), (1, 1)):
        o = Oracle()
        o.set_calibration(c)
        o.set_crop(1, crop_inside, region)
        o.trace_enable()
        o.process_packets(synth.as_bytes(single_return_packet(0, 3, 5000)), np.array([T0]))
        assert len(o.trace()["x"]) == n_expected
    for crop_inside, n_expected in ((0, 1), (1, 0)):  # a point outside the box
        o = Oracle()
        o.set_calibration(c)
        o.set_crop(1, crop_inside, region)
        o.trace_enable()
        o.process_packets(synth.as_bytes(single_return_packet(9000, 3, 5000)), np.array([T0]))
        assert len(o.trace()["x"]) == n_expected


# --- HDL-32 per-return adjustment (HDLParser.cxx:946-962, 133-137) ---------------------------
def test_hdl32_time_and_azimuth_adjustment():
    c = synth.calib_hdl32()
    az = (1000 + 20 * np.arange(12)) % 36000            # azimuthDiff = 20
    d = np.zeros((12, 32), np.uint16)
    d[3, 16] = 4000
    d[0, 31] = 4000
    _, tr = decode_one(c, make_packet(az, None, d))
    # emission order: block 0 first
    assert tr["tadj_us"][0] == round(0 * 46.08 + 31 * 1.152)        # 36
    assert tr["tadj_us"][1] == 157                                   # round(156.672)
    assert tr["azimuth"][1] == 1060 + 8                              # round(20 * 0.4)
    assert tr["azimuth"][0] == 1000 + round(20 * 31 * 1.152 / 46.08)


def test_median_is_seventh_smallest_of_eleven():
    c = synth.calib_hdl32()
    diffs = np.array([5, 5, 5, 5, 5, 5, 40, 40, 40, 40, 40])   # sorted -> index 6 == 40
    az = np.concatenate([[100], 100 + np.cumsum(diffs)])
    d = np.zeros((12, 32), np.uint16)
    d[0, 20] = 3000
    _, tr = decode_one(c, make_packet(az, None, d))
    assert tr["azimuth"][0] == 100 + round(40 * 20 * 1.152 / 46.08)  # 120
    diffs2 = np.array([5, 5, 5, 5, 5, 5, 5, 40, 40, 40, 40])   # index 6 == 5
    az2 = np.concatenate([[100], 100 + np.cumsum(diffs2)])
    _, tr2 = decode_one(c, make_packet(az2, None, d))
    assert tr2["azimuth"][0] == 103   # 5 * 0.5 = 2.5 -> 3: std::round is half away from zero


def test_hdl64_gets_no_adjustment():
    pk, t = synth.hdl64_packets(2, zero_frac=0.0)
    o = Oracle()
    o.set_calibration(synth.calib_hdl64())
    o.trace_enable()
    o.process_packets(synth.as_bytes(pk), t)
    tr = o.trace()
    assert np.all(tr["tadj_us"] == 0)
    az = pk["blocks"]["azimuth"]
    assert np.array_equal(tr["azimuth"], az[tr["packet"], tr["block"]] % 36000)


# --- pose interpolation (TransformManager.cxx:149-177, type_defs.h:102-146) ------------------
def _two_pose_oracle(R0, R1, T0v=(0, 0, 0), T1v=(0, 0, 0)):
    o = Oracle()
    o.set_calibration(synth.calib_identity(64))
    trv = np.zeros((2, 9))
    trv[0, 0:3], trv[1, 0:3] = T0v, T1v
    trv[0, 3:6], trv[1, 3:6] = R0, R1
    o.add_poses([T0, T0 + 10_000], trv)
    return o


def test_yaw_90_is_rz():
    m = Oracle.pose_matrix([0, 0, 0, 0, 0, 90, 0, 0, 0])
    assert np.allclose(m[:, :3] @ np.array([1.0, 0, 0]), [0, 1, 0], atol=1e-15)
    m = Oracle.pose_matrix([1, 2, 3, 90, 0, 0, 0, 0, 0])   # R[0] about Y
    assert np.allclose(m[:, :3] @ np.array([0, 0, 1.0]), [1, 0, 0], atol=1e-15)
    assert np.allclose(m[:, 3], [1, 2, 3])
    m = Oracle.pose_matrix([0, 0, 0, 0, 90, 0, 0, 0, 0])   # R[1] about X
    assert np.allclose(m[:, :3] @ np.array([0, 1.0, 0]), [0, 0, 1], atol=1e-15)


def test_rotation_order_is_ry_rx_rz():
    a, b, c = 10.0, 20.0, 30.0
    m = Oracle.pose_matrix([0, 0, 0, a, b, c, 0, 0, 0])[:, :3]

    def rot(axis, deg):
        r, s, co = np.eye(3), math.sin(math.radians(deg)), math.cos(math.radians(deg))
        i, j = [(1, 2), (2, 0), (0, 1)][axis]
        r[i, i], r[i, j], r[j, i], r[j, j] = co, -s, s, co
        return r
    assert np.allclose(m, rot(1, a) @ rot(0, b) @ rot(2, c), atol=1e-15)


def test_euler_lerp_goes_the_long_way_round():
    """170 -> -170 deg yaw lerps through 0, not 180 (F2: linear on Euler angles)."""
    o = _two_pose_oracle((0, 0, 170), (0, 0, -170))
    ok, trv, sp = o.interpolate(T0 + 5_000)
    assert ok and sp == 0 and trv[5] == 0.0


def test_interpolation_extrapolates_and_rebases():
    o = _two_pose_oracle((0, 0, 0), (0, 0, 0), (0, 0, 0), (1, 2, 3))
    ok, trv, _ = o.interpolate(T0 + 20_000)           # ratio 2
    assert np.allclose(trv[:3], (2, 4, 6))
    ok, trv, _ = o.interpolate(T0 - 10_000)           # ratio -1
    assert np.allclose(trv[:3], (-1, -2, -3))
    # decode: first packet defines the frame origin -> translation 0 for itself, relative after
    o.trace_enable()
    pk = np.concatenate([single_return_packet(0, 3, 5000), single_return_packet(200, 3, 5000)])
    o.process_packets(synth.as_bytes(pk), np.array([T0, T0 + 5_000]))
    tr = o.trace()
    assert np.allclose((tr["x"][0], tr["y"][0], tr["z"][0]), (0, 10, 0), atol=1e-6)
    s, c = math.sin(math.radians(2.0)), math.cos(math.radians(2.0))
    assert np.allclose((tr["x"][1], tr["y"][1], tr["z"][1]),
                       (10 * s + 0.5, 10 * c + 1.0, 1.5), atol=1e-5)


def test_empty_and_single_pose_timeline_give_no_transform():
    o = Oracle()
    o.set_calibration(synth.calib_identity(64))
    assert o.interpolate(T0)[0] is False
    o.add_poses([T0], np.array([[5, 6, 7, 0, 0, 90, 1, 1, 1.0]]))
    ok, trv, sp = o.interpolate(T0 + 2_000_000)
    assert ok and sp == -1                            # stays "invalid" for the parser
    assert np.allclose(trv[:3], (7, 8, 9))            # T + V * 2 s
    o.trace_enable()
    o.process_packets(synth.as_bytes(single_return_packet(0, 3, 5000)), np.array([T0]))
    tr = o.trace()
    assert np.allclose((tr["x"][0], tr["y"][0]), (0, 10), atol=1e-6)   # untransformed


def test_timeline_bracket_is_clamped_lower_bound():
    """Net semantics of TimeLine::getBoundaryData: i = clamp(lower_bound(ts, t), 1, N-1)."""
    rng = np.random.default_rng(3)
    n = 200
    ts = T0 + np.cumsum(rng.integers(5_000, 15_000, n)).astype(np.int64)
    trv = np.zeros((n, 9))
    trv[:, 0] = np.arange(n)                          # T[0] = index -> lerp reveals the bracket
    o = Oracle()
    o.add_poses(ts, trv)
    q = np.concatenate([ts[[0, 1, 5, 100, n - 2, n - 1]], ts[:-1] + 1, ts[1:] - 1,
                        [ts[0] - 7_000, ts[-1] + 9_000], rng.integers(ts[0], ts[-1], 300)])
    for t in q:
        i = int(np.clip(np.searchsorted(ts, t, side="left"), 1, n - 1))
        ratio = float(t - ts[i - 1]) / float(ts[i] - ts[i - 1])
        expect = (i - 1) + ratio
        ok, got, _ = o.interpolate(int(t))
        assert ok and abs(got[0] - expect) < 1e-9, (t, got[0], expect)


def test_timeline_out_of_order_insert_and_overwrite():
    o = Oracle()
    trv = np.zeros((1, 9))
    order = [0, 3, 1, 2, 6, 5, 4, 7, 8, 9, 10, 12, 11, 13]
    for k in order:
        trv[0, 0] = k
        o.add_poses([T0 + 10_000 * k], trv)
    trv[0, 0] = 100.0
    o.add_poses([T0 + 10_000 * 13], trv)              # duplicate timestamp overwrites
    assert o.num_poses() == 14
    ok, got, _ = o.interpolate(T0 + 10_000 * 11 + 5_000)
    assert abs(got[0] - 11.5) < 1e-12
    ok, got, _ = o.interpolate(T0 + 10_000 * 13)
    assert abs(got[0] - 100.0) < 1e-12


# --- segmentation (HDLParser.cxx:1013-1054, 867-897) ----------------------------------------
def _ramp_packets(start_az, n, step=100):
    az = (start_az + step * np.arange(12 * n)) % 36000
    d = np.full((12, 32), 1000, np.uint16)
    return np.concatenate([make_packet(az[12 * i:12 * i + 12], None, d) for i in range(n)])


def test_split_kat_streaming_quirks():
    """Wrap at block 5 of packet 1: frame 0 gets blocks <=4 of P1; frame 1 gets blocks 5-11 of
    P1 and 5-11 of P2 (blocks 0-4 of P2 silently dropped); its skips = 5 and its timestamp is
    the time of P2, not P1 (F4 a/b)."""
    start = (36000 - 100 * (12 + 5)) % 36000        # block 5 of packet 1 lands on azimuth 0
    pk = _ramp_packets(start, 4)
    az = pk["blocks"]["azimuth"]
    assert az[1, 5] == 0 and az[1, 4] == 35900
    t = T0 + 1000 * np.arange(4, dtype=np.int64)
    o = Oracle()
    o.set_calibration(synth.calib_identity(32, 32))
    o.trace_enable()
    o.process_packets(synth.as_bytes(pk), t)
    tr = o.trace()
    assert o.num_frames() == 1
    f0 = o.frame(0)
    assert f0.n_points == (12 + 5) * 32 and f0.skips == 0 and f0.timestamp_us == t[0]
    assert f0.n_packets == 3                           # first packet doubled (F4d)
    sel = tr["frame"] == 1
    blocks = set(zip(tr["packet"][sel].tolist(), tr["block"][sel].tolist()))
    expect = {(1, b) for b in range(5, 12)} | {(2, b) for b in range(5, 12)} | \
             {(3, b) for b in range(12)}
    assert blocks == expect
    assert o.state() == {"last_azimuth": int(az[3, 11]), "firing_skip": 0,
                         "frame_meta_inited": True, "is_hdl64": False}
    o.split_frame()
    f1 = o.frame(1)
    assert f1.skips == 5 and f1.timestamp_us == t[2]


def test_first_block_never_wraps_and_last_partial_frame_is_not_emitted():
    pk, t = synth.hdl64_packets(800)
    o = Oracle()
    o.set_calibration(synth.calib_hdl64())
    o.process_packets(synth.as_bytes(pk), t)
    az = pk["blocks"]["azimuth"].reshape(-1).astype(np.int64)
    assert o.num_frames() == int(np.sum(az[1:] < az[:-1])) == 2
    assert o.open_frame_points() > 0


def test_hdl64_beam_lut_reorders_lasers_at_split():
    pk, t = synth.hdl64_packets(400, zero_frac=0.0)
    o = Oracle()
    o.set_calibration(synth.calib_identity(64))
    sel = np.zeros(64, np.int32)
    sel[[0, 38]] = 1
    o.set_laser_selection(sel)
    o.process_packets(synth.as_bytes(pk), t)
    f = o.frame(0)
    assert f.is_hdl64_order
    nz = np.nonzero(f.laser_counts)[0]
    # new[i] = old[LUT[i]]: raw laser 38 -> row 0, raw laser 0 -> row 36
    assert list(nz) == [0, 36]


def test_offline_index_and_get_frame_have_no_drops():
    pk, t = synth.hdl64_packets(1100)
    b = synth.as_bytes(pk)
    sp, sk, ts = Oracle.read_frame_information(b, t)
    az = pk["blocks"]["azimuth"].reshape(-1).astype(np.int64)
    wraps = np.nonzero(az[1:] < az[:-1])[0] + 1
    assert len(sp) == len(wraps) + 1 and sp[0] == 0 and sk[0] == 0 and ts[0] == t[0]
    assert np.array_equal(sp[1:], wraps // 12) and np.array_equal(sk[1:], wraps % 12)
    assert np.array_equal(ts[1:], t[wraps // 12])
    o = Oracle()
    o.set_calibration(synth.calib_hdl64())
    f1 = o.get_frame(b, t, sp[1], sk[1])
    d = pk["blocks"]["returns"]["distance"].reshape(-1, 32)
    assert f1.n_points == int(np.count_nonzero(d[wraps[0]:wraps[1]]))
    assert f1.timestamp_us == t[sp[1]] and f1.skips == sk[1]
    last = o.get_frame(b, t, sp[-1], sk[-1])            # forced split at end of data
    assert last.n_points == int(np.count_nonzero(d[wraps[-1]:]))


def test_config1_hdl32_minute_on_the_oracle():
    """BASELINE.json configs[0], the reference's own CPU-runnable case, at full size: 60 s of
    HDL-32E at 10 Hz: 108 480 packets, 41.66 M return slots, 600 +- 1 frames."""
    from veloslam_b200 import synth
    n = 108_480
    pk, t = synth.hdl32_packets(n)
    o = Oracle()
    o.set_calibration(synth.calib_hdl32())
    o.process_packets(synth.as_bytes(pk), t)
    assert n * 384 == 41_656_320
    assert 599 <= o.num_frames() <= 601
    d = pk["blocks"]["returns"]["distance"]
    frames = o.frames()
    total = sum(f.n_points for f in frames) + o.open_frame_points()
    # streaming drops the blocks of the packet after a wrap that precede the wrap block (F4a)
    assert 0 <= int(np.count_nonzero(d)) - total < 600 * 11 * 32
    assert all(f.n_lasers == 32 for f in frames)
