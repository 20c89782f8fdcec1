"""Pins the CPU oracle to the reference itself: oracle/_ref is /root/reference's own
HDLParser.cxx / TransformManager.cxx / TimeLine.h / type_defs.* / HDLFrame.cxx /
vtkPacketFileWriter.cxx compiled verbatim against stand-in third-party headers
(oracle/ref_shim).  Every comparison here is exact (bit-for-bit floats included).

Skipped when oracle/_ref is not built (no /root/reference and no prebuilt library); the
committed fixtures under tests/golden/ (generated from oracle/_ref by
tests/golden/make_golden.py) still pin the oracle in that case (test_golden.py).
"""
import os
import time

import numpy as np
import pytest

from oracle import ref
from oracle.oracle import Oracle
from veloslam_b200 import pcapio, synth

from helpers import make_packet

# vtkPacketFileWriter stamps records with mktime(): make "local time" UTC for the byte compare
os.environ["TZ"] = "UTC"
time.tzset()

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref is not built")


def both(calib, poses=None, sel=None, skip=0, crop=None):
    out = []
    for cls in (ref.RefParser, Oracle):
        o = cls()
        o.set_calibration(calib)
        if sel is not None:
            o.set_laser_selection(sel)
        o.set_points_skip(skip)
        if crop is not None:
            o.set_crop(1, crop[0], crop[1])
        if poses is not None:
            o.add_poses(poses[0], poses[1])
        out.append(o)
    return out


def assert_same_frames(r, o, force_split=True):
    if force_split:
        r.split_frame()
        o.split_frame()
    fr, fo = r.frames(), o.frames()
    assert len(fr) == len(fo)
    for a, b in zip(fr, fo):
        assert a.n_points == b.n_points
        assert a.timestamp_us == b.timestamp_us
        # HDLFrame::skips is an uninitialised uint8_t in frames that never get their meta
        if b.skips >= 0:
            assert a.skips == b.skips
        assert a.n_packets == b.n_packets
        assert np.array_equal(a.laser_counts, b.laser_counts)
        assert np.array_equal(a.xyzi.view(np.uint32), b.xyzi.view(np.uint32))
        assert np.array_equal(a.azimuth, b.azimuth)
        assert np.array_equal(a.distance.view(np.uint32), b.distance.view(np.uint32))
        assert np.array_equal(a.carpose_TRV.view(np.uint64), b.carpose_TRV.view(np.uint64))
        assert a.carpose_valid == b.carpose_valid
    assert r.state() == o.state()
    return len(fr)


def test_hdl64_decode_deskew():
    pk, t = synth.hdl64_packets(2500)
    r, o = both(synth.calib_hdl64(), synth.ins_trajectory(90))
    for p in (r, o):
        p.process_packets(synth.as_bytes(pk), t)
    assert assert_same_frames(r, o) == 8


def test_hdl32_with_mid_packet_wraps():
    pk, t = synth.hdl32_packets(2000, az0=123.0)
    r, o = both(synth.calib_hdl32(), synth.ins_trajectory(130))
    for p in (r, o):
        p.process_packets(synth.as_bytes(pk), t)
    assert assert_same_frames(r, o) >= 10


def test_vlp16_mode():
    pk, t = synth.hdl32_packets(300, seed=16)
    c = synth.calib_hdl32()
    c.n_enabled = 16
    r, o = both(c)
    for p in (r, o):
        p.process_packets(synth.as_bytes(pk), t)
    assert r.num_channels() == 16
    assert_same_frames(r, o)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_azimuths(seed):
    """Several wraps per packet, non-constant skip chains, frames that never get their meta."""
    pk, t = synth.random_packets(700, seed)
    r, o = both(synth.calib_hdl64(), synth.ins_trajectory(40))
    for p in (r, o):
        p.process_packets(synth.as_bytes(pk), t)
    assert assert_same_frames(r, o) > 100


def test_filters_and_crop():
    pk, t = synth.hdl64_packets(900)
    sel = np.ones(64, np.int32)
    sel[[1, 17, 40]] = 0
    region = (-20.0, 20.0, -30.0, 30.0, -5.0, 5.0)
    for inside in (0, 1):
        r, o = both(synth.calib_hdl64(), synth.ins_trajectory(40), sel=sel, skip=2,
                    crop=(inside, region))
        for p in (r, o):
            p.process_packets(synth.as_bytes(pk), t)
        assert_same_frames(r, o)


def test_wrong_size_packets_are_dropped():
    pk, t = synth.hdl64_packets(400)
    b = synth.as_bytes(pk)
    lengths = np.full(400, 1206, np.int32)
    lengths[[3, 77, 200]] = [512, 1205, 0]
    r, o = both(synth.calib_hdl64())
    r.process_packets(b, t, lengths)
    for i in range(400):
        o.process_packet(b[i], t[i], length=int(lengths[i]))
    assert_same_frames(r, o)


def test_identity_kats_hold_for_the_reference_too():
    r, _ = both(synth.calib_identity(64))
    d = np.zeros((12, 32), np.uint16)
    d[0, 3] = 5000
    r.process_packets(synth.as_bytes(make_packet(9000 + 10 * np.arange(12), None, d)),
                      np.array([synth.T0_US]))
    r.split_frame()
    f = r.frames()[0]
    assert np.allclose(f.xyzi[0, :3], (10.0, 0.0, 0.0), atol=2e-6)


def test_timeline_interpolation_everywhere():
    rng = np.random.default_rng(5)
    n = 300
    ts = synth.T0_US + np.cumsum(rng.integers(2_000, 30_000, n)).astype(np.int64)
    trv = rng.normal(size=(n, 9)) * 50
    r, o = both(synth.calib_identity(64), (ts, trv))
    q = np.concatenate([ts, ts[:-1] + 1, ts[1:] - 1, [ts[0] - 5_000, ts[-1] + 123_456],
                        rng.integers(ts[0] - 1000, ts[-1] + 1000, 500)])
    for t in q:
        a, b = r.interpolate(int(t)), o.interpolate(int(t))
        assert a[0] == b[0] and a[2] == b[2]
        assert np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64)), t


def test_timeline_out_of_order_and_duplicates():
    r, o = both(synth.calib_identity(64))
    rng = np.random.default_rng(8)
    order = rng.permutation(40)
    for k in list(order) + [5, 17, 39]:
        trv = rng.normal(size=(1, 9))
        for p in (r, o):
            p.add_poses([synth.T0_US + 10_000 * int(k)], trv)
    assert r.num_poses() == o.num_poses()
    for t in range(synth.T0_US - 20_000, synth.T0_US + 420_000, 3_333):
        a, b = r.interpolate(t), o.interpolate(t)
        assert a[0] == b[0] and np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64)), t


@pytest.mark.parametrize("n_poses", [0, 1])
def test_short_timelines(n_poses):
    pt, trv = synth.ins_trajectory(1)
    r, o = both(synth.calib_hdl64(), (pt[:n_poses], trv[:n_poses]))
    a, b = r.interpolate(synth.T0_US + 1_234_567), o.interpolate(synth.T0_US + 1_234_567)
    assert a[0] == b[0] and a[2] == b[2] and np.array_equal(a[1], b[1])
    pk, t = synth.hdl64_packets(500)
    for p in (r, o):
        p.process_packets(synth.as_bytes(pk), t)
    assert_same_frames(r, o)


def test_pose_matrix():
    rng = np.random.default_rng(2)
    for _ in range(50):
        trv = rng.uniform(-180, 180, 9)
        a, b = ref.RefParser.pose_matrix(trv), Oracle.pose_matrix(trv)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_offline_index_and_get_frame_through_a_pcap_file(tmp_path):
    pk, t = synth.hdl64_packets(1200)
    b = synth.as_bytes(pk)
    path = str(tmp_path / "20160701T000000.pcap")      # an ISO name: the reference renames others
    pcapio.write_pcap(path, b, t)
    calib = synth.calib_hdl64()
    r, o = both(calib, synth.ins_trajectory(50))
    pos, sk, ts = r.read_frame_information(path)
    sp, sk2, ts2 = Oracle.read_frame_information(b, t)
    assert np.array_equal(pos, pcapio.GLOBAL_HEADER_BYTES + sp.astype(np.int64) * pcapio.RECORD_BYTES)
    assert np.array_equal(sk, sk2) and np.array_equal(ts, ts2)
    for i in range(len(pos)):
        fr = r.get_frame(path, pos[i], sk[i])
        fo = o.get_frame(b, t, sp[i], sk2[i])
        assert fr.n_points == fo.n_points
        assert np.array_equal(fr.xyzi.view(np.uint32), fo.xyzi.view(np.uint32))
        assert np.array_equal(fr.laser_counts, fo.laser_counts)
    assert os.path.exists(path)


def test_pcap_writer_matches_the_reference_byte_for_byte(tmp_path):
    pk, t = synth.hdl64_packets(37)
    b = synth.as_bytes(pk)
    path = str(tmp_path / "ref.pcap")
    # the reference stamps mktime(to_tm(t)) (local time, UTC here); its reader adds 8 hours
    assert ref.RefParser.write_pcap(path, b, t - pcapio.TZ_SHIFT_US)
    theirs = np.fromfile(path, dtype=np.uint8)
    ours = pcapio.write_pcap_image(b, t)
    assert theirs.shape == ours.shape and np.array_equal(theirs, ours)
    back, tb = pcapio.read_pcap_image(theirs)
    assert np.array_equal(back, b) and np.array_equal(tb, t)


# --- geodesy, TimeSolver, INS record -> pose ------------------------------------
R = ref
ORIG_XYZ = (-2781621.9891904, 4672106.75052387, 18.8910392)     # INSSource.cxx:334


def _ins_records(n, seed=3):
    from oracle.oracle import INS_DTYPE
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=INS_DTYPE)
    r["message_id"] = 508
    r["week_number"] = 1903
    r["milliseconds"] = 345_600_000 + 10 * np.arange(n)
    r["week_number_pos"] = 1903
    r["seconds_pos"] = 345_600.0 + 0.01 * np.arange(n) + rng.uniform(0, 0.004, n)
    r["LLH"][:, 0] = 39.8569901 + np.cumsum(rng.normal(0, 1e-6, n))
    r["LLH"][:, 1] = 116.1736406 + np.cumsum(rng.normal(0, 1e-6, n))
    r["LLH"][:, 2] = 89.09 + rng.normal(0, 0.05, n)
    r["V"] = rng.normal(0, 5, (n, 3))
    r["Eulr"] = rng.uniform(-180, 180, (n, 3))
    return r


def test_geodesy_matches_reference_bit_for_bit():
    from oracle import oracle as O
    assert R.sizeof_inspva() == O.INS_DTYPE.itemsize
    rng = np.random.default_rng(11)
    org = O.llh2xyz(np.radians([39.8569901, 116.1736406, 0]) + [0, 0, 89.09])
    for _ in range(200):
        llh = [np.radians(rng.uniform(-80, 80)), np.radians(rng.uniform(-179, 179)),
               rng.uniform(-100, 9000)]
        assert np.array_equal(O.llh2xyz(llh), R.llh2xyz(llh))
        for o in (ORIG_XYZ, org, O.llh2xyz(llh)):
            assert np.array_equal(O.llh2enu(llh, o), R.llh2enu(llh, o))
    # round trip sanity: a point 100 m east of the origin
    here = np.radians([39.8569901, 116.1736406, 0]) + [0, 0, 89.09]
    enu = O.llh2enu(here + [0, 100 / 6378137.0 / np.cos(here[0]), 0], O.llh2xyz(here))
    assert abs(enu[0] - 100) < 0.5 and abs(enu[1]) < 0.01 and abs(enu[2]) < 0.01


def test_time_solver_hdl_matches_reference():
    from oracle import oracle as O
    now0 = 1_467_331_234_567_890
    rng = np.random.default_rng(5)
    # two hour wraps, a duplicate, and a backwards step that the reference counts as a wrap
    gps = np.concatenate([np.arange(3_599_000_000, 3_600_000_000, 137_000),
                          np.arange(500, 2_000_000, 211_000), [2_000_000, 1_999_999],
                          np.arange(3_599_900_000, 3_600_000_000, 33_000), [7, 7, 8]]).astype(np.uint32)
    ref = R.RefTimeSolver(now0)
    mine = O.TimeSolver()
    for i, g in enumerate(gps):
        now = now0 + 553 * i + int(rng.integers(0, 50))     # the clock only matters at packet 0
        assert mine.hdl(g, now) == ref.hdl(g, now), i


def test_time_solver_ins_and_pose_match_reference():
    from oracle import oracle as O
    recs = _ins_records(300)
    recs["week_number_pos"][100:] += 1          # week roll-over between send and pose time
    now = 1_467_331_200_000_000 + 10_000 * np.arange(len(recs))
    ref = R.RefTimeSolver(int(now[0]))
    want = np.array([ref.ins(recs.ctypes.data + i * recs.dtype.itemsize, int(now[i]))
                     for i in range(len(recs))])
    assert np.array_equal(O.ins_times(recs, now), want)
    trv = O.ins_poses(recs, ORIG_XYZ)
    for i in (0, 1, 150, 299):
        llh = [np.radians(1.0) * 0 + recs["LLH"][i, 0] * np.pi / 180, recs["LLH"][i, 1] * np.pi / 180,
               recs["LLH"][i, 2]]
        assert np.array_equal(trv[i, :3], R.llh2enu(llh, ORIG_XYZ))
    assert np.array_equal(trv[:, 3:6], recs["Eulr"]) and np.array_equal(trv[:, 6:], recs["V"])
