"""Host-side logic of the C++ facade (no GPU): TransformManager/TimeLine interpolation, the
pcap reader/writer, readFrameInformation and the calibration XML reader, against the oracle,
the reference-made golden fixtures and (when built) the reference itself."""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from veloslam_b200 import calibxml, pcapio, synth

import facade_util as F

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def _driver():
    F.build()


def test_transform_manager_matches_reference_table(tmp_path):
    g = np.load(os.path.join(GOLDEN, "interpolate.npz"))
    F.write_poses(tmp_path / "poses.bin", g["pose_t"], g["pose_trv"])
    g["query"].astype("<i8").tofile(tmp_path / "q.bin")
    r = F.run(["interp", tmp_path / "poses.bin", tmp_path / "q.bin", tmp_path / "out.bin"])
    assert r.returncode == 0, r.stderr
    rec = np.fromfile(tmp_path / "out.bin", dtype=[("v", "<f8", (9,)), ("found", "<i4"), ("valid", "<i4")])
    assert np.all(rec["found"] == 1) and np.all(rec["valid"] == 1)
    # identical to the reference except on exact hits at a bucket start (<= 1 ulp)
    assert np.allclose(rec["v"], g["result"], rtol=0, atol=1e-12)
    exact = np.mean(rec["v"].view(np.uint64) == g["result"].view(np.uint64))
    assert exact > 0.95


def test_transform_manager_short_timelines(tmp_path):
    pt, trv = synth.ins_trajectory(1)
    q = np.array([synth.T0_US + 2_000_000], dtype="<i8")
    q.tofile(tmp_path / "q.bin")
    for n in (0, 1):
        F.write_poses(tmp_path / "poses.bin", pt[:n], trv[:n])
        assert F.run(["interp", tmp_path / "poses.bin", tmp_path / "q.bin", tmp_path / "o.bin"]).returncode == 0
        rec = np.fromfile(tmp_path / "o.bin", dtype=[("v", "<f8", (9,)), ("found", "<i4"), ("valid", "<i4")])
        o = Oracle()
        o.add_poses(pt[:n], trv[:n])
        ok, want, sp = o.interpolate(int(q[0]))
        assert bool(rec["found"][0]) == ok and bool(rec["valid"][0]) == (sp != -1)
        assert np.array_equal(rec["v"][0], want)


def test_pcap_writer_and_offline_index(tmp_path):
    pk, t = synth.hdl64_packets(900)
    b = synth.as_bytes(pk)
    b.tofile(tmp_path / "pk.bin")
    t.astype("<i8").tofile(tmp_path / "t.bin")
    path = tmp_path / "20160701T000000.pcap"
    assert F.run(["writepcap", tmp_path / "pk.bin", tmp_path / "t.bin", path]).returncode == 0
    img = np.fromfile(path, dtype=np.uint8)
    assert np.array_equal(img, pcapio.write_pcap_image(b, t))     # == the reference writer's bytes
    assert F.run(["index", path, tmp_path / "idx.bin"]).returncode == 0
    raw = open(tmp_path / "idx.bin", "rb").read()
    n = int(np.frombuffer(raw, "<i4", 1)[0])
    rec = np.frombuffer(raw, dtype=[("pos", "<i8"), ("sk", "<i4"), ("ts", "<i8")], count=n, offset=4)
    sp, sk, ts = Oracle.read_frame_information(b, t)
    assert np.array_equal(rec["pos"], 24 + sp.astype(np.int64) * 1264)
    assert np.array_equal(rec["sk"], sk) and np.array_equal(rec["ts"], ts)
    assert os.path.exists(path)


def test_index_renames_files_not_named_after_their_first_packet(tmp_path):
    pk, t = synth.hdl64_packets(20)
    path = tmp_path / "recording.pcap"
    pcapio.write_pcap(str(path), synth.as_bytes(pk), t)
    assert F.run(["index", path, tmp_path / "idx.bin"]).returncode == 0
    assert not os.path.exists(path)
    assert os.path.exists(tmp_path / "20160701T000000.pcap")     # T0_US as an ISO string


def test_calibration_xml_reader(tmp_path):
    calib = synth.calib_hdl64()
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    assert F.run(["calib", tmp_path / "db.xml", tmp_path / "c.bin"]).returncode == 0
    raw = open(tmp_path / "c.bin", "rb").read()
    assert int(np.frombuffer(raw, "<i4", 1)[0]) == 64
    assert np.array_equal(np.frombuffer(raw, "<f8", 64, 4), calib.vert_deg)


# --- HDLManager host logic: TimeLine queries, cache, hard-drive buffers, meta ------
def test_hdlmanager_host_logic(tmp_path):
    d = tmp_path / "buf"
    d.mkdir()
    r = F.run(["manager_host", d, tmp_path / "report.txt"])
    assert r.returncode == 0, r.stderr
    rep = [l.split() for l in open(tmp_path / "report.txt")]
    kv = {l[0]: l[1:] for l in rep if l[0] not in ("disk", "meta")}
    assert kv["frames"] == ["10"]
    assert kv["at3"] == ["3"] and kv["near4"] == ["4"] and kv["at_missing"] == ["0"]
    assert kv["range"] == ["4"]            # [a, b] inclusive on both ends (HDLManager.h:148)
    assert kv["recent"] == ["9"]
    # cache of 4 (HDLManager.cxx:400-421): the held frame survives and keeps its cache slot
    assert kv["cleared"] == ["6"] and kv["held1_alive"] == ["1"]
    assert kv["cleared_not_on_disk"] == ["0"]
    # hard-drive buffers of 3 frames x 3 packets: file names = first packet time, manual fpos
    disk = [l for l in rep if l[0] == "disk"]
    assert [int(l[3]) for l in disk] == [24, 24 + 3 * 1264, 24 + 6 * 1264] * 2 + [24]
    assert all(l[2] == "1" for l in disk)
    assert [l[4] for l in disk] == ["20160701T000000"] * 3 + ["20160701T000000.300000"] * 3 + \
        ["20160701T000000.600000"]
    # the files are what vtkPacketFileWriter writes: header + 1264-byte records, payload intact
    img = np.fromfile(d / "20160701T000000.pcap", dtype=np.uint8)
    assert img.size == 24 + 9 * 1264
    rec = img[24:].reshape(9, 1264)
    assert all(np.all(rec[k, 58:] == k) for k in range(9))
    # .hdlmeta round trip into a fresh manager
    assert kv["savemeta"] == ["1"] and kv["loadmeta"] == ["1", "7"]
    meta = [l for l in rep if l[0] == "meta"]
    assert [int(l[2]) for l in meta] == [synth.T0_US - synth.T0_US + 1467331200000000 + 100000 * i
                                         for i in range(7)]
    assert [int(l[3]) for l in meta] == [int(l[3]) for l in disk]
    assert all(l[5] == "1" for l in meta)


# --- facade TimeSolver / CoordiTran against the oracle, bit for bit -----------------
def test_facade_time_solver_and_geodesy(tmp_path):
    from oracle import oracle as O
    rng = np.random.default_rng(8)
    gps = ((3_599_500_000 + 288 * np.arange(4000, dtype=np.int64)) % 3_600_000_000).astype("<u4")
    gps[1234] = gps[1233] - 5
    gps.tofile(tmp_path / "gps.bin")
    recs = np.zeros(64, dtype=O.INS_DTYPE)
    recs["week_number"] = 1903
    recs["week_number_pos"] = 1903
    recs["milliseconds"] = 345_600_000 + 10 * np.arange(64)
    recs["seconds_pos"] = 345_600.0 + 0.01 * np.arange(64) + rng.uniform(0, 0.004, 64)
    recs["LLH"] = np.array([39.8569901, 116.1736406, 89.09]) + rng.normal(0, 1e-4, (64, 3))
    recs.tofile(tmp_path / "ins.bin")
    r = F.run(["frontend", tmp_path / "gps.bin", tmp_path / "ins.bin", tmp_path / "out.bin"])
    assert r.returncode == 0, r.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    t_pk = np.frombuffer(raw, "<i8", len(gps))
    # only the first packet reads the clock (TimeSolver.cxx:35-42)
    assert np.array_equal(t_pk, O.TimeSolver().hdl_many(gps, 1467331234567890))
    rec = np.frombuffer(raw, dtype=[("t", "<i8"), ("enu", "<f8", (3,))], offset=8 * len(gps))
    arrival = 1467331200000000 + 10000 * np.arange(64)
    assert np.array_equal(rec["t"], O.ins_times(recs, arrival))
    want = O.ins_poses(recs, (-2781621.9891904, 4672106.75052387, 18.8910392))[:, :3]
    assert np.array_equal(rec["enu"], want)


# --- HDLSource receive path: UDP -> ring -> TimeSolver stamp -> consumer -------------
def test_hdlsource_udp_receive_ring(tmp_path):
    from oracle import oracle as O
    n = 600
    pk, _ = synth.hdl64_packets(n, seed=21)
    gps = ((3_599_950_000 + 288 * np.arange(n, dtype=np.int64)) % 3_600_000_000).astype(np.uint32)
    pk["gps"] = gps                               # the sensor clock wraps the hour mid-stream
    b = synth.as_bytes(pk)
    feed = [row for row in b]
    feed.insert(100, np.zeros(512, dtype=np.uint8))   # a position packet: ignored (length != 1206)
    port = F.free_udp_port()
    rc, err = F.run_with_udp_feed(["udp_host", port, n, tmp_path / "out.bin"], port, feed)
    assert rc == 0, err
    assert f"received {n + 1} dropped 0" in err, err
    rec = np.fromfile(tmp_path / "out.bin", dtype=[("t", "<i8"), ("head", "u1", (8,)), ("len", "<u4")])
    assert len(rec) == n and np.all(rec["len"] == 1206)
    assert np.array_equal(rec["head"], b[:, :8])                      # in order, intact
    assert np.array_equal(rec["t"], O.TimeSolver().hdl_many(gps, 1467331234567890))


def test_inssource_udp_to_transform_manager(tmp_path):
    from oracle import oracle as O
    rng = np.random.default_rng(12)
    n = 150
    recs = np.zeros(n, dtype=O.INS_DTYPE)
    recs["message_id"] = 508
    recs["week_number"] = 1903
    recs["week_number_pos"] = 1903
    recs["milliseconds"] = 345_600_000 + 10 * np.arange(n)
    # pose time - send time: distinct per record (equal timestamps overwrite each other in the
    # TimeLine, as in the reference), not monotonic
    recs["seconds_pos"] = 345_600.0 + 0.01 * np.arange(n) + (np.arange(n) % 7) * 1e-3 + np.arange(n) * 2e-6
    recs["LLH"] = np.array([39.8569901, 116.1736406, 89.09]) + np.cumsum(rng.normal(0, 1e-6, (n, 3)), axis=0)
    recs["V"] = rng.normal(0, 5, (n, 3))
    recs["Eulr"] = rng.uniform(-180, 180, (n, 3))
    feed = [np.frombuffer(recs[i].tobytes(), dtype=np.uint8) for i in range(n)]
    other = np.zeros(24, dtype=np.uint8)
    other[:2] = np.frombuffer(np.uint16(325).tobytes(), dtype=np.uint8)    # RAWINS: ignored
    feed.insert(40, other)
    port = F.free_udp_port()
    rc, err = F.run_with_udp_feed(["ins_host", port, n, tmp_path / "out.bin", tmp_path / "a.insmeta"],
                                  port, feed, pace_s=100e-6)
    assert rc == 0, err
    assert f"poses {n} timeline {n} meta {n}" in err, err
    got = np.fromfile(tmp_path / "out.bin", dtype=[("t", "<i8"), ("trv", "<f8", (9,))])
    want_t = O.ins_times(recs, 1467331200000000)
    order = np.argsort(want_t, kind="stable")
    assert np.array_equal(got["t"], want_t[order])
    want = O.ins_poses(recs, (-2781621.9891904, 4672106.75052387, 18.8910392))
    assert np.array_equal(got["trv"], want[order])                   # host libm on both sides


def test_frame_arena_adoption_is_zero_copy(tmp_path):
    """HDLFrame::points[row]->points adopts a slice of the frame's arena (FrameArena.h): same
    address, ordinary std::vector afterwards, arena released with its last vector."""
    r = F.run(["arena", tmp_path / "rep.txt"])
    assert r.returncode == 0, r.stderr
    rep = dict(line.split(" ", 1) for line in open(tmp_path / "rep.txt").read().splitlines())
    assert rep["alive_while_adopted"] == "1" and rep["zero_copy"] == "1"
    assert rep["size"] == "50" and rep["first"] == "10 last 59"
    assert rep["copy_is_heap"] == "1"
    assert rep["grown_size"] == "250 kept 59 new 1000"
    assert rep["resized_zero"] == "0"
    assert rep["moved"] == "1 1900" and rep["empty"] == "0" and rep["released"] == "1"
