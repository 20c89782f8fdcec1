"""The C++ facade (HDLParser / HDLFrame / TransformManager with the reference's names) driven
like the reference's consumers, on the GPU, against the oracle: laser-major frame contents."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from veloslam_b200 import calibxml, pcapio, synth

import facade_util as F
import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _driver():
    F.build()


def compare(frames, oracle_frames, tol):
    assert len(frames) == len(oracle_frames)
    for a, b in zip(frames, oracle_frames):
        assert a.n_points == b.n_points
        assert a.timestamp_us == b.timestamp_us and a.skips == b.skips
        assert a.n_packets == b.n_packets
        assert a.carpose_valid == b.carpose_valid
        assert np.allclose(a.carpose_TRV, b.carpose_TRV, rtol=0, atol=1e-9)
        assert np.array_equal(a.laser_counts, b.laser_counts)
        assert np.array_equal(a.azimuth, b.azimuth)
        assert np.array_equal(a.distance, b.distance)
        assert np.array_equal(a.xyzi[:, 3], b.xyzi[:, 3])
        d = np.abs(a.xyzi[:, :3].astype(np.float64) - b.xyzi[:, :3])
        assert d.size == 0 or d.max() <= tol


@pytest.mark.parametrize("sensor,batch,pipelined", [("hdl64", 4096, 0), ("hdl64", 100, 0), ("hdl32", 4096, 0),
                                                   ("hdl64", 400, 1), ("hdl64", 64, 1), ("hdl32", 500, 1)])
def test_streaming_like_hdlsource(tmp_path, sensor, batch, pipelined):
    if sensor == "hdl64":
        pk, t = synth.hdl64_packets(1500)
        calib = synth.calib_hdl64()
    else:
        pk, t = synth.hdl32_packets(1200, az0=123.0)
        calib = synth.calib_hdl32()
    poses = synth.ins_trajectory(80)
    b = synth.as_bytes(pk)
    b.tofile(tmp_path / "pk.bin")
    t.astype("<i8").tofile(tmp_path / "t.bin")
    F.write_poses(tmp_path / "poses.bin", *poses)
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    r = F.run(["stream", tmp_path / "db.xml", tmp_path / "pk.bin", tmp_path / "t.bin",
               tmp_path / "poses.bin", tmp_path / "out.bin", batch, pipelined])
    assert r.returncode == 0, r.stderr
    o = P.make_oracle(calib, poses)
    o.process_packets(b, t)
    compare(F.read_frames(tmp_path / "out.bin"), o.frames(), P.TOL_DESKEW)


def test_offline_like_hdlmanager(tmp_path):
    pk, t = synth.hdl64_packets(1300)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(60)
    b = synth.as_bytes(pk)
    path = tmp_path / "20160701T000000.pcap"
    pcapio.write_pcap(str(path), b, t)
    F.write_poses(tmp_path / "poses.bin", *poses)
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    r = F.run(["offline", tmp_path / "db.xml", path, tmp_path / "poses.bin", tmp_path / "out.bin"])
    assert r.returncode == 0, r.stderr
    frames = F.read_frames(tmp_path / "out.bin")
    o = P.make_oracle(calib, poses)
    sp, sk, ts = Oracle.read_frame_information(b, t)
    assert len(frames) == len(sp)
    for i, f in enumerate(frames):
        want = o.get_frame(b, t, sp[i], sk[i])
        assert f.n_points == want.n_points and f.timestamp_us == ts[i] and f.skips == sk[i]
        assert np.array_equal(f.laser_counts, want.laser_counts)
        assert np.array_equal(f.azimuth, want.azimuth)
        d = np.abs(f.xyzi[:, :3].astype(np.float64) - want.xyzi[:, :3])
        assert d.max() <= P.TOL_DESKEW


# --- HDLManager: loadOffline + getFrameAt, recording resident in HBM vs file ------
@pytest.mark.parametrize("mode", ["manager", "manager_file"])
def test_hdlmanager_load_offline_and_get_frame(tmp_path, mode):
    pk, t = synth.hdl64_packets(1300, seed=64)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(60)
    b = synth.as_bytes(pk)
    path = tmp_path / "20160701T000000.pcap"
    pcapio.write_pcap(str(path), b, t)
    F.write_poses(tmp_path / "poses.bin", *poses)
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    r = F.run([mode, tmp_path / "db.xml", path, tmp_path / "poses.bin", tmp_path / "out.bin"])
    assert r.returncode == 0, r.stderr
    assert "resident 1" in r.stderr        # fixed-stride packet file: uploaded to HBM once
    frames = F.read_frames(tmp_path / "out.bin")
    o = P.make_oracle(calib, poses)
    sp, sk, ts = Oracle.read_frame_information(b, t)
    assert len(frames) == len(sp)          # GPU index == reference index
    for i, f in enumerate(frames):
        want = o.get_frame(b, t, sp[i], sk[i])
        assert f.n_points == want.n_points and f.timestamp_us == ts[i] and f.skips == sk[i]
        ok, trv, spos = o.interpolate(int(ts[i]))     # HDLManager::loadOffline sets carpose
        assert np.allclose(f.carpose_TRV, trv, rtol=0, atol=1e-12)
        assert np.array_equal(f.laser_counts, want.laser_counts)
        assert np.array_equal(f.azimuth, want.azimuth)
        assert np.array_equal(f.xyzi[:, 3], want.xyzi[:, 3])
        d = np.abs(f.xyzi[:, :3].astype(np.float64) - want.xyzi[:, :3])
        assert d.max() <= P.TOL_DESKEW


def test_hdlmanager_renames_recording_and_index_spans_batches(tmp_path):
    """A file not named after its first packet is renamed (HDLParser.cxx:1150-1158) and keeps
    being served from HBM; 9000 packets > the parser's 4096-packet batch: the GPU index runs in
    chunks linked by the carry."""
    pk, t = synth.hdl32_packets(9000, seed=5)
    calib = synth.calib_hdl32()
    b = synth.as_bytes(pk)
    path = tmp_path / "drive.pcap"
    pcapio.write_pcap(str(path), b, t)
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    r = F.run(["manager", tmp_path / "db.xml", path, "-", tmp_path / "out.bin"])
    assert r.returncode == 0, r.stderr
    assert not path.exists() and len(list(tmp_path.glob("2016*T*.pcap"))) == 1
    frames = F.read_frames(tmp_path / "out.bin")
    sp, sk, ts = Oracle.read_frame_information(b, t)
    assert len(frames) == len(sp) and len(sp) > 40
    o = P.make_oracle(calib)
    for i in (0, 1, len(sp) // 2, len(sp) - 1):
        want = o.get_frame(b, t, sp[i], sk[i])
        f = frames[i]
        assert f.timestamp_us == ts[i] and f.skips == sk[i] and f.n_points == want.n_points
        assert np.array_equal(f.azimuth, want.azimuth)
        assert np.array_equal(f.xyzi.view(np.uint32), want.xyzi.view(np.uint32))


@pytest.mark.parametrize("sensor,n,devices", [("hdl64", 5000, "0,0,0"), ("hdl32", 3000, "0,0"),
                                              ("hdl64", 3, "0,0,0,0,0")])
def test_hdlmanager_recording_spread_over_devices(tmp_path, sensor, n, devices):
    """HDLManager::setDevices: one range of the recording per CUDA context (here several on one
    GPU), indices concatenated, each frame decoded where it lives -- same frames as the
    reference's one-file index + getFrame, including the rotations that cross a range end."""
    import torch
    if devices == "0,0,0" and torch.cuda.device_count() >= 2:
        devices = ",".join(str(d) for d in range(min(torch.cuda.device_count(), 8)))
    if sensor == "hdl64":
        pk, t = synth.hdl64_packets(n, seed=11)
        calib = synth.calib_hdl64()
    else:
        pk, t = synth.hdl32_packets(n, seed=12)
        calib = synth.calib_hdl32()
    poses = synth.ins_trajectory(max(60, n // 20))
    b = synth.as_bytes(pk)
    path = tmp_path / "drive.pcap"                  # renamed by the touch before the split
    pcapio.write_pcap(str(path), b, t)
    F.write_poses(tmp_path / "poses.bin", *poses)
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    r = F.run(["manager_shards", tmp_path / "db.xml", path, tmp_path / "poses.bin", tmp_path / "out.bin", devices])
    assert r.returncode == 0, r.stderr
    assert f"shards {len(devices.split(','))}" in r.stderr
    frames = F.read_frames(tmp_path / "out.bin")
    o = P.make_oracle(calib, poses)
    sp, sk, ts = Oracle.read_frame_information(b, t)
    assert [f.timestamp_us for f in frames] == [int(x) for x in ts]
    assert [f.skips for f in frames] == [int(x) for x in sk]
    for i, f in enumerate(frames):
        want = o.get_frame(b, t, sp[i], sk[i])
        assert f.n_points == want.n_points
        assert np.array_equal(f.laser_counts, want.laser_counts)
        assert np.array_equal(f.azimuth, want.azimuth)
        assert np.array_equal(f.xyzi[:, 3], want.xyzi[:, 3])
        d = np.abs(f.xyzi[:, :3].astype(np.float64) - want.xyzi[:, :3])
        assert d.size == 0 or d.max() <= P.TOL_DESKEW


# --- the whole online path over loopback UDP -------------------------------------------
def test_online_udp_to_hdlmanager(tmp_path):
    n = 1500
    pk, t = synth.hdl64_packets(n, seed=33)
    # the sensor clock: microseconds past the hour, consistent with the packet times
    pk["gps"] = ((t - synth.T0_US) % 3_600_000_000).astype(np.uint32)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(80)
    b = synth.as_bytes(pk)
    F.write_poses(tmp_path / "poses.bin", *poses)
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    port = F.free_udp_port()
    rc, err = F.run_with_udp_feed(["udp", tmp_path / "db.xml", port, tmp_path / "poses.bin",
                                   tmp_path / "out.bin", n], port, [row for row in b], pace_s=60e-6)
    assert rc == 0, err
    assert f"received {n} dropped 0 consumed {n}" in err, err
    # TimeSolver: first packet = the (injected) clock, later ones by the sensor clock: == t
    o = P.make_oracle(calib, poses)
    o.process_packets(b, t)
    compare(F.read_frames(tmp_path / "out.bin"), o.frames(), P.TOL_DESKEW)


@pytest.mark.parametrize("store,meta,pipelined,batch", [(1, 1, 1, 700), (0, 0, 1, 512), (0, 1, 0, 4096)])
def test_consumer_loop_counts_match_the_oracle(tmp_path, store, meta, pipelined, batch):
    """facade_driver bench (what bench.py's e2e_facade runs): processHDLPacket per packet +
    getAllFrames/clearAllFrames over the packet array twice (warm-up pass + 1 timed pass, one
    continuous stream).  Frames and points handed out in the timed pass == the oracle's."""
    import json
    n = 2500
    pk, t = synth.hdl64_packets(n)
    calib = synth.calib_hdl64()
    b = synth.as_bytes(pk)
    b.tofile(tmp_path / "pk.bin")
    t.astype("<i8").tofile(tmp_path / "t.bin")
    calibxml.write_db_xml(str(tmp_path / "db.xml"), calib)
    r = F.run(["bench", tmp_path / "db.xml", tmp_path / "pk.bin", tmp_path / "t.bin", "-", batch, 1,
               store, meta, pipelined, 0])
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    span = int(t[-1] - t[0]) + int(t[-1] - t[0]) // (n - 1)
    o = P.make_oracle(calib)
    o.process_packets(b, t)
    warm = o.num_frames()                    # closed during the warm-up pass: handed out untimed
    o.process_packets(b, t + span)
    frames = o.frames()
    assert got["frames"] == len(frames) - warm
    assert got["points"] == sum(f.n_points for f in frames[warm:])
