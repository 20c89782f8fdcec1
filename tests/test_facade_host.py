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
