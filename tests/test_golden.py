"""The CPU oracle against the committed golden fixtures (tests/golden/*.npz), which were
produced by the reference's own code (oracle/_ref, see tests/golden/make_golden.py).  Exact."""
import glob
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from veloslam_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if not p.endswith("interpolate.npz"))


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = int(g["calib_rows"])
    c = g["calib"]
    calib = synth.Calibration(c[0, :n], c[1, :n], c[2, :n], c[3, :n], c[4, :n], int(g["n_enabled"]))
    return g, calib


def configure(o, g, calib):
    o.set_calibration(calib)
    o.set_laser_selection(g["laser_selection"])
    o.set_points_skip(int(g["points_skip"]))
    if int(g["crop"][0]):
        o.set_crop(1, int(g["crop_inside"][0]), g["crop_region"])
    if len(g["pose_t"]):
        o.add_poses(g["pose_t"], g["pose_trv"])


def test_fixtures_exist():
    assert set(CASES) >= {"hdl64_deskew", "hdl32_nopose", "random_azimuth", "hdl64_filters"}


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_output(name):
    g, calib = load_case(name)
    o = Oracle()
    configure(o, g, calib)
    o.process_packets(g["packets"], g["t_us"])
    frames = o.frames()
    assert len(frames) == int(g["n_frames"])
    st = o.state()
    assert [st["last_azimuth"], st["firing_skip"], int(st["frame_meta_inited"]),
            int(st["is_hdl64"])] == list(g["state"])
    assert o.open_frame_points() == int(g["open_frame_points"])
    for i, f in enumerate(frames):
        assert np.array_equal(f.xyzi.view(np.uint32), g[f"f{i}_xyzi"].view(np.uint32)), i
        assert np.array_equal(f.azimuth, g[f"f{i}_azimuth"]), i
        assert np.array_equal(f.distance.view(np.uint32), g[f"f{i}_distance"].view(np.uint32)), i
        assert np.array_equal(f.laser_counts, g[f"f{i}_laser_counts"]), i
        ts, skips, npk, valid = g[f"f{i}_meta"]
        assert f.timestamp_us == ts and f.n_packets == npk and int(f.carpose_valid) == valid
        if f.skips >= 0:          # uninitialised uint8_t in the reference otherwise
            assert f.skips == skips
        assert np.array_equal(f.carpose_TRV.view(np.uint64), g[f"f{i}_carpose"].view(np.uint64))


def test_oracle_interpolation_table():
    g = np.load(os.path.join(GOLDEN, "interpolate.npz"))
    o = Oracle()
    o.add_poses(g["pose_t"], g["pose_trv"])
    for q, want, m in zip(g["query"], g["result"], g["matrix"]):
        ok, got, sp = o.interpolate(int(q))
        assert ok and sp == 0
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        assert np.array_equal(Oracle.pose_matrix(got).view(np.uint64), m.view(np.uint64))
