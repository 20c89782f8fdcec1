"""The CUDA path (through the C ABI) against the committed golden fixtures that the reference's
own code produced (tests/golden/make_golden.py): integers exact, coordinates within the
north_star tolerances (1e-4 m decode, 1e-3 m after deskew)."""
import numpy as np
import pytest

from veloslam_b200 import capi
from veloslam_b200.frames import assemble_frame, point_meta_distance

import parity as P
from test_golden import CASES, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_reference_output(name):
    g, calib = load_case(name)
    poses = (g["pose_t"], g["pose_trv"]) if len(g["pose_t"]) else None
    crop = (int(g["crop_inside"][0]), tuple(g["crop_region"])) if int(g["crop"][0]) else None
    ctx = P.make_ctx(calib, poses, laser_selection=g["laser_selection"],
                     points_skip=int(g["points_skip"]), crop=crop)
    t = g["t_us"]
    r = ctx.decode(np.ascontiguousarray(g["packets"]), np.ascontiguousarray(t), t_base_us=int(t[0]))
    cols = r.fetch()
    assert r.n_closed == int(g["n_frames"])
    co = r.carry_out
    assert [co.last_azimuth, co.firing_skip, co.frame_meta_inited, co.is_hdl64] == list(g["state"])
    assert r.frames[-1].n_points == int(g["open_frame_points"])
    tol = P.TOL_DESKEW if poses is not None else P.TOL_DECODE
    exact = []
    for i in range(r.n_closed):
        f = r.frames[i]
        c = {k: v[f.first_point:f.first_point + f.n_points] for k, v in cols.items()}
        want = g[f"f{i}_xyzi"]
        n_lasers = len(g[f"f{i}_laser_counts"])
        a = assemble_frame(c, f, n_lasers)
        n_rows = 64 if f.hdl64_order else n_lasers
        assert np.array_equal(a.laser_counts[:n_rows], g[f"f{i}_laser_counts"][:n_rows]), i
        assert np.array_equal(a.azimuth, g[f"f{i}_azimuth"]), i
        assert np.array_equal(a.xyzi[:, 3], want[:, 3]), i
        rows = np.repeat(a.laser_rows[:n_rows], a.laser_counts[:n_rows])
        corr = np.zeros(64)
        corr[:calib.n_rows] = calib.dist_cm
        assert np.array_equal(point_meta_distance(a.distance_raw, rows, corr), g[f"f{i}_distance"]), i
        d = np.abs(a.xyzi[:, :3].astype(np.float64) - want[:, :3])
        assert d.size == 0 or d.max() <= tol, (i, d.max())
        exact.append(np.mean(a.xyzi[:, :3].view(np.uint32) == want[:, :3].view(np.uint32))
                     if d.size else 1.0)
        ts, skips, _, valid = g[f"f{i}_meta"]
        assert f.timestamp_us == ts and int(f.carpose_valid) == valid
        if f.skips >= 0:
            assert f.skips == skips
        assert np.allclose(f.carpose, g[f"f{i}_carpose"], rtol=0, atol=1e-9)
    assert min(exact) > 0.999
    ctx.close()


def test_cuda_host_interpolation_matches_reference_table():
    import os
    from test_golden import GOLDEN
    g = np.load(os.path.join(GOLDEN, "interpolate.npz"))
    ctx = capi.Context(0, 1024, 128, 1)
    ctx.set_poses(g["pose_t"], g["pose_trv"])
    for q, want in zip(g["query"], g["result"]):
        found, got, valid = ctx.interpolate(int(q))
        assert found and valid
        assert np.allclose(got, want, rtol=0, atol=1e-10)
    ctx.close()
