"""Parity harness: run the same packets through the CPU oracle and the CUDA C-ABI and compare.

Bars (BASELINE.json north_star): laser id, intensity, raw distance, azimuth, per-point time
offset, frame boundaries and point counts bit-exact; coordinates within 1e-4 m after decode
and 1e-3 m after deskew.
"""
import numpy as np

from oracle.oracle import Oracle
from veloslam_b200 import capi, synth
from veloslam_b200.frames import assemble_frame, point_meta_distance

TOL_DECODE = 1e-4   # metres, decode only
TOL_DESKEW = 1e-3   # metres, after the per-packet rigid transform


def make_oracle(calib, poses=None, laser_selection=None, points_skip=0, crop=None):
    o = Oracle()
    o.set_calibration(calib)
    if laser_selection is not None:
        o.set_laser_selection(laser_selection)
    o.set_points_skip(points_skip)
    if crop is not None:
        o.set_crop(1, crop[0], crop[1])
    if poses is not None:
        o.add_poses(poses[0], poses[1])
    return o


def make_ctx(calib, poses=None, laser_selection=None, points_skip=0, crop=None,
             max_batch_packets=1 << 16, n_slots=1):
    ctx = capi.Context(0, max_batch_packets=max_batch_packets,
                       max_poses=max(2, 0 if poses is None else len(poses[0])), n_slots=n_slots)
    ctx.set_calibration(calib)
    ctx.set_filters(laser_selection, points_skip, crop is not None,
                    crop[0] if crop is not None else 0,
                    crop[1] if crop is not None else (0,) * 6)
    if poses is not None:
        ctx.set_poses(poses[0], poses[1])
    return ctx


def gpu_stream(ctx, pk_bytes, t_us, splits=(), mode=capi.MODE_STREAMING, t_base=None):
    """Decode in batches cut at `splits`, chaining the carry.  Returns [(result, cols)]."""
    n = pk_bytes.shape[0]
    cuts = [0] + sorted({int(s) for s in splits if 0 < s < n}) + [n]
    carry = capi.carry_init()
    out = []
    t_base = int(t_us[0]) if t_base is None else t_base
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            continue
        r = ctx.decode(np.ascontiguousarray(pk_bytes[a:b]), np.ascontiguousarray(t_us[a:b]),
                       mode=mode, carry=carry, t_base_us=t_base)
        cols = r.fetch()
        out.append((r, cols, a))
        carry = r.carry_out
    return out


def concat_cols(batches):
    keys = [c for c, _ in capi.COLUMNS]
    return {k: np.concatenate([b[1][k] for b in batches]) if batches else np.zeros(0)
            for k in keys}


def point_frame_ids(batches):
    """Global frame id of every emitted point, from the per-batch frame tables."""
    ids = []
    base = 0
    for r, cols, _ in batches:
        f = np.zeros(r.n_points, dtype=np.int64)
        for i, fr in enumerate(r.frames):
            f[fr.first_point:fr.first_point + fr.n_points] = base + i
        ids.append(f)
        base += r.n_closed
    return np.concatenate(ids) if ids else np.zeros(0, np.int64)


def collect_frames(batches):
    """Merge per-batch frame entries into global frames."""
    frames = {}
    base = 0
    for r, cols, pkt0 in batches:
        for i, fr in enumerate(r.frames):
            g = base + i
            e = frames.setdefault(g, {"n_points": 0, "laser_counts": np.zeros(64, np.int64),
                                      "closed": False, "hdl64_order": False, "segments": []})
            e["n_points"] += fr.n_points
            e["laser_counts"] = e["laser_counts"] + fr.laser_counts
            e["segments"].append((r, cols, fr))
            # the latest batch that shows the frame has its meta (carried or initialised)
            e["timestamp_us"] = fr.timestamp_us
            e["skips"] = fr.skips
            e["carpose"] = fr.carpose
            e["carpose_valid"] = fr.carpose_valid
            if fr.closed:
                e["closed"] = True
                e["hdl64_order"] = fr.hdl64_order
        base += r.n_closed
    return frames


def frame_columns(entry):
    keys = [c for c, _ in capi.COLUMNS]
    return {k: np.concatenate([cols[k][fr.first_point:fr.first_point + fr.n_points]
                               for _, cols, fr in entry["segments"]]) for k in keys}


def xyz_stats(got, want):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    return {"max": float(d.max()) if d.size else 0.0,
            "exact": float(np.mean(got.view(np.uint32) == want.view(np.uint32))) if d.size else 1.0}


def assert_stream_parity(o, batches, tol, t_us, t_base=None, check_frames=True, calib=None):
    """o: oracle after processing the whole stream with trace enabled."""
    tr = o.trace()
    g = concat_cols(batches)
    assert len(g["x"]) == len(tr["x"]), (len(g["x"]), len(tr["x"]))
    for k in ("laser", "intensity", "azimuth", "distance"):
        assert np.array_equal(g[k], tr[k]), k
    t_base = int(t_us[0]) if t_base is None else t_base
    want_t = (np.asarray(t_us, np.int64)[tr["packet"]] - t_base).astype(np.uint32) + tr["tadj_us"]
    assert np.array_equal(g["t_us"], want_t), "t_us"
    assert np.array_equal(point_frame_ids(batches), tr["frame"].astype(np.int64)), "frame ids"
    stats = {}
    for k in ("x", "y", "z"):
        s = xyz_stats(g[k], tr[k])
        stats[k] = s
        assert s["max"] <= tol, (k, s)
    if check_frames:
        assert_frames_parity(o, batches, tol, calib)
    # carry-out state == oracle state
    st = o.state()
    co = batches[-1][0].carry_out
    assert co.last_azimuth == st["last_azimuth"]
    assert co.firing_skip == st["firing_skip"]
    assert bool(co.frame_meta_inited) == st["frame_meta_inited"]
    assert bool(co.is_hdl64) == st["is_hdl64"]
    return stats


def assert_frames_parity(o, batches, tol, calib=None):
    """Closed frames: counts, meta and laser-major contents (HDLFrame shape)."""
    gf = collect_frames(batches)
    of = o.frames()
    closed = [k for k in sorted(gf) if gf[k]["closed"]]
    assert len(closed) == len(of), (len(closed), len(of))
    for k, f in zip(closed, of):
        e = gf[k]
        assert e["n_points"] == f.n_points, (k, e["n_points"], f.n_points)
        assert e["timestamp_us"] == (f.timestamp_us if f.timestamp_us != -(2 ** 63)
                                     else capi.VS_TIME_NONE), (k, e["timestamp_us"], f.timestamp_us)
        assert e["skips"] == f.skips, (k, e["skips"], f.skips)
        assert e["carpose_valid"] == f.carpose_valid, k
        assert np.allclose(e["carpose"], f.carpose_TRV, rtol=0, atol=1e-9), k
        assert e["hdl64_order"] == f.is_hdl64_order, k
        seg0 = e["segments"][-1][2]
        fr_like = type("F", (), {"hdl64_order": e["hdl64_order"]})
        a = assemble_frame(frame_columns(e), fr_like, f.n_lasers)
        n_rows = 64 if e["hdl64_order"] else f.n_lasers
        assert np.array_equal(a.laser_counts[:n_rows], f.laser_counts[:n_rows]), k
        assert np.array_equal(a.xyzi[:, 3], f.xyzi[:, 3]), k
        assert np.array_equal(a.azimuth, f.azimuth), k
        d = np.abs(a.xyzi[:, :3].astype(np.float64) - f.xyzi[:, :3].astype(np.float64))
        assert d.size == 0 or d.max() <= tol, (k, d.max())
        if calib is not None and calib.n_enabled in (32, 64):
            rows = np.repeat(a.laser_rows[:n_rows], a.laser_counts[:n_rows])
            corr = np.zeros(64)
            corr[:calib.n_rows] = calib.dist_cm
            assert np.array_equal(point_meta_distance(a.distance_raw, rows, corr), f.distance), k
        del seg0


# ------------------------------------------------------------------------------------------
# HDLFrame layout built on the device (vs_layout_frames)
# ------------------------------------------------------------------------------------------
def gpu_layout_stream(ctx, pk_bytes, t_us, splits=(), mode=capi.MODE_STREAMING, with_meta=True):
    """Decode in batches cut at `splits` and take every frame in the device-built HDLFrame
    layout, the way the C++ facade does: the batch's first frame continues the open frame, its
    rows leave room for the carried points, which are copied into the gaps.  Returns the closed
    frames as dicts {xyzi (n, 4), meta, row_count (64,), order} -- NO host-side sort or scatter."""
    n = pk_bytes.shape[0]
    cuts = [0] + sorted({int(s) for s in splits if 0 < s < n}) + [n]
    carry = capi.carry_init()
    carried = np.zeros(64, np.uint32)
    partial = None
    out = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            continue
        r = ctx.decode(np.ascontiguousarray(pk_bytes[a:b]), np.ascontiguousarray(t_us[a:b]),
                       mode=mode, carry=carry, t_base_us=int(t_us[0]))
        xyzi, meta, rows = ctx.fetch_frames_layout(r.ticket, carried if carried.any() else None,
                                                   with_meta)
        assert len(rows) == r.n_frames
        for i, fr in enumerate(r.frames):
            rw = rows[i]
            s0, ns = int(rw["first_slot"]), int(rw["n_slots"])
            fx = xyzi[s0:s0 + ns].copy()
            fm = meta[s0:s0 + ns].copy() if with_meta else None
            if i == 0 and partial is not None:
                px, pm, pstart, pcount = partial
                for rr in range(64):
                    c = int(rw["row_carried"][rr])
                    if c == 0:
                        continue
                    laser = int(rw["row_laser"][rr])
                    assert c == pcount[laser]
                    d0 = int(rw["row_start"][rr])
                    fx[d0:d0 + c] = px[pstart[laser]:pstart[laser] + c]
                    if with_meta:
                        fm[d0:d0 + c] = pm[pstart[laser]:pstart[laser] + c]
            else:
                assert not rw["row_carried"].any()
            if fr.closed:
                out.append({"xyzi": fx, "meta": fm, "row_count": rw["row_count"].astype(np.int64),
                            "order": bool(fr.hdl64_order), "row_laser": rw["row_laser"].copy()})
                partial = None
            else:
                assert np.array_equal(rw["row_laser"], np.arange(64))   # open frames: by laser id
                partial = (fx, fm, rw["row_start"].astype(np.int64), rw["row_count"].astype(np.int64))
                carried = rw["row_count"].astype(np.uint32)
        carry = r.carry_out
    return out


def assert_layout_parity(o, frames, tol, with_meta=True):
    """Device-built frames == the oracle's HDLFrames, element for element in the reference's own
    order (rows of points[]/pointsMeta[], HDL64BeamLUT applied)."""
    of = o.frames()
    assert len(frames) == len(of), (len(frames), len(of))
    worst = 0.0
    for k, (g, f) in enumerate(zip(frames, of)):
        n_rows = len(f.laser_counts)
        assert g["order"] == f.is_hdl64_order, k
        assert np.array_equal(g["row_count"][:n_rows], f.laser_counts), k
        assert int(g["row_count"][n_rows:].sum()) == 0, k
        assert g["xyzi"].shape[0] == f.n_points, (k, g["xyzi"].shape, f.n_points)
        if f.n_points == 0:
            continue
        assert np.array_equal(g["xyzi"][:, 3], f.xyzi[:f.n_points, 3]), k
        d = np.abs(g["xyzi"][:, :3].astype(np.float64) - f.xyzi[:f.n_points, :3].astype(np.float64))
        worst = max(worst, float(d.max()))
        assert d.max() <= tol, (k, d.max())
        if with_meta:
            assert np.array_equal(g["meta"]["azimuth"], f.azimuth[:f.n_points]), k
            assert np.array_equal(g["meta"]["distance"], f.distance[:f.n_points]), k
            assert not g["meta"]["flags"].any() and not g["meta"]["intensityFlag"].any()
    return worst
