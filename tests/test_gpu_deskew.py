"""Per-point deskew extension (VS_FLAG_DESKEW_PER_POINT) on the GPU.

The reference has no per-point motion compensation, so this mode cannot have reference parity.
It is checked (1) against the numpy statement of its semantics (oracle/deskew_port.py, "parity
unpinned") fed with the reference-pinned oracle's sensor-frame points, frame structure and
firing offsets, and (2) through properties that tie it to the reference-pinned per-packet path:
a stationary vehicle gives the sensor-frame points back; a pure translation with zero firing
offsets equals the per-packet mode; integer columns and frame tables never change.
Tolerance: 1e-3 m (north_star's bar after deskew); observed ~1e-5 m (float32 output).
"""
import numpy as np
import pytest

from oracle import deskew_port as D
from veloslam_b200 import capi, synth

import parity as P

pytestmark = pytest.mark.gpu


def _oracle_sensor_frame(b, t, calib):
    """Sensor-frame points, packet and firing offset of every emitted point, and the time of the
    origin packet for every packet, all from the reference-pinned oracle (no poses)."""
    o = P.make_oracle(calib)
    o.trace_enable()
    o.process_packets(b, t)
    tr = o.trace()
    o.split_frame()                      # close the open frame to read its timestamp
    stamps = sorted({int(t[0])} | {int(f.timestamp_us) for f in o.frames()
                                    if f.timestamp_us != -(2 ** 63)})
    stamps = np.array(stamps, dtype=np.int64)
    origin_time = stamps[np.searchsorted(stamps, t, side="right") - 1]
    xyz = np.stack([tr["x"], tr["y"], tr["z"]], axis=1).astype(np.float64)
    return tr, xyz, origin_time


def _decode(ctx, b, t, flags=0, splits=()):
    n = b.shape[0]
    cuts = [0] + sorted(s for s in splits if 0 < s < n) + [n]
    carry = capi.carry_init()
    cols = []
    for a, e in zip(cuts[:-1], cuts[1:]):
        r = ctx.wait(ctx.submit(np.ascontiguousarray(b[a:e]), np.ascontiguousarray(t[a:e]),
                                flags=flags, t_base_us=int(t[0]), carry=carry))
        cols.append(r.fetch())
        carry = r.carry_out
    return {k: np.concatenate([c[k] for c in cols]) for k in cols[0]}


@pytest.mark.parametrize("splits", [(), (700, 1301)])
def test_hdl32_per_point_deskew_matches_the_numpy_statement(splits):
    pk, t = synth.hdl32_packets(2600, az0=30000.0)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl32()
    poses = synth.ins_trajectory(200, yaw_amp_deg=80.0, yaw_period_s=6.0, speed=20.0)
    tr, xyz, origin_time = _oracle_sensor_frame(b, t, calib)
    want = D.deskew_points(xyz, tr["packet"], tr["tadj_us"], t, origin_time, *poses)
    ctx = P.make_ctx(calib, poses)
    got = _decode(ctx, b, t, capi.FLAG_DESKEW_PER_POINT, splits)
    ref = _decode(ctx, b, t, 0, splits)
    ctx.close()
    for k in ("laser", "intensity", "azimuth", "distance", "t_us"):     # untouched by the mode
        assert np.array_equal(got[k], ref[k]), k
    g = np.stack([got["x"], got["y"], got["z"]], axis=1).astype(np.float64)
    err = np.abs(g - want).max()
    assert err <= P.TOL_DESKEW, err
    assert err < 5e-5                       # float32 rounding of ~100 m coordinates
    # and it is a different answer from the reference's per-packet pose (up to ~2 cm here)
    r = np.stack([ref["x"], ref["y"], ref["z"]], axis=1).astype(np.float64)
    assert np.abs(g - r).max() > 1e-3


def test_hdl64_with_a_firing_table():
    pk, t = synth.hdl64_packets(1500)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(80, yaw_amp_deg=80.0, yaw_period_s=6.0, speed=20.0)
    # S2-like timing: the 6 block pairs of a packet fire 48 us apart, lasers 1.5 us apart
    off = np.zeros((12, 32), dtype=np.uint16)
    for j in range(12):
        off[j] = np.round((j // 2) * 48.0 + np.arange(32) * 1.5).astype(np.uint16)
    tr, xyz, origin_time = _oracle_sensor_frame(b, t, calib)
    point_off = off[tr["block"], tr["dsr"]]
    want = D.deskew_points(xyz, tr["packet"], point_off, t, origin_time, *poses)
    ctx = P.make_ctx(calib, poses)
    ctx.set_firing_offsets(off)
    got = _decode(ctx, b, t, capi.FLAG_DESKEW_PER_POINT)
    g = np.stack([got["x"], got["y"], got["z"]], axis=1).astype(np.float64)
    assert np.abs(g - want).max() <= P.TOL_DESKEW
    # the time column now carries the firing offsets
    want_t = (t[tr["packet"]] - t[0]).astype(np.uint32) + point_off
    assert np.array_equal(got["t_us"], want_t)
    # without the flag HDL-64 keeps the reference's "packet time for all 384 returns" (F5)
    plain = _decode(ctx, b, t, 0)
    assert np.array_equal(plain["t_us"], (t[tr["packet"]] - t[0]).astype(np.uint32))
    ctx.close()
    # HDL-32 has the reference's own table: the call is refused
    ctx = P.make_ctx(synth.calib_hdl32(), poses)
    with pytest.raises(capi.VeloError):
        ctx.set_firing_offsets(off)
    ctx.close()


def test_stationary_vehicle_returns_sensor_frame_points():
    pk, t = synth.hdl32_packets(1200)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl32()
    pt, trv = synth.ins_trajectory(120)
    trv[:] = trv[17]                      # every sample is the same pose
    _, xyz, _ = _oracle_sensor_frame(b, t, calib)
    ctx = P.make_ctx(calib, (pt, trv))
    got = _decode(ctx, b, t, capi.FLAG_DESKEW_PER_POINT)
    ctx.close()
    g = np.stack([got["x"], got["y"], got["z"]], axis=1).astype(np.float64)
    assert np.abs(g - xyz).max() < 2e-5


def test_pure_translation_without_offsets_equals_per_packet_mode():
    pk, t = synth.hdl64_packets(1500)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    pt, trv = synth.ins_trajectory(80)
    trv[:, 3:6] = 0.0                     # identity rotation: Euler lerp == slerp
    ctx = P.make_ctx(calib, (pt, trv))
    a = _decode(ctx, b, t, capi.FLAG_DESKEW_PER_POINT, splits=(400,))
    r = _decode(ctx, b, t, 0, splits=(400,))
    ctx.close()
    for k in ("x", "y", "z"):
        assert np.abs(a[k].astype(np.float64) - r[k]).max() < 2e-5, k


def test_flag_is_ignored_without_a_pose_bracket():
    pk, t = synth.hdl64_packets(300)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    ctx = P.make_ctx(calib, synth.ins_trajectory(1))
    a = _decode(ctx, b, t, capi.FLAG_DESKEW_PER_POINT)
    r = _decode(ctx, b, t, 0)
    ctx.close()
    for k in a:
        assert np.array_equal(a[k], r[k]), k


@pytest.mark.parametrize("sensor", ["hdl64", "hdl32"])
def test_per_point_deskew_against_an_independent_50_digit_evaluation(sensor):
    """The pin of this extension that does not share code, formulation or precision with the
    CUDA path: tests/mp_deskew.py evaluates the closed form with mpmath at 50 digits through the
    SO(3) geodesic (axis-angle / Rodrigues), no quaternions.  400 points sampled over a stream
    with a fast yaw (80 deg amplitude, 6 s period: up to 84 deg/s) -- which also bounds the error
    of the kernel's per-packet linearisation of the slerp weights.  Bar: 1e-3 m (north_star)."""
    import mp_deskew as M
    if sensor == "hdl64":
        pk, t = synth.hdl64_packets(1500)
        calib = synth.calib_hdl64()
        off = np.zeros((12, 32), dtype=np.uint16)
        for j in range(12):
            off[j] = np.round((j // 2) * 48.0 + np.arange(32) * 1.5).astype(np.uint16)
    else:
        pk, t = synth.hdl32_packets(2000, az0=30000.0)
        calib = synth.calib_hdl32()
        off = None
    b = synth.as_bytes(pk)
    poses = synth.ins_trajectory(200, yaw_amp_deg=80.0, yaw_period_s=6.0, speed=20.0)
    tr, xyz, origin_time = _oracle_sensor_frame(b, t, calib)
    point_off = off[tr["block"], tr["dsr"]] if off is not None else tr["tadj_us"]
    ctx = P.make_ctx(calib, poses)
    if off is not None:
        ctx.set_firing_offsets(off)
    got = _decode(ctx, b, t, capi.FLAG_DESKEW_PER_POINT, splits=(611,))
    ctx.close()
    g = np.stack([got["x"], got["y"], got["z"]], axis=1).astype(np.float64)
    tl = M.Timeline(*poses)
    rng = np.random.default_rng(2016)
    idx = np.sort(rng.choice(len(xyz), 400, replace=False))
    worst = 0.0
    for i in idx:
        P_ = int(tr["packet"][i])
        want = M.deskew_point(tl, xyz[i], int(t[P_]), int(point_off[i]), int(origin_time[P_]))
        worst = max(worst, float(np.abs(g[i] - np.array(want)).max()))
    assert worst <= P.TOL_DESKEW, worst
    assert worst < 1e-4, worst        # float32 outputs + float32 sensor-frame inputs of the check
