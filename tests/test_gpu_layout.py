"""The reference's HDLFrame layout built ON THE DEVICE (vs_layout_frames, k_layout):
points[row] / pointsMeta[row] lists, rows permuted by HDL64BeamLUT as splitFrame does
(/root/reference/HDLParser.cxx:733-751, 880-893), compared element for element with the
oracle's frames -- no argsort / scatter on the host side of the comparison."""
import numpy as np
import pytest

from veloslam_b200 import capi, synth

import parity as P

pytestmark = pytest.mark.gpu


def run(pk, t, calib, poses=None, splits=(), tol=P.TOL_DESKEW, **kw):
    b = synth.as_bytes(pk)
    o = P.make_oracle(calib, poses, **kw)
    o.process_packets(b, t)
    ctx = P.make_ctx(calib, poses, **kw)
    try:
        frames = P.gpu_layout_stream(ctx, b, t, splits)
        worst = P.assert_layout_parity(o, frames, tol)
    finally:
        ctx.close()
    return len(frames), worst


def test_hdl64_frames_in_one_batch():
    pk, t = synth.hdl64_packets(2500)
    n, worst = run(pk, t, synth.calib_hdl64(), synth.ins_trajectory(120))
    assert n >= 6 and worst <= 1e-4


@pytest.mark.parametrize("splits", [(347,), (100, 200, 300, 1000), (1, 2, 3, 700, 701), tuple(range(50, 2000, 50))])
def test_hdl64_batches_with_carried_rows(splits):
    """Frames that span batches: the first frame of a batch leaves room for the carried points
    (several batches in a row without a wrap included)."""
    pk, t = synth.hdl64_packets(2000)
    n, _ = run(pk, t, synth.calib_hdl64(), synth.ins_trajectory(100), splits=splits)
    assert n >= 5


def test_hdl32_and_no_poses():
    pk, t = synth.hdl32_packets(1500, az0=77.0)
    n, worst = run(pk, t, synth.calib_hdl32(), None, splits=(400, 900), tol=P.TOL_DECODE)
    assert n >= 7


def test_vlp16_rows_interleave_both_firings():
    """VLP-16: return slots l and l + 16 of a block are the same laser (HDLParser.cxx:935-943);
    within a row the first firing of a block comes before the second."""
    pk, t = synth.hdl32_packets(1200, az0=10.0)
    c = synth.calib_hdl32()
    c.n_enabled = 16
    n, _ = run(pk, t, c, None, splits=(333,), tol=P.TOL_DECODE)
    assert n >= 5


def test_random_azimuths_many_wraps_per_packet():
    """Several wraps inside one packet, frames without points, skipped blocks."""
    rng = np.random.default_rng(5)
    pk, t = synth.hdl64_packets(600)
    pk = pk.copy()
    pk["blocks"]["azimuth"] = rng.integers(0, 36000, size=pk["blocks"]["azimuth"].shape)
    n, _ = run(pk, t, synth.calib_hdl64(), synth.ins_trajectory(40), splits=(123, 124, 400))
    assert n > 500


def test_filters_and_crop():
    pk, t = synth.hdl64_packets(1200)
    sel = [1] * 64
    for i in (3, 17, 40, 63):
        sel[i] = 0
    n, _ = run(pk, t, synth.calib_hdl64(), synth.ins_trajectory(60), splits=(500,),
               laser_selection=sel, points_skip=1, crop=(0, (-20, 20, -20, 20, -3, 3)))
    assert n >= 3


def test_xyzi_stride_32_is_pcl_pointxyzi_and_meta_is_optional():
    pk, t = synth.hdl64_packets(800)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    ctx = P.make_ctx(calib, synth.ins_trajectory(40))
    try:
        r = ctx.decode(b, t, t_base_us=int(t[0]))
        x16, meta, rows = ctx.fetch_frames_layout(r.ticket)
        lay, rows32 = ctx.layout_frames(r.ticket, None, 32, False)
        assert lay.meta is None and lay.xyzi_stride == 32 and lay.n_slots == r.n_points
        raw = np.zeros((r.n_points, 8), np.float32)
        ctx.fetch_layout(r.ticket, 0, r.n_points, raw, None)
        ctx.sync(r.ticket)
        assert np.array_equal(raw[:, :3], x16[:, :3])
        assert np.all(raw[:, 3] == 1.0)                      # PCL's data[3]
        assert np.array_equal(raw[:, 4], x16[:, 3])          # intensity behind the 16-byte xyz1
        assert not raw[:, 5:].any()
        assert np.array_equal(rows["row_start"], rows32["row_start"])
        with pytest.raises(capi.VeloError):
            ctx.fetch_layout(r.ticket, 0, 1, None, raw)      # no PointMeta in this layout
    finally:
        ctx.close()


def test_halo_shard_layout_matches_whole_stream():
    """A shard that rebuilds its state from a halo lays out the same frames as the whole stream."""
    pk, t = synth.hdl64_packets(3000)
    b = synth.as_bytes(pk)
    calib, poses = synth.calib_hdl64(), synth.ins_trajectory(140)
    o = P.make_oracle(calib, poses)
    o.process_packets(b, t)
    of = o.frames()
    ctx = P.make_ctx(calib, poses)
    try:
        first, halo = 1500, 512
        r = ctx.decode(np.ascontiguousarray(b[first - halo:]), np.ascontiguousarray(t[first - halo:]),
                       n_halo=halo, t_base_us=int(t[0]))
        xyzi, meta, rows = ctx.fetch_frames_layout(r.ticket)
        # frames of the shard from its second entry on are whole frames of the stream
        k0 = len(of) - (r.n_frames - 1)   # index of the shard's first (partial) frame in the stream
        for i in range(1, r.n_frames - 1):
            f = of[k0 + i]
            rw = rows[i]
            s0, ns = int(rw["first_slot"]), int(rw["n_slots"])
            assert ns == f.n_points
            assert np.array_equal(rw["row_count"].astype(np.int64), f.laser_counts)
            assert np.array_equal(meta["azimuth"][s0:s0 + ns], f.azimuth[:ns])
            d = np.abs(xyzi[s0:s0 + ns, :3].astype(np.float64) - f.xyzi[:ns, :3])
            assert d.max() <= P.TOL_DESKEW
    finally:
        ctx.close()


def test_layout_needs_a_finished_decode_batch():
    pk, t = synth.hdl64_packets(100)
    ctx = P.make_ctx(synth.calib_hdl64())
    try:
        with pytest.raises(capi.VeloError):
            ctx.layout_frames(12345)
        tk = ctx.submit(synth.as_bytes(pk), t, t_base_us=int(t[0]))
        with pytest.raises(capi.VeloError):
            ctx.layout_frames(tk)             # not waited yet
        r = ctx.wait(tk)
        with pytest.raises(capi.VeloError):
            ctx.layout_frames(r.ticket, None, 24, True)   # stride must be 16 or 32
        ctx.fetch_frames_layout(r.ticket)
    finally:
        ctx.close()


def test_offline_mode_layout_matches_get_frame():
    """VS_MODE_OFFLINE (readFrameInformation + getFrame): frame i = blocks [start_i, start_i+1),
    no dropped blocks; laid out on the device == oracle.get_frame of every index entry."""
    from oracle.oracle import Oracle
    pk, t = synth.hdl64_packets(1300)
    b = synth.as_bytes(pk)
    calib, poses = synth.calib_hdl64(), synth.ins_trajectory(60)
    sp, sk, ts = Oracle.read_frame_information(b, t)
    o = P.make_oracle(calib, poses)
    ctx = P.make_ctx(calib, poses)
    try:
        frames = P.gpu_layout_stream(ctx, b, t, splits=(450, 451), mode=capi.MODE_OFFLINE)
        assert len(frames) == len(sp) - 1          # the last index entry stays open
        for i, g in enumerate(frames):
            want = o.get_frame(b, t, sp[i], sk[i], first=True)
            n_rows = len(want.laser_counts)
            assert np.array_equal(g["row_count"][:n_rows], want.laser_counts), i
            assert np.array_equal(g["meta"]["azimuth"], want.azimuth[:want.n_points]), i
            assert np.array_equal(g["meta"]["distance"], want.distance[:want.n_points]), i
            d = np.abs(g["xyzi"][:, :3].astype(np.float64) - want.xyzi[:want.n_points, :3])
            assert d.max() <= P.TOL_DESKEW, i
    finally:
        ctx.close()


def test_layout_of_an_empty_and_a_wrap_free_batch():
    """No points at all (every distance 0) and a batch without a single wrap (one open frame)."""
    pk, t = synth.hdl64_packets(100)
    pk = pk.copy()
    pk["blocks"]["returns"]["distance"] = 0
    ctx = P.make_ctx(synth.calib_hdl64())
    try:
        r = ctx.decode(synth.as_bytes(pk), t, t_base_us=int(t[0]))
        assert r.n_points == 0
        xyzi, meta, rows = ctx.fetch_frames_layout(r.ticket)
        assert xyzi.shape[0] == 0 and len(rows) == r.n_frames and not rows["n_slots"].any()
        pk2, t2 = synth.hdl64_packets(120, az0=100.0)          # 120 packets: a third of a rotation
        r = ctx.decode(synth.as_bytes(pk2), t2, t_base_us=int(t2[0]))
        assert r.n_frames == 1 and r.n_closed == 0
        xyzi, meta, rows = ctx.fetch_frames_layout(r.ticket)
        assert int(rows["n_slots"][0]) == r.n_points == xyzi.shape[0]
        cols = r.fetch()
        # rows by laser id, time order inside a row: a stable sort of the stream order
        order = np.argsort(cols["laser"], kind="stable")
        assert np.array_equal(xyzi[:, 0], cols["x"][order])
        assert np.array_equal(meta["azimuth"], cols["azimuth"][order])
        assert np.array_equal(rows["row_count"][0], np.bincount(cols["laser"], minlength=64))
    finally:
        ctx.close()
