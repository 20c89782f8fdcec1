"""Parity at the sizes that are benchmarked (BASELINE.json configs[1]/[2] at 60 s, and the
bench's 1 Mi-packet batch): every point of the full stream against the oracle, the oracle run in
chunks to bound memory; and a sharded decode of a >= 1 Mi-packet stream (8 halo shards, as
configs[3] splits a recording) against the whole-stream decode and the oracle's frame list."""
import numpy as np
import pytest

from veloslam_b200 import capi, sharding, synth

import parity as P

pytestmark = pytest.mark.gpu

CHUNK = 16384


def _oracle_chunks(o, b, t):
    """Yield (trace dict, closed-frame summaries) per chunk of packets; trace and frames are
    dropped after each chunk (the oracle's packet / frame counters keep running)."""
    o.trace_enable()
    for a in range(0, b.shape[0], CHUNK):
        o.process_packets(b[a:a + CHUNK], t[a:a + CHUNK])
        tr = o.trace()
        frames = [o.frame_summary(i) for i in range(o.num_frames())]
        o.clear_frames()
        o.trace_enable(False)
        o.trace_enable(True)
        yield tr, frames


@pytest.mark.parametrize("with_poses", [False, True])
def test_every_point_of_a_full_minute_hdl64(with_poses):
    """configs[1] (decode + segmentation) and configs[2] (+ deskew against 100 Hz INS) at their
    stated size: 208 320 packets, 80.0 M return slots, ~76 M points, ~600 frames."""
    n = 208_320
    pk, t = synth.hdl64_packets(n)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(6100) if with_poses else None
    tol = P.TOL_DESKEW if with_poses else P.TOL_DECODE
    ctx = P.make_ctx(calib, poses, max_batch_packets=n)
    try:
        r = ctx.decode(b, t, t_base_us=int(t[0]))
        tab = r.frame_table
        first_points = tab["first_point"].astype(np.int64)
        o = P.make_oracle(calib, poses)
        off = 0
        k_frame = 0
        worst = 0.0
        exact = 0
        for tr, frames in _oracle_chunks(o, b, t):
            m = len(tr["x"])
            c = r.fetch(off, m)
            for k in ("laser", "intensity", "azimuth", "distance"):
                assert np.array_equal(c[k], tr[k]), (k, off)
            want_t = (t[tr["packet"]] - int(t[0])).astype(np.uint32) + tr["tadj_us"]
            assert np.array_equal(c["t_us"], want_t), off
            # frame id of every point from the GPU frame table == the oracle's
            fid = np.searchsorted(first_points, np.arange(off, off + m), side="right") - 1
            assert np.array_equal(fid, tr["frame"].astype(np.int64)), off
            for k in ("x", "y", "z"):
                d = np.abs(c[k].astype(np.float64) - tr[k].astype(np.float64))
                worst = max(worst, float(d.max()) if m else 0.0)
                exact += int(np.count_nonzero(c[k].view(np.uint32) == tr[k].view(np.uint32)))
            for ts, skips, npts, order, counts in frames:
                fr = tab[k_frame]
                assert fr["closed"] == 1 and int(fr["n_points"]) == npts, k_frame
                assert int(fr["timestamp_us"]) == ts and int(fr["skips"]) == skips, k_frame
                assert bool(fr["hdl64_order"]) == order
                lut = synth.HDL64_BEAM_LUT if order else np.arange(len(counts))
                assert np.array_equal(fr["laser_counts"][lut].astype(np.int64), counts), k_frame
                k_frame += 1
            off += m
        assert off == r.n_points and k_frame == r.n_closed and 598 <= k_frame <= 601
        assert worst <= tol, worst
        assert exact >= 0.999 * 3 * off          # bit-identical floats in > 99.9 % of the coordinates
    finally:
        ctx.close()


def test_eight_halo_shards_of_the_bench_batch_equal_the_whole_stream():
    """1 Mi packets (the bench's batch): rank g of 8 decodes [g N/8 - 512, (g+1) N/8) with a
    512-packet halo.  Concatenated shard columns == whole-stream columns, bit for bit; the
    stitched global frame index == the whole-stream frame table == the oracle's frame list."""
    n, world = 1 << 20, 8
    pk, t = synth.hdl64_stream_tiled(n)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(int(n * 288e-6 * 100) + 40)
    whole_ctx = P.make_ctx(calib, poses, max_batch_packets=n)
    shard_ctx = P.make_ctx(calib, poses, max_batch_packets=n // world + sharding.HALO_HDL64)
    try:
        whole = whole_ctx.decode(b, t, t_base_us=int(t[0]))
        wtab = whole.frame_table
        tables = []
        off = 0
        for g, (first, halo, end) in enumerate(sharding.shard_ranges(n, world, sharding.HALO_HDL64)):
            r = shard_ctx.decode(np.ascontiguousarray(b[first - halo:end]),
                                 np.ascontiguousarray(t[first - halo:end]), n_halo=halo,
                                 t_base_us=int(t[0]))
            tables.append(sharding.local_table(r.frame_table, g, first, halo))
            # columns in pieces: 48 M points per shard
            for a in range(0, r.n_points, 1 << 23):
                m = min(1 << 23, r.n_points - a)
                cs, cw = r.fetch(a, m), whole.fetch(off + a, m)
                for k in cs:
                    assert np.array_equal(cs[k].view(np.uint8), cw[k].view(np.uint8)), (g, k, a)
            off += r.n_points
        assert off == whole.n_points
        frames = sharding.stitch(tables)
        assert len(frames) == whole.n_frames
        assert not any("timestamp_mismatch" in f for f in frames)
        for i, f in enumerate(frames):
            w = wtab[i]
            assert f["n_points"] == int(w["n_points"]), i
            assert f["timestamp_us"] == int(w["timestamp_us"]) and f["skips"] == int(w["skips"]), i
            if i > 0:
                assert (f["start_packet"], f["start_block"]) == (int(w["start_packet"]), int(w["start_block"])), i
            assert f["closed"] == bool(w["closed"]) and f["hdl64_order"] == bool(w["hdl64_order"]), i
            # a global frame is one run of points: its segments are adjacent in the concatenation
            assert len(f["segments"]) <= 2
        # the oracle's frame list (frame by frame, chunked)
        o = P.make_oracle(calib, poses)
        k = 0
        for a in range(0, n, 32768):
            o.process_packets(b[a:a + 32768], t[a:a + 32768])
            for i in range(o.num_frames()):
                ts, skips, npts, order, _ = o.frame_summary(i)
                f = frames[k]
                assert f["closed"] and f["n_points"] == npts and f["timestamp_us"] == ts, k
                assert f["skips"] == skips and f["hdl64_order"] == order, k
                k += 1
            o.clear_frames()
        assert k == len(frames) - 1 and k > 3000          # + the open frame at the end
    finally:
        whole_ctx.close()
        shard_ctx.close()


def test_packet_times_outside_the_t_us_range_are_rejected():
    """u32(packet time - t_base) must not wrap silently (the 1-h config spans 84 % of the range)."""
    pk, t = synth.hdl64_packets(400)
    b = synth.as_bytes(pk)
    ctx = P.make_ctx(synth.calib_hdl64())
    try:
        with pytest.raises(capi.VeloError) as e:       # a packet earlier than t_base
            ctx.decode(b, t, t_base_us=int(t[10]))
        assert e.value.code == 1
        late = t.copy()
        late[300:] += 4_295_000_000                     # ~71.6 min later: past 2^32 - 65536 us
        with pytest.raises(capi.VeloError) as e:
            ctx.decode(b, late, t_base_us=int(t[0]))
        assert e.value.code == 1
        # the same through device input: caught by the kernels, reported by vs_wait
        import torch
        d_b = torch.from_numpy(b).cuda()
        d_t = torch.from_numpy(late).cuda()
        tk = ctx.submit(d_b, d_t, n=400, stride=1206, flags=capi.FLAG_DEVICE_INPUT, t_base_us=int(t[0]))
        with pytest.raises(capi.VeloError) as e:
            ctx.wait(tk)
        assert e.value.code == 1
        ok = late.copy()
        ok[300:] -= 100_000_000                         # back inside [0, 2^32 - 65536) us
        r = ctx.decode(b, ok, t_base_us=int(t[0]))
        c = r.fetch(columns=["t_us"])
        assert int(c["t_us"].max()) == int(ok[-1] - t[0])
        # a halo packet may lie before t_base: it emits nothing
        r = ctx.decode(np.ascontiguousarray(b), np.ascontiguousarray(t), n_halo=380, t_base_us=int(t[380]))
        assert r.n_points > 0
    finally:
        ctx.close()
