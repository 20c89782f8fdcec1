"""vs_frame_table_rows_device (k_table_rows): the multi-GPU exchange rows built on the device
behind a batch's kernels must be the rows vs_frame_table_rows gives for vs_wait's frame list --
which the parity tests pin against the oracle -- in every situation the host code distinguishes:
fresh stream, carried open frame, halo shard (streaming / offline), index-only passes; and
VS_FLAG_NO_FRAME_LIST must change nothing but the list."""
import ctypes as C

import numpy as np
import pytest
import torch

from veloslam_b200 import capi, synth

import parity as P

pytestmark = pytest.mark.gpu


def device_rows(ctx, ticket, rank, first_packet, cap_rows=4096):
    d = torch.full((cap_rows + 1, capi.FRAME_ROW_COLS), -99, dtype=torch.int64, device="cuda")
    ctx.frame_table_rows_device(ticket, rank, first_packet, d, cap_rows)
    return d


def check(ctx, ticket, rank, first_packet, n_halo):
    d = device_rows(ctx, ticket, rank, first_packet)
    res = ctx.wait(ticket)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    want = capi.frame_table_rows(res.frame_table, rank, first_packet, n_halo)
    assert got[0, 0] == want.shape[0] == res.n_frames and not got[0, 1:].any()
    assert np.array_equal(got[1:1 + want.shape[0]], want)
    assert (got[1 + want.shape[0]:] == -99).all()          # nothing written past the table
    return res, want


@pytest.mark.parametrize("sensor", ["hdl64", "hdl32"])
def test_fresh_stream_and_carried_open_frame(sensor):
    if sensor == "hdl64":
        pk, t = synth.hdl64_packets(3000, seed=3)
        calib = synth.calib_hdl64()
    else:
        pk, t = synth.hdl32_packets(2500, az0=77.0, seed=4)
        calib = synth.calib_hdl32()
    b = synth.as_bytes(pk)
    ctx = P.make_ctx(calib, synth.ins_trajectory(120))
    carry = capi.carry_init()
    cuts = (0, 1100, 1101, 1900, len(t))       # a one-packet batch too
    seen = 0
    for a, e in zip(cuts[:-1], cuts[1:]):
        tk = ctx.submit(np.ascontiguousarray(b[a:e]), np.ascontiguousarray(t[a:e]), t_base_us=int(t[0]),
                        carry=carry)
        res, rows = check(ctx, tk, 5, a, 0)
        assert (rows[:, 9] == 5).all()
        carry = res.carry_out
        seen += res.n_closed
    assert seen > 5
    ctx.close()


@pytest.mark.parametrize("mode", [capi.MODE_STREAMING, capi.MODE_OFFLINE])
def test_halo_shard(mode):
    pk, t = synth.hdl64_packets(3000, seed=8)
    b = synth.as_bytes(pk)
    ctx = P.make_ctx(synth.calib_hdl64(), synth.ins_trajectory(120))
    cut, halo = 1700, 512
    tk = ctx.submit(np.ascontiguousarray(b[cut - halo:]), np.ascontiguousarray(t[cut - halo:]), n_halo=halo,
                    mode=mode, t_base_us=int(t[0]))
    res, rows = check(ctx, tk, 1, cut, halo)
    assert rows[0, 2] == -1 and rows[1, 2] >= cut      # global packet indices
    ctx.close()


def test_too_many_frames_for_the_buffer():
    pk, t = synth.hdl64_packets(2000)
    ctx = P.make_ctx(synth.calib_hdl64())
    tk = ctx.submit(synth.as_bytes(pk), t, t_base_us=int(t[0]))
    d = device_rows(ctx, tk, 0, 0, cap_rows=3)
    res = ctx.wait(tk)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    assert res.n_frames > 3 and got[0, 0] == -res.n_frames
    assert (got[1:] == -99).all()
    ctx.close()


def test_no_frame_list_flag_changes_nothing_but_the_list():
    pk, t = synth.hdl64_packets(4000, seed=21)
    b = synth.as_bytes(pk)
    poses = synth.ins_trajectory(160)
    ctx = P.make_ctx(synth.calib_hdl64(), poses)
    carry = capi.carry_init()
    for a, e in ((0, 2100), (2100, 4000)):
        sub = dict(t_base_us=int(t[0]), carry=carry)
        full = ctx.wait(ctx.submit(np.ascontiguousarray(b[a:e]), np.ascontiguousarray(t[a:e]), **sub))
        want = capi.frame_table_rows(full.frame_table, 0, a, 0)
        cols = full.fetch()
        tk = ctx.submit(np.ascontiguousarray(b[a:e]), np.ascontiguousarray(t[a:e]),
                        flags=capi.FLAG_NO_FRAME_LIST, **sub)
        d = device_rows(ctx, tk, 0, a)
        lean = ctx.wait(tk)
        torch.cuda.synchronize()
        assert lean.frames is None and len(lean.frame_table) == 0
        assert (lean.n_frames, lean.n_closed, lean.n_points) == (full.n_frames, full.n_closed, full.n_points)
        assert bytes(lean.carry_out) == bytes(full.carry_out)
        got = d.cpu().numpy()
        assert got[0, 0] == want.shape[0] and np.array_equal(got[1:1 + want.shape[0]], want)
        lc = lean.fetch()
        assert all(np.array_equal(cols[k], lc[k]) for k in cols)
        with pytest.raises(capi.VeloError):
            ctx.layout_frames(tk)
        carry = full.carry_out
    # halo shard with the flag: the first-frame patch and the carry still work
    cut, halo = 1700, 512
    sub = dict(n_halo=halo, t_base_us=int(t[0]))
    full = ctx.wait(ctx.submit(np.ascontiguousarray(b[cut - halo:]), np.ascontiguousarray(t[cut - halo:]), **sub))
    lean = ctx.wait(ctx.submit(np.ascontiguousarray(b[cut - halo:]), np.ascontiguousarray(t[cut - halo:]),
                               flags=capi.FLAG_NO_FRAME_LIST, **sub))
    assert bytes(lean.carry_out) == bytes(full.carry_out) and lean.n_frames == full.n_frames
    ctx.close()


def test_rows_feed_the_stitch_like_the_host_rows():
    """Two halo shards + the device rows -> vs_stitch_frame_tables == the whole stream's table."""
    pk, t = synth.hdl64_packets(5000, seed=2)
    b = synth.as_bytes(pk)
    ctx = P.make_ctx(synth.calib_hdl64(), synth.ins_trajectory(200))
    whole = ctx.wait(ctx.submit(b, t, t_base_us=int(t[0])))
    wt = whole.frame_table.copy()
    tabs = []
    for g in range(2):
        first, halo, end = capi.shard_range(len(t), 2, g, 512)
        tk = ctx.submit(np.ascontiguousarray(b[first - halo:end]), np.ascontiguousarray(t[first - halo:end]),
                        n_halo=halo, t_base_us=int(t[0]), flags=capi.FLAG_NO_FRAME_LIST)
        d = device_rows(ctx, tk, g, first)
        ctx.wait(tk)
        torch.cuda.synchronize()
        got = d.cpu().numpy()
        tabs.append(got[1:1 + got[0, 0]].copy())
    gf, segs = capi.stitch_frame_tables(tabs)
    assert gf["n_points"].tolist() == wt["n_points"].tolist()
    assert gf["timestamp_us"].tolist() == wt["timestamp_us"].tolist()
    assert gf["closed"].tolist() == wt["closed"].tolist()
    assert not gf["timestamp_mismatch"].any()
    ctx.close()
