"""Rotation-sized contexts issue every batch as ONE CUDA graph per result slot (recorded once by
stream capture, node parameters refreshed per batch): same results as the direct launches, for
varying batch sizes, pinned host input, device input, pcap framing, filter / mode changes that
force a re-capture, and the fallback for pageable input."""
import numpy as np
import pytest
import torch

from veloslam_b200 import capi, pcapio, synth

import parity as P

pytestmark = pytest.mark.gpu


def _pinned(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def _stream(ctx, b, t, cuts, pinned=True, flags=0, mode=capi.MODE_STREAMING):
    carry = capi.carry_init()
    out = []
    hb, ht = (_pinned(b), _pinned(t)) if pinned else (b, t)
    for a, e in zip(cuts[:-1], cuts[1:]):
        r = ctx.wait(ctx.submit(hb[a:e], ht[a:e], n=e - a, stride=1206, mode=mode, flags=flags,
                                t_base_us=int(t[0]), carry=carry))
        out.append((r, r.fetch(), a))
        carry = r.carry_out
    return out


@pytest.mark.parametrize("pinned", [True, False])
def test_graph_batches_of_varying_size_match_the_oracle(pinned):
    """Rotation-sized batches of 300..400 packets through one small context: the first is
    captured, the rest refresh the node parameters (pinned input) -- or everything falls back to
    direct launches (pageable input)."""
    pk, t = synth.hdl64_packets(3000)
    b = synth.as_bytes(pk)
    calib, poses = synth.calib_hdl64(), synth.ins_trajectory(130)
    rng = np.random.default_rng(3)
    cuts = [0]
    while cuts[-1] < 3000:
        cuts.append(min(3000, cuts[-1] + int(rng.integers(300, 401))))
    o = P.make_oracle(calib, poses)
    o.trace_enable()
    o.process_packets(b, t)
    ctx = P.make_ctx(calib, poses, max_batch_packets=512)
    try:
        batches = _stream(ctx, b, t, cuts, pinned)
        P.assert_stream_parity(o, batches, P.TOL_DESKEW, t, calib=calib)
        # and as device-built HDLFrames on top of graph-issued batches
        o2 = P.make_oracle(calib, poses)
        o2.process_packets(b, t)
        P.assert_layout_parity(o2, P.gpu_layout_stream(ctx, b, t, cuts[1:-1]), P.TOL_DESKEW)
    finally:
        ctx.close()


def test_graph_is_recaptured_when_the_operation_sequence_changes():
    """Filters (crop kernel variant), per-point deskew, offline mode, index-only and pcap framing
    change the kernels / operations of a batch: each switch re-records, results stay right."""
    pk, t = synth.hdl64_packets(900)
    b = synth.as_bytes(pk)
    calib, poses = synth.calib_hdl64(), synth.ins_trajectory(50)
    ctx = P.make_ctx(calib, poses, max_batch_packets=1024)
    big = P.make_ctx(calib, poses, max_batch_packets=1 << 16)      # direct launches
    try:
        hb, ht = _pinned(b), _pinned(t)

        def both(**kw):
            ra = ctx.wait(ctx.submit(hb, ht, n=900, stride=1206, t_base_us=int(t[0]), **kw))
            rb = big.wait(big.submit(b, t, t_base_us=int(t[0]), **kw))
            ca, cb = ra.fetch(), rb.fetch()
            assert ra.n_points == rb.n_points and ra.n_frames == rb.n_frames
            for k in ca:
                assert np.array_equal(ca[k].view(np.uint8), cb[k].view(np.uint8)), (k, kw)
            assert np.array_equal(ra.frame_table, rb.frame_table)

        both()
        both()                                            # refreshed graph
        both(flags=capi.FLAG_DESKEW_PER_POINT)            # other decode kernel
        both(mode=capi.MODE_OFFLINE)
        for c in (ctx, big):
            c.set_filters(None, 1, True, 0, (-20, 20, -20, 20, -3, 3))   # crop variant of k_scan
        both()
        for c in (ctx, big):
            c.set_filters()
        both()
        sp_a = ctx.read_frame_information(b, t)            # index-only batches (pageable: direct)
        sp_b = big.read_frame_information(b, t)
        assert all(np.array_equal(x, y) for x, y in zip(sp_a, sp_b))
        both()
        # pcap framing: 1264-byte records, times from the record headers on the GPU
        img = pcapio.write_pcap_image(b, t)
        recs, nrec = pcapio.payload_view(img)
        himg = _pinned(img)
        base = himg.data_ptr() + (recs.ctypes.data - img.ctypes.data)
        ra = ctx.wait(ctx.submit(base, None, n=nrec, stride=pcapio.RECORD_BYTES, flags=capi.FLAG_PCAP_TIMES))
        rb = big.decode(recs, None, n=nrec, stride=pcapio.RECORD_BYTES, flags=capi.FLAG_PCAP_TIMES)
        ca, cb = ra.fetch(), rb.fetch()
        for k in ca:
            assert np.array_equal(ca[k].view(np.uint8), cb[k].view(np.uint8)), k
        both()
    finally:
        ctx.close()
        big.close()


def test_graph_with_device_input_and_two_slots():
    pk, t = synth.hdl64_packets(2000)
    b = synth.as_bytes(pk)
    calib, poses = synth.calib_hdl64(), synth.ins_trajectory(90)
    ctx = capi.Context(0, max_batch_packets=1024, max_poses=128, n_slots=2)
    ctx.set_calibration(calib)
    ctx.set_poses(*poses)
    ref = P.make_ctx(calib, poses)
    try:
        d_b = torch.from_numpy(b).cuda()
        d_t = torch.from_numpy(t).cuda()
        want = ref.decode(b[:1000], t[:1000], t_base_us=int(t[0]))
        wc = want.fetch()
        tickets = []
        for _ in range(6):                                # both slots, graphs re-used
            tickets.append(ctx.submit(d_b[:1000], d_t[:1000], n=1000, stride=1206,
                                      flags=capi.FLAG_DEVICE_INPUT, t_base_us=int(t[0])))
            if len(tickets) == 2:
                r = ctx.wait(tickets.pop(0))
                c = r.fetch()
                assert r.n_points == want.n_points
                for k in c:
                    assert np.array_equal(c[k].view(np.uint8), wc[k].view(np.uint8)), k
        ctx.wait(tickets.pop(0))
    finally:
        ctx.close()
        ref.close()
