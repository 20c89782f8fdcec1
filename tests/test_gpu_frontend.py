"""SURVEY 8f row N3 on the GPU: TimeSolver's hour-wrap scan over packet arrays and the INS
record -> ENU pose conversion, through the C ABI, against the CPU oracle (which is pinned
bit-for-bit to the reference's TimeSolver.cxx / CoordiTran.cpp in tests/test_oracle_vs_ref.py)."""
import numpy as np
import pytest

from oracle import oracle as O
from veloslam_b200 import capi, synth

import parity as P

pytestmark = pytest.mark.gpu

ORIG_XYZ = (-2781621.9891904, 4672106.75052387, 18.8910392)     # INSSource.cxx:334
NOW0 = 1_467_331_234_567_890


def _packets_with_gps(gps):
    pk, _ = synth.hdl64_packets(len(gps), seed=9)
    pk["gps"] = np.asarray(gps, dtype=np.uint32)
    return synth.as_bytes(pk)


def _gps_sequence(n, rng):
    """Sensor clock: 288 us per packet from just before an hour boundary, plus a few backward
    steps (each one is an hour for the reference) and duplicates."""
    g = (3_599_000_000 + 288 * np.arange(n, dtype=np.int64)) % 3_600_000_000
    for i in rng.integers(10, n - 10, 5):
        g[i] = g[i - 1] - 1          # backwards by 1 us
    for i in rng.integers(10, n - 10, 5):
        g[i] = g[i - 1]              # duplicate: not a wrap
    return g.astype(np.uint32)


@pytest.mark.parametrize("n", [1, 5, 1023, 1024, 1025, 20000])
def test_packet_times_match_time_solver(n):
    rng = np.random.default_rng(n)
    gps = _gps_sequence(n, rng) if n > 30 else (3_599_999_000 + 400 * np.arange(n)) % 3_600_000_000
    b = _packets_with_gps(gps)
    ctx = capi.Context(0, max_batch_packets=4096, max_poses=16)
    t, st = ctx.solve_packet_times(b, NOW0)
    want = O.TimeSolver()
    ref_t = want.hdl_many(gps, NOW0)
    assert np.array_equal(t, ref_t)
    assert (st.base_us, st.last_report, st.inited) == \
        (want.state.base_us, want.state.last_report, want.state.inited)
    ctx.close()


def test_packet_times_state_carries_across_calls_and_device_input():
    import torch
    rng = np.random.default_rng(77)
    gps = _gps_sequence(9000, rng)
    b = _packets_with_gps(gps)
    ctx = capi.Context(0, max_batch_packets=4096, max_poses=16)
    st = capi.TimeSolver()
    parts = []
    for a, e in ((0, 1), (1, 2500), (2500, 2501), (2501, 9000)):
        t, st = ctx.solve_packet_times(np.ascontiguousarray(b[a:e]), NOW0 + a, state=st)
        parts.append(t)
    want = O.TimeSolver().hdl_many(gps, NOW0)
    assert np.array_equal(np.concatenate(parts), want)
    # device packets in, device times out: what vs_submit(VS_FLAG_DEVICE_INPUT) consumes
    d_pk = torch.from_numpy(b).cuda()
    d_t = torch.empty(len(gps), dtype=torch.int64, device="cuda")
    _, st2 = ctx.solve_packet_times(d_pk, NOW0, n=len(gps), stride=1206, flags=capi.FLAG_DEVICE_INPUT,
                                    out=d_t)
    assert np.array_equal(d_t.cpu().numpy(), want) and st2.last_report == int(gps[-1])
    # ... and decode with them
    calib = synth.calib_hdl64()
    ctx.set_calibration(calib)
    r = ctx.wait(ctx.submit(d_pk[:4000], d_t[:4000], n=4000, stride=1206, flags=capi.FLAG_DEVICE_INPUT,
                            t_base_us=int(want[0])))
    o = P.make_oracle(calib)
    o.trace_enable()
    o.process_packets(b[:4000], want[:4000])
    P.assert_stream_parity(o, [(r, r.fetch(), 0)], P.TOL_DECODE, want[:4000], calib=calib)
    ctx.close()


def _ins_records(n, seed=3):
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=capi.INS_PVA_DTYPE)
    r["message_id"] = 508
    r["week_number"] = 1903
    r["milliseconds"] = 345_600_000 + 10 * np.arange(n)
    r["week_number_pos"] = 1903
    r["week_number_pos"][n // 2:] += 1
    r["seconds_pos"] = 345_600.0 + 0.01 * np.arange(n) + rng.uniform(0, 0.004, n)
    r["llh"][:, 0] = 39.8569901 + np.cumsum(rng.normal(0, 1e-6, n))
    r["llh"][:, 1] = 116.1736406 + np.cumsum(rng.normal(0, 1e-6, n))
    r["llh"][:, 2] = 89.09 + rng.normal(0, 0.05, n)
    r["v"] = rng.normal(0, 5, (n, 3))
    r["eulr"] = rng.uniform(-180, 180, (n, 3))
    return r


def test_ins_records_to_poses():
    recs = _ins_records(5000)
    arrival = 1_467_331_200_000_000 + 10_000 * np.arange(len(recs))
    ctx = capi.Context(0, max_batch_packets=1024, max_poses=8192)
    t, trv = ctx.poses_from_ins(recs, ORIG_XYZ, arrival)
    o_recs = recs.view(O.INS_DTYPE)
    assert np.array_equal(t, O.ins_times(o_recs, arrival))            # integer: bit-exact
    want = O.ins_poses(o_recs, ORIG_XYZ)
    assert np.array_equal(trv[:, 3:], want[:, 3:])                    # R, V copied
    # T: device sin/cos/tan vs host libm on ECEF coordinates of 6.4e6 m -> a few 1e-9 m
    assert np.max(np.abs(trv[:, :3] - want[:, :3])) < 1e-6
    # the result is a pose snapshot: feed it straight back
    order = np.argsort(t, kind="stable")
    keep = np.concatenate([[True], np.diff(t[order]) > 0])
    ctx.set_poses(t[order][keep], trv[order][keep])
    found, got, valid = ctx.interpolate(int(t[order][keep][10]) + 5)
    assert found and valid
    ctx.close()
