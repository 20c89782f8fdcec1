"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
identical seeded inputs.  Run on the B200 box with `pytest -m gpu`.
"""
import numpy as np
import pytest

from veloslam_b200 import capi, synth

import parity as P
from helpers import make_packet

pytestmark = pytest.mark.gpu


def _run_streaming(pk, t, calib, poses=None, splits=(), tol=None, **filt):
    b = synth.as_bytes(pk)
    o = P.make_oracle(calib, poses, **filt)
    o.trace_enable()
    o.process_packets(b, t)
    ctx = P.make_ctx(calib, poses, **filt)
    batches = P.gpu_stream(ctx, b, t, splits)
    tol = tol if tol is not None else (P.TOL_DESKEW if poses is not None and len(poses[0]) >= 2
                                       else P.TOL_DECODE)
    stats = P.assert_stream_parity(o, batches, tol, t, calib=calib)
    ctx.close()
    return stats


# --- config 2: HDL-64E S2 decode + frame segmentation ---------------------------------------
def test_hdl64_decode_and_segmentation():
    pk, t = synth.hdl64_packets(4000)
    stats = _run_streaming(pk, t, synth.calib_hdl64())
    # the reference evaluates sin/cos(rad(az/100 - rotCorrection)) with libm; the kernel uses
    # the angle-difference identity on its LUT: doubles agree to ~1e-16, floats almost always
    assert min(s["exact"] for s in stats.values()) > 0.999
    assert max(s["max"] for s in stats.values()) <= 2e-5


def test_hdl64_identity_calibration_is_bit_exact():
    """rotCorrection == 0 is the reference's LUT branch: doubles are identical, so are floats."""
    pk, t = synth.hdl64_packets(1500)
    stats = _run_streaming(pk, t, synth.calib_identity(64), tol=0.0)
    assert all(s["exact"] == 1.0 for s in stats.values())


# --- config 1: HDL-32E, no poses ----------------------------------------------------------------
def test_hdl32_stream_is_bit_exact():
    pk, t = synth.hdl32_packets(3000)
    stats = _run_streaming(pk, t, synth.calib_hdl32(), tol=0.0)
    assert all(s["exact"] == 1.0 for s in stats.values())


def test_hdl32_azimuth_adjust_ties():
    """azimuthDiff = 20 makes azimuthDiff * ratio land on x.5 for every odd dsr multiple:
    std::round (half away from zero) must be reproduced exactly."""
    n = 40
    az = (np.arange(12 * n) * 20 + 35000) % 36000
    rng = np.random.default_rng(5)
    pks = []
    for i in range(n):
        d = rng.integers(1, 40000, (12, 32)).astype(np.uint16)
        pks.append(make_packet(az[12 * i:12 * i + 12], None, d,
                               rng.integers(0, 256, (12, 32)).astype(np.uint8)))
    pk = np.concatenate(pks)
    t = synth.T0_US + 553 * np.arange(n, dtype=np.int64)
    _run_streaming(pk, t, synth.calib_hdl32(), tol=0.0)


def test_vlp16_mode():
    pk, t = synth.hdl32_packets(400, seed=16)
    c = synth.calib_hdl32()
    c.n_enabled = 16
    _run_streaming(pk, t, c, tol=0.0)


# --- config 3: decode + deskew -------------------------------------------------------------------
def test_hdl64_deskew_against_ins_timeline():
    pk, t = synth.hdl64_packets(4000)
    poses = synth.ins_trajectory(140)          # 1.4 s of 100 Hz poses around the 1.15 s stream
    stats = _run_streaming(pk, t, synth.calib_hdl64(), poses)
    assert max(s["max"] for s in stats.values()) <= 1e-4


def test_deskew_across_yaw_wrap_and_extrapolation():
    """Euler lerp through +-180 deg (F2) and packets before / after the pose timeline."""
    pk, t = synth.hdl64_packets(1200)
    pt, trv = synth.ins_trajectory(20, t0_us=synth.T0_US + 60_000, yaw_amp_deg=0.0)
    trv[:, 5] = np.where(np.arange(20) % 2 == 0, 179.0, -179.0)   # flips sign every sample
    _run_streaming(pk, t, synth.calib_hdl64(), (pt, trv))


def test_pose_exact_hits_and_irregular_timeline():
    pk, t = synth.hdl64_packets(900)
    rng = np.random.default_rng(11)
    pt = np.sort(rng.choice(t, 40, replace=False)).astype(np.int64)   # poses exactly at packet times
    _, trv = synth.ins_trajectory(40)
    _run_streaming(pk, t, synth.calib_hdl64(), (pt, trv))


@pytest.mark.parametrize("n_poses", [0, 1])
def test_short_timeline_means_no_transform(n_poses):
    pk, t = synth.hdl64_packets(800)
    pt, trv = synth.ins_trajectory(max(n_poses, 1))
    poses = (pt[:n_poses], trv[:n_poses])
    _run_streaming(pk, t, synth.calib_hdl64(), poses, splits=(300,), tol=P.TOL_DECODE)


# --- carried state across batches ---------------------------------------------------------------------
def test_batches_with_carry_match_one_stream():
    pk, t = synth.hdl64_packets(2500)
    az = pk["blocks"]["azimuth"]
    wrap_pk = int(np.nonzero(az[1:, 0] < az[:-1, 11])[0][0]) + 1 if np.any(az[1:, 0] < az[:-1, 11]) \
        else int(np.nonzero(np.any(az[:, 1:] < az[:, :-1], axis=1))[0][0])
    poses = synth.ins_trajectory(100)
    # cuts: mid-frame, right after the packet that holds a wrap, single-packet batches
    splits = (7, 8, 333, wrap_pk, wrap_pk + 1, wrap_pk + 2, 1000, 1001, 2499)
    _run_streaming(pk, t, synth.calib_hdl64(), poses, splits=splits)


def test_wrap_inside_packet_streaming_quirks():
    """F4 a-d on a stream whose wraps fall in the middle of packets."""
    pk, t = synth.hdl32_packets(1500, az0=123.0)
    az = pk["blocks"]["azimuth"]
    inner = np.any(az[:, 1:] < az[:, :-1], axis=1)
    assert inner.sum() >= 5
    w = int(np.nonzero(inner)[0][2])
    poses = synth.ins_trajectory(100)
    _run_streaming(pk, t, synth.calib_hdl32(), poses, splits=(w, w + 1))


# --- adversarial segmentation -----------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_azimuths_streaming(seed):
    pk, t = synth.random_packets(1500, seed)
    poses = synth.ins_trajectory(60)
    _run_streaming(pk, t, synth.calib_hdl64(), poses, splits=(1, 513, 1024))


def test_decreasing_azimuths_chain_of_nonconstant_maps():
    """Every block wraps: long chains of non-constant skip maps across tiles."""
    n = 1300
    az = (35999 - (np.arange(12 * n) * 7) % 36000).astype(np.int64)
    rng = np.random.default_rng(9)
    pks = [make_packet(az[12 * i:12 * i + 12], None,
                       rng.integers(0, 3, (12, 32)).astype(np.uint16) * 1000) for i in range(n)]
    pk = np.concatenate(pks)
    t = synth.T0_US + 300 * np.arange(n, dtype=np.int64)
    _run_streaming(pk, t, synth.calib_identity(64), tol=0.0)


def test_long_chain_of_nonconstant_skip_maps():
    """Identical packets with a wrap at block 6: every packet's skip map is non-constant
    (f = [6 x7, 7, 8, 9, 10, 0]), so the tile look-back has to compose maps all the way back."""
    n = 1700
    az = np.array([700, 800, 900, 1000, 1100, 1200, 100, 200, 300, 400, 500, 600])
    rng = np.random.default_rng(4)
    pks = [make_packet(az, None, rng.integers(0, 2, (12, 32)).astype(np.uint16) * 2000)
           for _ in range(n)]
    pk = np.concatenate(pks)
    t = synth.T0_US + 300 * np.arange(n, dtype=np.int64)
    _run_streaming(pk, t, synth.calib_identity(64), synth.ins_trajectory(60), splits=(600,))


# --- filters ---------------------------------------------------------------------------------------------
def test_laser_selection_points_skip_and_crop():
    pk, t = synth.hdl64_packets(1200)
    sel = np.ones(64, np.int32)
    sel[[0, 5, 33, 63]] = 0
    poses = synth.ins_trajectory(60)
    _run_streaming(pk, t, synth.calib_hdl64(), poses, laser_selection=sel, points_skip=1)
    region = (-20.0, 20.0, -30.0, 30.0, -5.0, 5.0)
    for inside in (0, 1):
        _run_streaming(pk, t, synth.calib_hdl64(), poses, crop=(inside, region))


def test_calibration_with_fewer_rows_than_lasers_on_the_wire():
    """0xddff blocks with a 32-laser calibration: ids >= 32 are out of range in the reference
    (undefined behaviour there, dropped by the oracle and by the kernel)."""
    pk, t = synth.hdl64_packets(500)
    _run_streaming(pk, t, synth.calib_hdl32(), tol=0.0)


# --- edge sizes -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 511, 512, 513])
def test_ragged_sizes(n):
    pk, t = synth.hdl64_packets(n, az0=35900.0)
    _run_streaming(pk, t, synth.calib_hdl64(), synth.ins_trajectory(10))


# --- staged output of k_decode: whole 16-element granules by TMA bulk store, scalar head / tail ------
@pytest.mark.parametrize("n", [3, 7, 8, 9, 15, 16, 17, 23, 24, 25, 40])
def test_sizes_around_the_decode_tile(n):
    """k_decode tiles are 8 packets, a warp pair owns 2 of them: every remainder."""
    pk, t = synth.hdl64_packets(n, az0=35950.0, seed=n)
    _run_streaming(pk, t, synth.calib_hdl64(), synth.ins_trajectory(10))


@pytest.mark.parametrize("zero_frac", [0.5, 0.9, 0.97, 0.995])
def test_sparse_returns_partial_granules(zero_frac):
    """Few points per packet pair: ranges shorter than one 16-element granule (scalar stores
    only), ranges that straddle one granule boundary, pairs and packets that emit nothing."""
    pk, t = synth.hdl64_packets(700, zero_frac=zero_frac, seed=int(zero_frac * 1000))
    d = pk["blocks"]["returns"]["distance"]
    d[100:140] = 0          # whole packets without a point, inside and across tiles
    d[333] = 0
    _run_streaming(pk, t, synth.calib_hdl64(), synth.ins_trajectory(30), splits=(97, 98, 355))


def test_sparse_hdl32_and_random_streams():
    pk, t = synth.hdl32_packets(500, zero_frac=0.93, seed=93)
    _run_streaming(pk, t, synth.calib_hdl32(), tol=0.0)
    pk, t = synth.random_packets(300, seed=77, zero_frac=0.9)
    _run_streaming(pk, t, synth.calib_hdl64(), synth.ins_trajectory(20))


@pytest.mark.parametrize("halo", [349, 350, 351, 356])
def test_odd_and_even_halos(halo):
    """The point-offset column is copied from an even index: odd first packets shift it."""
    pk, t = synth.hdl64_packets(1500, seed=halo)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(60)
    ctx = P.make_ctx(calib, poses)
    whole = ctx.decode(b, t, t_base_us=int(t[0]))
    wc = whole.fetch()
    cut = 801
    r = ctx.decode(np.ascontiguousarray(b[cut - halo:]), np.ascontiguousarray(t[cut - halo:]),
                   n_halo=halo, t_base_us=int(t[0]))
    c = r.fetch()
    first = whole.n_points - r.n_points
    for k in wc:
        assert np.array_equal(wc[k][first:], c[k]), k
    ctx.close()


def test_all_zero_distances_emit_nothing():
    pk, t = synth.hdl64_packets(100, zero_frac=1.1)
    ctx = P.make_ctx(synth.calib_hdl64())
    r = ctx.decode(synth.as_bytes(pk), t)
    assert r.n_points == 0 and r.n_frames == 1 + r.n_closed
    ctx.close()


def test_decode_before_calibration_is_rejected():
    ctx = capi.Context(0, 1024, 16, 1)
    pk, t = synth.hdl64_packets(4)
    with pytest.raises(capi.VeloError) as e:
        ctx.decode(synth.as_bytes(pk), t)
    assert e.value.code == 2
    ctx.close()


# --- offline path (readFrameInformation + getFrame) -----------------------------------------------------------
def test_offline_mode_matches_get_frame():
    pk, t = synth.hdl64_packets(1500)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(60)
    o = P.make_oracle(calib, poses)
    sp, sk, ts = o.read_frame_information(b, t)
    ctx = P.make_ctx(calib, poses)
    gsp, gsk, gts = ctx.read_frame_information(b, t)
    assert np.array_equal(sp, gsp) and np.array_equal(sk, gsk) and np.array_equal(ts, gts)
    r = ctx.decode(b, t, mode=capi.MODE_OFFLINE, t_base_us=int(t[0]))
    cols = r.fetch()
    assert r.n_frames == len(sp)
    from veloslam_b200.frames import assemble_frame
    for i in range(len(sp)):
        f = o.get_frame(b, t, sp[i], sk[i])
        g = r.frames[i]
        assert g.n_points == f.n_points, i
        assert (g.start_packet, g.start_block) == ((-1, -1) if i == 0 else (sp[i], sk[i]))
        assert g.timestamp_us == ts[i] and g.skips == sk[i]
        assert np.allclose(g.carpose, f.carpose_TRV, atol=1e-9)
        c = {k: v[g.first_point:g.first_point + g.n_points] for k, v in cols.items()}
        fr_like = type("F", (), {"hdl64_order": f.is_hdl64_order})
        a = assemble_frame(c, fr_like, 64)
        assert np.array_equal(a.laser_counts, f.laser_counts), i
        assert np.array_equal(a.azimuth, f.azimuth), i
        d = np.abs(a.xyzi.astype(np.float64) - f.xyzi.astype(np.float64))
        assert d.max() <= P.TOL_DESKEW, (i, d.max())
    ctx.close()


def test_offline_random_azimuths():
    pk, t = synth.random_packets(600, 21)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(30)
    o = P.make_oracle(calib, poses)
    sp, sk, ts = o.read_frame_information(b, t)
    ctx = P.make_ctx(calib, poses)
    r = ctx.decode(b, t, mode=capi.MODE_OFFLINE, t_base_us=int(t[0]))
    assert r.n_frames == len(sp)
    cols = r.fetch()
    # several wraps per packet: index entry i is the blocks [start_i, start_{i+1}); getFrame's
    # frames.back() quirk (see Oracle.get_frame) is side-stepped with first=True
    for i in (0, 1, 2, len(sp) // 2, len(sp) - 1):
        f = o.get_frame(b, t, sp[i], sk[i], first=True)
        g = r.frames[i]
        assert g.n_points == f.n_points, i
        c = {k: v[g.first_point:g.first_point + g.n_points] for k, v in cols.items()}
        assert np.isclose(float(np.sum(c["x"].astype(np.float64))),
                          float(np.sum(f.xyzi[:, 0].astype(np.float64))), atol=1e-3 * max(1, g.n_points))
    ctx.close()


# --- input framing ----------------------------------------------------------------------------------------------
def test_pcap_record_stride_and_gpu_side_times():
    from veloslam_b200 import pcapio
    pk, t = synth.hdl64_packets(700)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(40)
    img = pcapio.write_pcap_image(synth.as_bytes(pk), t)
    recs, n = pcapio.payload_view(img)
    assert n == 700
    o = P.make_oracle(calib, poses)
    o.trace_enable()
    o.process_packets(synth.as_bytes(pk), t)
    ctx = P.make_ctx(calib, poses)
    r = ctx.decode(recs, None, n=n, stride=pcapio.RECORD_BYTES, flags=capi.FLAG_PCAP_TIMES)
    batches = [(r, r.fetch(), 0)]
    P.assert_stream_parity(o, batches, P.TOL_DESKEW, t, calib=calib)
    ctx.close()


# --- halo shards (multi-GPU decomposition on one device) ----------------------------------------------------------
@pytest.mark.parametrize("mode", [capi.MODE_STREAMING, capi.MODE_OFFLINE])
def test_halo_shard_matches_whole_stream(mode):
    pk, t = synth.hdl64_packets(3000)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(120)
    ctx = P.make_ctx(calib, poses)
    whole = ctx.decode(b, t, mode=mode, t_base_us=int(t[0]))
    wc = whole.fetch()
    cut, halo = 1700, 400
    r = ctx.decode(np.ascontiguousarray(b[cut - halo:]), np.ascontiguousarray(t[cut - halo:]),
                   n_halo=halo, mode=mode, t_base_us=int(t[0]))
    c = r.fetch()
    # points of packets >= cut are the tail of the whole-stream output
    first = whole.n_points - r.n_points
    for k in wc:
        assert np.array_equal(wc[k][first:], c[k]), k
    assert r.frames[-1].n_points == whole.frames[-1].n_points
    assert r.frames[-1].timestamp_us == whole.frames[-1].timestamp_us
    assert r.frames[0].timestamp_us == whole.frames[whole.n_frames - r.n_frames].timestamp_us
    ctx.close()


def test_halo_without_wrap_is_rejected():
    pk, t = synth.hdl64_packets(600)
    ctx = P.make_ctx(synth.calib_hdl64(), synth.ins_trajectory(30))
    with pytest.raises(capi.VeloError) as e:
        ctx.decode(np.ascontiguousarray(synth.as_bytes(pk)[100:]), np.ascontiguousarray(t[100:]),
                   n_halo=50, t_base_us=int(t[0]))
    assert e.value.code == 6
    ctx.close()


# --- size-independent properties at scale (config 2/3 sizes) ---------------------------------------------------------
def test_full_minute_hdl64_properties():
    """60 s of HDL-64E (208 320 packets, 80.0 M slots): counts and frame structure against
    closed forms of the input; per-point check against the oracle on a window."""
    n = 208_320
    pk, t = synth.hdl64_packets(n)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl64()
    poses = synth.ins_trajectory(6100)
    ctx = P.make_ctx(calib, poses, max_batch_packets=n)
    r = ctx.decode(b, t, t_base_us=int(t[0]))
    d = pk["blocks"]["returns"]["distance"]
    az = pk["blocks"]["azimuth"].reshape(-1).astype(np.int64)
    wraps = np.nonzero(az[1:] < az[:-1])[0] + 1
    assert r.n_closed == len(wraps) and 598 <= r.n_closed <= 601
    nz = np.count_nonzero(d.reshape(-1, 32), axis=1)
    cum = np.concatenate([[0], np.cumsum(nz)])
    # streaming drops blocks [0, k) of the packet after a wrap at block k (F4a)
    dropped = 0
    for i, w in enumerate(wraps):
        p, k = divmod(int(w), 12)
        assert r.frames[i + 1].start_packet == p and r.frames[i + 1].start_block == k
        if k and p + 1 < n:
            dropped += int(nz[(p + 1) * 12:(p + 1) * 12 + k].sum())
    assert r.n_points == int(cum[-1]) - dropped
    assert sum(f.n_points for f in r.frames) == r.n_points
    assert all(int(f.laser_counts.sum()) == f.n_points for f in r.frames)
    # window parity against the oracle (first 3000 packets)
    o = P.make_oracle(calib, poses)
    o.trace_enable()
    o.process_packets(b[:3000], t[:3000])
    tr = o.trace()
    m = len(tr["x"])
    c = r.fetch(0, m)
    for k in ("laser", "intensity", "azimuth", "distance"):
        assert np.array_equal(c[k], tr[k])
    for k in ("x", "y", "z"):
        assert np.max(np.abs(c[k].astype(np.float64) - tr[k])) <= P.TOL_DESKEW
    ctx.close()


def test_config1_hdl32_minute_pcap_file_no_poses():
    """BASELINE.json configs[0] at full size: 60 s of HDL-32E (108 480 packets, 41.66 M slots) in
    vtkPacketFileWriter format, decoded in place out of the file image (1264-byte records, record
    timestamps on the GPU), no poses.  Bit-exact against the oracle end to end."""
    from veloslam_b200 import pcapio
    n = 108_480
    pk, t = synth.hdl32_packets(n)
    b = synth.as_bytes(pk)
    calib = synth.calib_hdl32()
    img = pcapio.write_pcap_image(b, t)
    assert img.size == 137_118_744
    recs, nrec = pcapio.payload_view(img)
    assert nrec == n
    ctx = P.make_ctx(calib, max_batch_packets=n)
    r = ctx.decode(recs, None, n=n, stride=pcapio.RECORD_BYTES, flags=capi.FLAG_PCAP_TIMES)
    assert 599 <= r.n_closed <= 601
    o = P.make_oracle(calib)
    o.trace_enable()
    o.process_packets(b, t)
    tr = o.trace()
    c = r.fetch()
    assert r.n_points == len(tr["x"])
    for k in ("laser", "intensity", "azimuth", "distance"):
        assert np.array_equal(c[k], tr[k]), k
    for k in ("x", "y", "z"):                            # rotCorrection == 0: LUT branch, bit-exact
        assert np.array_equal(c[k].view(np.uint32), tr[k].astype(np.float32).view(np.uint32)), k
    of = o.frames()
    assert len(of) == r.n_closed
    assert [f.n_points for f in of] == [fr.n_points for fr in r.frames[:r.n_closed]]
    assert [f.timestamp_us for f in of[1:]] == [fr.timestamp_us for fr in r.frames[1:r.n_closed]]
    ctx.close()
