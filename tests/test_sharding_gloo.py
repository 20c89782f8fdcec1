"""Host-side multi-GPU logic on CPU: shard ranges, the frame-index all-gather (gloo,
world_size 2) and the stitch of the rotation that straddles a shard boundary."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from veloslam_b200 import capi, sharding


def fake_table(rows):
    t = np.zeros(len(rows), dtype=capi.FRAME_TABLE_DTYPE)
    for i, (n, first, sp, sb, ts, closed) in enumerate(rows):
        t[i]["n_points"], t[i]["first_point"] = n, first
        t[i]["start_packet"], t[i]["start_block"] = sp, sb
        t[i]["timestamp_us"], t[i]["closed"] = ts, closed
        t[i]["meta_packet"] = sp + 1 if sp >= 0 else -1
        t[i]["skips"] = max(sb, 0)
    return t


def rank_tables():
    # rank 0 decodes packets [0, 1000): frame 0 (partial), frame 1 closed, frame 2 open
    r0 = fake_table([(5000, 0, -1, -1, 100, 1), (9000, 5000, 300, 4, 400, 1), (2000, 14000, 650, 2, 750, 0)])
    # rank 1 decodes [1000, 2000) with a 400-packet halo starting at packet 600: its first frame
    # continues rank 0's open one (same timestamp, rebuilt from the halo)
    r1 = fake_table([(7000, 0, -1, -1, 750, 1), (8000, 7000, 400, 6, 1100, 0)])
    return [sharding.local_table(r0, 0, 0, 0), sharding.local_table(r1, 1, 1000, 400)]


def test_shard_ranges_cover_the_recording():
    r = sharding.shard_ranges(12_499_200, 8, sharding.HALO_HDL64)
    assert r[0] == (0, 0, 1_562_400)
    assert all(a[2] == b[0] for a, b in zip(r[:-1], r[1:])) and r[-1][2] == 12_499_200
    assert all(h == sharding.HALO_HDL64 for _, h, _ in r[1:])
    assert sharding.shard_ranges(100, 3, 512)[1] == (33, 33, 66)      # halo clipped at the start


def test_local_table_uses_global_packet_indices():
    t = rank_tables()[1]
    assert t[1, sharding.COLS.index("start_packet")] == 1000 - 400 + 400
    assert t[0, sharding.COLS.index("start_packet")] == -1
    assert np.all(t[:, sharding.COLS.index("rank")] == 1)


def test_stitch_merges_the_straddling_rotation():
    frames = sharding.stitch(rank_tables())
    assert [f["n_points"] for f in frames] == [5000, 9000, 2000 + 7000, 8000]
    assert frames[2]["segments"] == [(0, 14000, 2000), (1, 0, 7000)]
    assert frames[2]["closed"] and not frames[3]["closed"]
    assert frames[2]["start_packet"] == 650 and frames[2]["timestamp_us"] == 750
    assert not any("timestamp_mismatch" in f for f in frames)
    assert sum(f["n_points"] for f in frames) == 31000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tables = sharding.all_gather_tables(rank_tables()[rank])
    frames = sharding.stitch(tables)
    out[rank] = [f["n_points"] for f in frames]
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_of_frame_tables_over_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == out[1] == [5000, 9000, 9000, 8000]


def test_stitch_arrays_frame_spanning_three_ranks_and_mismatch_flag():
    """The C-ABI stitch (vs_stitch_frame_tables): a shard with no wrap of its own is the middle
    of a frame that spans three ranks; a boundary whose two sides disagree on the frame's
    timestamp is flagged."""
    r0 = fake_table([(100, 0, -1, -1, 10, 1), (50, 100, 20, 3, 77, 0)])
    r1 = fake_table([(60, 0, -1, -1, 77, 0)])                        # no wrap in this shard
    r2 = fake_table([(70, 0, -1, -1, 77, 1), (5, 70, 90, 0, 999, 0)])
    r3 = fake_table([(9, 0, -1, -1, 1234, 0)])                       # disagrees with rank 2's 999
    tabs = [sharding.local_table(t, g, 100 * g, 0) for g, t in enumerate((r0, r1, r2, r3))]
    gf, segs = sharding.stitch_arrays(tabs)
    assert gf["n_points"].tolist() == [100, 180, 14]
    assert gf["n_segments"].tolist() == [1, 3, 2] and gf["first_segment"].tolist() == [0, 1, 4]
    assert segs["rank"].tolist() == [0, 0, 1, 2, 2, 3]
    assert gf["closed"].tolist() == [1, 1, 0]
    assert gf["timestamp_mismatch"].tolist() == [0, 0, 1]
    assert gf["start_packet"].tolist() == [-1, 20, 290]
    # empty world / empty tables
    gf, segs = sharding.stitch_arrays([np.zeros((0, 10), np.int64)])
    assert len(gf) == 0 and len(segs) == 0


def test_shard_range_edge_cases():
    assert capi.shard_range(0, 4, 2, 512) == (0, 0, 0)
    assert capi.shard_range(7, 8, 7, 512) == (6, 6, 7)
    big = (1 << 62) + 12345
    f, h, e = capi.shard_range(big, 8, 7, 512)
    assert e == big and f == (big * 7) // 8 and h == 512         # no 64-bit overflow in n * rank
    with pytest.raises(capi.VeloError):
        capi.shard_range(100, 4, 4, 0)


def test_stitcher_reads_the_all_gather_buffer_in_place():
    """Stitcher: table g at row g * stride (+ a header row) of one fixed-size gathered buffer."""
    tabs = rank_tables()
    cap = 8
    buf = np.full((2, cap + 1, capi.FRAME_ROW_COLS), -7, dtype=np.int64)
    for g, t in enumerate(tabs):
        buf[g, 0, 0] = t.shape[0]
        buf[g, 1:1 + t.shape[0]] = t
    st = capi.Stitcher(2, 2 * cap)
    gf, segs = st(buf, buf[:, 0, 0], cap + 1, first_row=1)
    want_f, want_s = sharding.stitch_arrays(tabs)
    assert np.array_equal(gf, want_f) and np.array_equal(segs, want_s)
    with pytest.raises(capi.VeloError):
        st(buf, [cap + 2, 1], cap + 1, first_row=1)      # more rows than the stride holds
