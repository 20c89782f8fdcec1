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
