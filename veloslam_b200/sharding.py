"""Packet-range sharding of a long recording across GPUs.

Packets are independent given four carried scalars (lastAzimuth, firingSkip, the open frame's
origin/meta packet, sticky isHDL64Data), so rank g decodes the contiguous range
[g*N/G, (g+1)*N/G) preceded by a halo of H packets (>= one rotation) that only rebuilds that
state.  There is no data-path collective: after the local decode each rank holds a frame-index
table; one all-gather of those tables (NCCL over NVLink on the GPU box, gloo in the CPU tests)
plus a local prefix turns local frame ids / point offsets into global ones and stitches the
frame that straddles each shard boundary.
"""
from __future__ import annotations

import numpy as np

HALO_HDL64 = 512   # > 348 packets = one 10 Hz HDL-64E rotation
HALO_HDL32 = 256   # > 181 packets = one 10 Hz HDL-32E rotation

# Columns of the exchanged per-frame rows (int64).
COLS = ("n_points", "first_point", "start_packet", "start_block", "timestamp_us", "skips",
        "closed", "hdl64_order", "meta_packet", "rank")
NCOL = len(COLS)


def shard_ranges(n_packets, world, halo):
    """[(first, n_halo, end)] per rank: decode [first, end), state rebuilt from [first-n_halo, first).
    (vs_shard_range of the C ABI.)"""
    from . import capi
    return [capi.shard_range(n_packets, world, g, halo) for g in range(world)]


def local_table(frame_table, rank, first_packet, n_halo):
    """vs_frame rows of one shard -> (n_frames, NCOL) int64 with GLOBAL packet indices
    (vs_frame_table_rows of the C ABI)."""
    from . import capi
    return capi.frame_table_rows(frame_table, rank, first_packet, n_halo)


def all_gather_tables(table, group=None):
    """All-gather variable-length (n, NCOL) int64 tables with torch.distributed; returns the
    list of per-rank tables.  Works on NCCL (tensors on the current CUDA device) and gloo."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) \
        if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = torch.tensor([table.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts + [1])
    mine = torch.zeros((m, NCOL), dtype=torch.int64, device=dev)
    if table.shape[0]:
        mine[:table.shape[0]] = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    bufs = [torch.zeros((m, NCOL), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    return [b[:c].cpu().numpy() for b, c in zip(bufs, counts)]


def stitch_arrays(tables):
    """Global frame index from per-rank tables (rank order) as structured arrays
    (vs_global_frame rows, vs_frame_segment rows): vs_stitch_frame_tables of the C ABI."""
    from . import capi
    return capi.stitch_frame_tables(tables)


def stitch(tables):
    """The same index as a list of dicts (tests, small tables): built from stitch_arrays().

    The open (last) frame of rank g and the first frame of rank g+1 are the same rotation;
    they are merged into one global frame whose points live in two ranks.  Returns a list of
    dicts: {"segments": [(rank, first_point, n_points)], "n_points", "start_packet",
    "start_block", "timestamp_us", "skips", "closed", "hdl64_order"}.
    """
    gf, segs = stitch_arrays(tables)
    frames = []
    for f in gf:
        a, n = int(f["first_segment"]), int(f["n_segments"])
        d = {"segments": [(int(sg["rank"]), int(sg["first_point"]), int(sg["n_points"]))
                          for sg in segs[a:a + n]],
             "n_points": int(f["n_points"]), "start_packet": int(f["start_packet"]),
             "start_block": int(f["start_block"]), "timestamp_us": int(f["timestamp_us"]),
             "skips": int(f["skips"]), "closed": bool(f["closed"]),
             "hdl64_order": bool(f["hdl64_order"])}
        if f["timestamp_mismatch"]:
            d["timestamp_mismatch"] = True
        frames.append(d)
    return frames
