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
    """[(first, n_halo, end)] per rank: decode [first, end), state rebuilt from [first-n_halo, first)."""
    out = []
    for g in range(world):
        first = (n_packets * g) // world
        end = (n_packets * (g + 1)) // world
        h = min(halo, first)
        out.append((first, h, end))
    return out


def local_table(frame_table, rank, first_packet, n_halo):
    """vs_frame rows of one shard -> (n_frames, NCOL) int64 with GLOBAL packet indices."""
    n = frame_table.shape[0]
    t = np.zeros((n, NCOL), dtype=np.int64)
    sp = frame_table["start_packet"].astype(np.int64)
    mp = frame_table["meta_packet"].astype(np.int64)
    base = first_packet - n_halo               # index 0 of the submitted array
    t[:, 0] = frame_table["n_points"]
    t[:, 1] = frame_table["first_point"]
    t[:, 2] = np.where(sp >= 0, sp + base, -1)
    t[:, 3] = frame_table["start_block"]
    t[:, 4] = frame_table["timestamp_us"]
    t[:, 5] = frame_table["skips"]
    t[:, 6] = frame_table["closed"]
    t[:, 7] = frame_table["hdl64_order"]
    t[:, 8] = np.where(mp >= 0, mp + base, mp)
    t[:, 9] = rank
    return t


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


def stitch(tables):
    """Global frame index from per-rank tables (rank order).

    The open (last) frame of rank g and the first frame of rank g+1 are the same rotation;
    they are merged into one global frame whose points live in two ranks.  Returns a list of
    dicts: {"segments": [(rank, first_point, n_points)], "n_points", "start_packet",
    "start_block", "timestamp_us", "skips", "closed", "hdl64_order"}.
    """
    frames = []
    for g, t in enumerate(tables):
        for i in range(t.shape[0]):
            row = t[i]
            seg = (int(row[9]), int(row[1]), int(row[0]))
            cont = (i == 0 and g > 0 and frames)
            if cont:
                f = frames[-1]
                f["segments"].append(seg)
                f["n_points"] += int(row[0])
                f["closed"] = bool(row[6])
                f["hdl64_order"] = bool(row[7])
                # meta comes from wherever the frame started; both sides agree (the shard
                # rebuilt it from its halo) -- keep the owner's and check the timestamp
                if int(row[4]) != f["timestamp_us"]:
                    f["timestamp_mismatch"] = (f["timestamp_us"], int(row[4]))
            else:
                frames.append({"segments": [seg], "n_points": int(row[0]),
                               "start_packet": int(row[2]), "start_block": int(row[3]),
                               "timestamp_us": int(row[4]), "skips": int(row[5]),
                               "closed": bool(row[6]), "hdl64_order": bool(row[7])})
    return frames
