"""Deterministic synthetic inputs for the ingest hot path.

Velodyne HDL-32E / HDL-64E S2 packet streams in the wire layout the reference parses
(HDLParser.cxx:61-87: 12 x {u16 blockId, u16 rotationalPosition, 32 x {u16 distance,
u8 intensity}} + u32 gpsTimestamp + 2 status bytes = 1206 B), per-laser calibration tables
in the raw db.xml units (HDLParser.cxx:822-839) and a 100 Hz INS trajectory in the
PoseTransform convention (type_defs.h:86-131: T = ENU metres, R = roll/pitch/yaw degrees,
V = ENU m/s).  Everything is seeded; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import numpy as np

PACKET_BYTES = 1206
BLOCKS_PER_PACKET = 12
RETURNS_PER_BLOCK = 32
SLOTS_PER_PACKET = BLOCKS_PER_PACKET * RETURNS_PER_BLOCK  # 384
BLOCK_LOWER = 0xEEFF
BLOCK_UPPER = 0xDDFF

RETURN_DTYPE = np.dtype([("distance", "<u2"), ("intensity", "u1")])
BLOCK_DTYPE = np.dtype([("id", "<u2"), ("azimuth", "<u2"), ("returns", RETURN_DTYPE, (32,))])
PACKET_DTYPE = np.dtype([("blocks", BLOCK_DTYPE, (12,)), ("gps", "<u4"), ("status", "u1", (2,))])
assert PACKET_DTYPE.itemsize == PACKET_BYTES

# Default epoch for synthetic streams: 2016-07-01 00:00:00 UTC in microseconds.
T0_US = 1467331200 * 1_000_000

HDL64_US_PER_PACKET = 288      # ~3472 packets/s
HDL32_US_PER_PACKET = 553      # ~1808 packets/s (HDLParser.cxx:1121 uses 553 us too)
HDL64_TICKS_PER_PAIR = 17.28   # 0.01 deg ticks between consecutive (lower, upper) pairs
HDL32_TICKS_PER_BLOCK = 16.59

# HDL-64 beam re-order table applied by the reference when it closes a frame
# (HDLParser.cxx:179-182); new[i] = old[LUT[i]].
HDL64_BEAM_LUT = np.array(
    [38, 39, 42, 43, 32, 33, 36, 37, 40, 41, 46, 47, 50, 51, 54, 55, 44, 45, 48, 49, 52, 53, 58, 59,
     62, 63, 34, 35, 56, 57, 60, 61, 6, 7, 10, 11, 0, 1, 4, 5, 8, 9, 14, 15, 18, 19, 22, 23, 12, 13,
     16, 17, 20, 21, 26, 27, 30, 31, 2, 3, 24, 25, 28, 29], dtype=np.int32)


class Calibration:
    """Per-laser calibration in raw db.xml units plus the enabled-laser count."""

    def __init__(self, rot_deg, vert_deg, dist_cm, voff_cm, hoff_cm, n_enabled):
        self.rot_deg = np.array(rot_deg, dtype=np.float64, copy=True)
        self.vert_deg = np.array(vert_deg, dtype=np.float64, copy=True)
        self.dist_cm = np.array(dist_cm, dtype=np.float64, copy=True)
        self.voff_cm = np.array(voff_cm, dtype=np.float64, copy=True)
        self.hoff_cm = np.array(hoff_cm, dtype=np.float64, copy=True)
        self.n_enabled = int(n_enabled)
        n = self.rot_deg.shape[0]
        assert n <= 64 and all(a.shape == (n,) for a in
                               (self.vert_deg, self.dist_cm, self.voff_cm, self.hoff_cm))

    @property
    def n_rows(self):
        return int(self.rot_deg.shape[0])

    def padded64(self):
        """(5, 64) float64, rows beyond n_rows zero."""
        out = np.zeros((5, 64), dtype=np.float64)
        for i, a in enumerate((self.rot_deg, self.vert_deg, self.dist_cm, self.voff_cm, self.hoff_cm)):
            out[i, :a.shape[0]] = a
        return out


def calib_identity(n_lasers=64, n_enabled=None):
    z = np.zeros(n_lasers)
    return Calibration(z, z, z, z, z, n_lasers if n_enabled is None else n_enabled)


def calib_hdl32():
    """32 lasers, rotCorrection = 0 (LUT branch), vertCorrection -30.67..+10.67 in 1.333 deg steps."""
    vert = -30.67 + (41.34 / 31.0) * np.arange(32)
    z = np.zeros(32)
    return Calibration(z, vert, z, z, z, 32)


def calib_hdl64(seed=0xC0FFEE):
    """64 lasers with non-zero rotCorrection (libm branch of the reference)."""
    rng = np.random.default_rng(seed)
    rot = rng.uniform(-5.0, 5.0, 64)
    vert = np.linspace(-24.8, 2.0, 64)[rng.permutation(64)]
    dist = rng.uniform(90.0, 160.0, 64)
    voff = rng.uniform(18.0, 22.0, 64)
    hoff = np.where(np.arange(64) % 2 == 0, 2.6, -2.6)
    return Calibration(rot, vert, dist, voff, hoff, 64)


def _fill_returns(pk, rng, zero_frac, dist_lo, dist_hi):
    n = pk.shape[0]
    shape = (n, 12, 32)
    d = rng.integers(dist_lo, dist_hi + 1, size=shape, dtype=np.uint16)
    if zero_frac > 0:
        d[rng.random(shape, dtype=np.float32) < zero_frac] = 0
    pk["blocks"]["returns"]["distance"] = d
    pk["blocks"]["returns"]["intensity"] = rng.integers(0, 256, size=shape, dtype=np.uint8)


def hdl64_packets(n_packets, seed=0xC0FFEE, first_packet=0, az0=12345.0, t0_us=T0_US,
                  zero_frac=0.05, dist_lo=500, dist_hi=60000):
    """HDL-64E S2 stream: packets of 6 x (0xeeff, 0xddff) pairs sharing one azimuth.

    Packets are a pure function of (seed, first_packet + index) for azimuth/time, so a shard
    generated with first_packet=k continues the stream of a shard that ended at k.
    Returns (packets[n] PACKET_DTYPE, t_us[n] int64).
    """
    rng = np.random.default_rng([seed, first_packet])
    pk = np.zeros(n_packets, dtype=PACKET_DTYPE)
    idx = np.arange(first_packet, first_packet + n_packets, dtype=np.int64)
    pair = idx[:, None] * 6 + np.arange(6)[None, :]
    az = np.floor(az0 + pair * HDL64_TICKS_PER_PAIR).astype(np.int64) % 36000
    pk["blocks"]["azimuth"] = np.repeat(az, 2, axis=1).astype(np.uint16)
    pk["blocks"]["id"] = np.tile(np.array([BLOCK_LOWER, BLOCK_UPPER], dtype=np.uint16), 6)[None, :]
    _fill_returns(pk, rng, zero_frac, dist_lo, dist_hi)
    t_us = t0_us + idx * HDL64_US_PER_PACKET
    pk["gps"] = (t_us % 3_600_000_000).astype(np.uint32)
    return pk, t_us


def hdl32_packets(n_packets, seed=0x32E, first_packet=0, az0=777.0, t0_us=T0_US,
                  zero_frac=0.05, dist_lo=500, dist_hi=50000):
    """HDL-32E stream: all blocks 0xeeff, azimuth advancing 16.59 ticks/block."""
    rng = np.random.default_rng([seed, first_packet])
    pk = np.zeros(n_packets, dtype=PACKET_DTYPE)
    idx = np.arange(first_packet, first_packet + n_packets, dtype=np.int64)
    blk = idx[:, None] * 12 + np.arange(12)[None, :]
    az = np.floor(az0 + blk * HDL32_TICKS_PER_BLOCK).astype(np.int64) % 36000
    pk["blocks"]["azimuth"] = az.astype(np.uint16)
    pk["blocks"]["id"] = BLOCK_LOWER
    _fill_returns(pk, rng, zero_frac, dist_lo, dist_hi)
    t_us = t0_us + idx * HDL32_US_PER_PACKET
    pk["gps"] = (t_us % 3_600_000_000).astype(np.uint32)
    return pk, t_us


def random_packets(n_packets, seed, t0_us=T0_US, upper_frac=0.3, zero_frac=0.3):
    """Adversarial stream: random azimuths (wraps everywhere), random block ids."""
    rng = np.random.default_rng(seed)
    pk = np.zeros(n_packets, dtype=PACKET_DTYPE)
    pk["blocks"]["azimuth"] = rng.integers(0, 36000, size=(n_packets, 12), dtype=np.uint16)
    ids = np.where(rng.random((n_packets, 12)) < upper_frac, BLOCK_UPPER, BLOCK_LOWER)
    pk["blocks"]["id"] = ids.astype(np.uint16)
    _fill_returns(pk, rng, zero_frac, 0, 65535)
    t_us = t0_us + np.arange(n_packets, dtype=np.int64) * 300
    pk["gps"] = (t_us % 3_600_000_000).astype(np.uint32)
    return pk, t_us


def as_bytes(packets):
    """View a PACKET_DTYPE array as (n, 1206) uint8."""
    return np.ascontiguousarray(packets).view(np.uint8).reshape(-1, PACKET_BYTES)


def ins_trajectory(n_poses, t0_us=T0_US - 50_000, dt_us=10_000, seed=7, yaw_amp_deg=60.0,
                   yaw_period_s=20.0, speed=10.0, yaw_offset_deg=0.0):
    """100 Hz INS samples: (t_us[n] int64, TRV[n, 9] float64 = T(3), R(3) deg, V(3)).

    10 m/s along the heading, yaw = offset + amp * sin(2 pi t / period) (|rate| <= 18.9 deg/s
    at the defaults), roll/pitch +-3 deg sinusoids.
    """
    rng = np.random.default_rng(seed)
    ph = rng.uniform(0, 2 * np.pi, 3)
    i = np.arange(n_poses, dtype=np.int64)
    t_us = t0_us + i * dt_us
    ts = (i * dt_us) / 1e6
    yaw = yaw_offset_deg + yaw_amp_deg * np.sin(2 * np.pi * ts / yaw_period_s + ph[0])
    roll = 3.0 * np.sin(2 * np.pi * ts / 7.0 + ph[1])
    pitch = 3.0 * np.sin(2 * np.pi * ts / 11.0 + ph[2])
    heading = np.deg2rad(yaw)
    v = np.stack([speed * np.cos(heading), speed * np.sin(heading), 0.2 * np.cos(ts)], axis=1)
    T = np.concatenate([np.zeros((1, 3)), np.cumsum(v[:-1] * (dt_us / 1e6), axis=0)], axis=0)
    T += np.array([431000.0, 3391000.0, 12.0])[None, :] * 0.001  # a non-zero ENU origin offset
    trv = np.concatenate([T, np.stack([roll, pitch, yaw], axis=1), v], axis=1)
    return t_us, np.ascontiguousarray(trv, dtype=np.float64)


def hdl64_stream_tiled(n_packets, first_packet=0, base_packets=16384, seed=0xC0FFEE, az0=12345.0,
                       t0_us=T0_US, zero_frac=0.05):
    """A long HDL-64E stream built quickly: azimuths and times are those of packets
    [first_packet, first_packet + n) of the infinite stream; the returns repeat a seeded base
    block of `base_packets` packets (phase-locked to the global packet index, so two shards of
    the same stream agree where they overlap)."""
    base, _ = hdl64_packets(base_packets, seed=seed, zero_frac=zero_frac)
    idx = np.arange(first_packet, first_packet + n_packets, dtype=np.int64)
    pk = np.take(as_bytes(base), idx % base_packets, axis=0).view(PACKET_DTYPE).reshape(-1)
    pair = idx[:, None] * 6 + np.arange(6)[None, :]
    az = np.floor(az0 + pair * HDL64_TICKS_PER_PAIR).astype(np.int64) % 36000
    pk["blocks"]["azimuth"] = np.repeat(az, 2, axis=1).astype(np.uint16)
    t_us = t0_us + idx * HDL64_US_PER_PACKET
    pk["gps"] = (t_us % 3_600_000_000).astype(np.uint32)
    return pk, t_us
