"""Shared helpers for the parity tests."""
import numpy as np

from veloslam_b200 import synth


def make_packet(az, ids=None, dist=None, inten=None, gps=0):
    """One 1206-byte packet from 12 block azimuths (+ optional ids / returns)."""
    pk = np.zeros(1, dtype=synth.PACKET_DTYPE)
    pk["blocks"]["azimuth"][0] = np.asarray(az, dtype=np.uint16)
    pk["blocks"]["id"][0] = synth.BLOCK_LOWER if ids is None else np.asarray(ids, dtype=np.uint16)
    if dist is not None:
        pk["blocks"]["returns"]["distance"][0] = np.asarray(dist, dtype=np.uint16)
    if inten is not None:
        pk["blocks"]["returns"]["intensity"][0] = np.asarray(inten, dtype=np.uint8)
    pk["gps"] = gps
    return pk


def single_return_packet(az_ticks, dsr, dist, inten=7, block=0, upper=False, step=10):
    """A packet whose only non-zero return is (block, dsr); azimuths rise by `step`."""
    az = (az_ticks - block * step + step * np.arange(12)) % 36000
    d = np.zeros((12, 32), dtype=np.uint16)
    d[block, dsr] = dist
    i = np.zeros((12, 32), dtype=np.uint8)
    i[block, dsr] = inten
    ids = np.full(12, synth.BLOCK_LOWER, dtype=np.uint16)
    if upper:
        ids[block] = synth.BLOCK_UPPER
    return make_packet(az, ids, d, i)
