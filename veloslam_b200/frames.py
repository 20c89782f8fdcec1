"""Host-side assembly of reference-shaped frames from the C-ABI's stream-order columns.

The GPU emits points in the order the reference pushes them (packet, block, return slot) and
a frame index; the reference's HDLFrame holds one list per laser (HDLFrame.h:17-18), re-ordered
by HDL64BeamLUT when the frame is closed on HDL-64 data (HDLParser.cxx:880-893).  This is the
same scatter the C++ facade does (veloslam_b200/cpp/HDLParser.cpp).
"""
from __future__ import annotations

import numpy as np

from .synth import HDL64_BEAM_LUT


class AssembledFrame:
    def __init__(self, xyzi, azimuth, distance_raw, laser_counts, laser_rows):
        self.xyzi = xyzi                    # (n, 4) float32, laser-major
        self.azimuth = azimuth              # (n,) uint16
        self.distance_raw = distance_raw    # (n,) uint16
        self.laser_counts = laser_counts    # per final row
        self.laser_rows = laser_rows        # raw laser id of each final row


def assemble_frame(cols, frame, n_lasers):
    """cols: dict of numpy columns covering exactly the frame's point range."""
    laser = cols["laser"].astype(np.int64)
    order = np.argsort(laser, kind="stable")        # time order is kept inside each laser
    rows = np.arange(max(n_lasers, 1), dtype=np.int64)
    if frame.hdl64_order:
        # new[i] = old[LUT[i]]
        rows = HDL64_BEAM_LUT.astype(np.int64)
        rank = np.empty(64, dtype=np.int64)
        rank[rows] = np.arange(64)
        order = np.argsort(rank[laser], kind="stable")
    xyzi = np.stack([cols["x"][order], cols["y"][order], cols["z"][order],
                     cols["intensity"][order].astype(np.float32)], axis=1)
    counts = np.bincount(laser, minlength=64)[rows] if len(rows) else np.zeros(0, np.int64)
    return AssembledFrame(xyzi, cols["azimuth"][order], cols["distance"][order], counts, rows)


def point_meta_distance(distance_raw, laser_row, dist_correction_cm):
    """PointMeta::distance (HDLParser.cxx:614, 747): float(dist * 0.002 + distCorrection/100)."""
    corr = np.asarray(dist_correction_cm, dtype=np.float64) / 100.0
    return (distance_raw.astype(np.float64) * 0.002 + corr[laser_row]).astype(np.float32)
