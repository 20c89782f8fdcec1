#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Runs /root/reference's own HDLParser / TransformManager (oracle/_ref, compiled verbatim against
oracle/ref_shim) on small seeded inputs and stores inputs + outputs.  The reference cannot
travel to the GPU box; these files can.  Re-run only in the build container:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref  # noqa: E402
from veloslam_b200 import synth  # noqa: E402


def run_case(name, pk, t, calib, poses, sel=None, skip=0, crop=None):
    r = ref.RefParser()
    r.set_calibration(calib)
    if sel is not None:
        r.set_laser_selection(sel)
    r.set_points_skip(skip)
    if crop is not None:
        r.set_crop(1, crop[0], crop[1])
    if poses is not None:
        r.add_poses(poses[0], poses[1])
    b = synth.as_bytes(pk)
    r.process_packets(b, t)
    state = r.state()
    frames = r.frames()            # closed frames only: what getAllFrames() returns
    out = {
        "packets": b, "t_us": t, "calib": calib.padded64(), "calib_rows": calib.n_rows,
        "n_enabled": calib.n_enabled,
        "pose_t": poses[0] if poses is not None else np.zeros(0, np.int64),
        "pose_trv": poses[1] if poses is not None else np.zeros((0, 9)),
        "laser_selection": np.ones(64, np.int32) if sel is None else sel,
        "points_skip": skip, "crop": np.array([0] if crop is None else [1]),
        "crop_inside": np.array([0 if crop is None else crop[0]]),
        "crop_region": np.zeros(6) if crop is None else np.array(crop[1], dtype=np.float64),
        "n_frames": len(frames),
        "state": np.array([state["last_azimuth"], state["firing_skip"],
                           int(state["frame_meta_inited"]), int(state["is_hdl64"])]),
        "open_frame_points": r.open_frame_points(),
    }
    for i, f in enumerate(frames):
        out[f"f{i}_xyzi"] = f.xyzi
        out[f"f{i}_azimuth"] = f.azimuth
        out[f"f{i}_distance"] = f.distance
        out[f"f{i}_laser_counts"] = f.laser_counts
        out[f"f{i}_meta"] = np.array([f.timestamp_us, f.skips, f.n_packets, int(f.carpose_valid)],
                                     dtype=np.int64)
        out[f"f{i}_carpose"] = f.carpose_TRV
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, len(frames), "frames", sum(f.n_points for f in frames), "points")


def main():
    assert ref.available(), "oracle/_ref is not built"
    # two full HDL-64 rotations need ~700 packets; start close to the wrap to keep it small
    pk, t = synth.hdl64_packets(420, az0=33000.0, zero_frac=0.75)
    run_case("hdl64_deskew", pk, t, synth.calib_hdl64(), synth.ins_trajectory(24))
    pk, t = synth.hdl32_packets(260, az0=30000.0, zero_frac=0.6)
    run_case("hdl32_nopose", pk, t, synth.calib_hdl32(), None)
    pk, t = synth.random_packets(120, 77)
    run_case("random_azimuth", pk, t, synth.calib_hdl64(), synth.ins_trajectory(12))
    pk, t = synth.hdl64_packets(400, az0=34000.0, seed=5, zero_frac=0.6)
    sel = np.ones(64, np.int32)
    sel[[2, 3, 50]] = 0
    run_case("hdl64_filters", pk, t, synth.calib_hdl64(), synth.ins_trajectory(24), sel=sel, skip=1,
             crop=(0, (-15.0, 15.0, -15.0, 15.0, -3.0, 3.0)))
    # pose interpolation table
    rng = np.random.default_rng(12)
    ts = synth.T0_US + np.cumsum(rng.integers(3_000, 20_000, 64)).astype(np.int64)
    trv = rng.normal(size=(64, 9)) * 30
    r = ref.RefParser()
    r.add_poses(ts, trv)
    q = np.concatenate([ts[::7], ts[:-1:5] + 1, [ts[0] - 9_000, ts[-1] + 40_000],
                        rng.integers(ts[0], ts[-1], 60)]).astype(np.int64)
    res = np.stack([r.interpolate(int(x))[1] for x in q])
    mats = np.stack([ref.RefParser.pose_matrix(v) for v in res])
    np.savez_compressed(os.path.join(HERE, "interpolate.npz"), pose_t=ts, pose_trv=trv, query=q,
                        result=res, matrix=mats)
    print("interpolate", len(q), "queries")


if __name__ == "__main__":
    main()
