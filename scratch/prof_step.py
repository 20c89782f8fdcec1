"""Profiling target: a few device-resident steps of the bench batch (k_scan, k_pose, k_decode,
k_frames on 1 Mi packets) followed by the HDLFrame layout of the same batch (k_layout_rows,
k_layout).  Run under ncu by scratch/run_r2_prof.sh."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from veloslam_b200 import capi, synth
n = int(os.environ.get("PROF_PACKETS", 1 << 20))
steps = int(os.environ.get("PROF_STEPS", 4))
dev = torch.device("cuda", 0)
pk, t = synth.hdl64_stream_tiled(n)
poses = synth.ins_trajectory(int(n * 288e-6 * 100) + 40)
ctx = capi.Context(0, max_batch_packets=n, max_poses=len(poses[0]) + 8, n_slots=1)
ctx.set_calibration(synth.calib_hdl64())
ctx.set_poses(*poses)
d_pk = torch.from_numpy(synth.as_bytes(pk)).to(dev)
d_t = torch.from_numpy(np.ascontiguousarray(t)).to(dev)
for i in range(steps):
    r = ctx.wait(ctx.submit(d_pk, d_t, n=n, stride=1206, flags=capi.FLAG_DEVICE_INPUT,
                            t_base_us=int(t[0])), frames=False)
lay, rows = ctx.layout_frames(r.ticket, None, 16, True)
ms = ctx.sync(r.ticket)
print("points", r.n_points, "gpu_ms", r.gpu_ms, "decode_ms", r.decode_ms, "layout_ms", ms)
ctx.close()
