"""Where the per-rotation latency goes (host wall clock): submit vs wait vs device time."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from veloslam_b200 import capi, synth
rot = 347
n_rot = 300
pk, t = synth.hdl64_stream_tiled(rot * n_rot)
b = synth.as_bytes(pk)
poses = synth.ins_trajectory(int(rot * n_rot * 288e-6 * 100) + 40)
ctx = capi.Context(0, max_batch_packets=512, max_poses=len(poses[0]) + 8, n_slots=1)
ctx.set_calibration(synth.calib_hdl64())
ctx.set_poses(*poses)
h_pk = torch.from_numpy(b).pin_memory()
h_t = torch.from_numpy(np.ascontiguousarray(t)).pin_memory()
carry = capi.carry_init()
ts, tw, tg = [], [], []
for r in range(n_rot):
    a = r * rot
    t0 = time.perf_counter()
    tk = ctx.submit(h_pk[a:a + rot], h_t[a:a + rot], n=rot, stride=1206, flags=0, t_base_us=int(t[0]), carry=carry)
    t1 = time.perf_counter()
    res = ctx.wait(tk, frames=False)
    t2 = time.perf_counter()
    carry = res.carry_out
    ts.append(t1 - t0); tw.append(t2 - t1); tg.append(res.gpu_ms)
ts, tw, tg = [np.array(x[20:]) for x in (ts, tw, tg)]
print("graph launches of the slot:", res.graph_launches)
print("submit p50 %.1f us  wait p50 %.1f us  device(kernels) p50 %.1f us  total p50 %.1f us" %
      (np.median(ts) * 1e6, np.median(tw) * 1e6, np.median(tg) * 1e3, np.median(ts + tw) * 1e6))
