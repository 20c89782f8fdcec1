"""HDL-32E (BASELINE configs[0] sensor, ADJ=1 kernels: per-return azimuth/time adjustment, LUT
branch) device-resident decode rate on one GPU: 262 144 packets, no poses."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from veloslam_b200 import capi, synth
n = 1 << 18
pk, t = synth.hdl32_packets(n)
b = synth.as_bytes(pk)
ctx = capi.Context(0, max_batch_packets=n, max_poses=8, n_slots=2)
ctx.set_calibration(synth.calib_hdl32())
dev = torch.device("cuda", 0)
d_pk = torch.from_numpy(b).to(dev); d_t = torch.from_numpy(np.ascontiguousarray(t)).to(dev)
def sub():
    return ctx.submit(d_pk, d_t, n=n, stride=1206, n_halo=0, mode=capi.MODE_STREAMING,
                      flags=capi.FLAG_DEVICE_INPUT, t_base_us=int(t[0]))
for _ in range(3):
    r = ctx.wait(sub(), frames=False)
torch.cuda.synchronize(); t0 = time.perf_counter(); K = 20
pend = sub(); dec = []
for i in range(K):
    nxt = sub() if i + 1 < K else None
    r = ctx.wait(pend, frames=False); dec.append(r.decode_ms); pend = nxt
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
print("HDL-32E: %d points/batch, %.3f ms/step, %.1f G points/s, k_decode<1> %.3f ms" %
      (r.n_points, dt * 1e3, r.n_points / dt / 1e9, np.mean(dec)))
ctx.close()
