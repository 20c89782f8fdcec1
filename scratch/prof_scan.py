"""Development aid: cycles per phase of k_scan (temporary clock64 instrumentation build)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from veloslam_b200 import capi, synth
n = 1 << 20
pk, t = synth.hdl64_stream_tiled(n, first_packet=0)
b = synth.as_bytes(pk)
ctx = capi.Context(0, max_batch_packets=n, max_poses=64, n_slots=1)
ctx.set_calibration(synth.calib_hdl64())
dev = torch.device("cuda", 0)
d_pk = torch.from_numpy(b).to(dev); d_t = torch.from_numpy(np.ascontiguousarray(t)).to(dev)
L = capi.load_library(); out = (C.c_ulonglong * 16)()
def run(k):
    for _ in range(k):
        ctx.wait(ctx.submit(d_pk, d_t, n=n, stride=1206, n_halo=0, mode=capi.MODE_STREAMING,
                            flags=capi.FLAG_DEVICE_INPUT, t_base_us=int(synth.T0_US)), frames=False)
run(2); L.vs_debug_fused_prof(out, 1); run(4); L.vs_debug_fused_prof(out, 1)
tiles = 4 * (n // 32)
names = ["wait tma", "A (w0) / B1 (w1)", "mid barrier", "B2 (+copy issue w0)", "epilogue (w0)", "end barrier"]
for w in range(2):
    tot = sum(out[w * 6 + i] for i in range(6))
    print("warp", w, "cycles/tile %.0f" % (tot / tiles))
    for i, nm in enumerate(names):
        print("   %-22s %7.0f" % (nm, out[w * 6 + i] / tiles))
ctx.close()
