"""Development aid: run the bench batch through a -DVS_PROFILE_FUSED build of the library and print
the cycles per wait site of the single-pass kernel (see VS_PROF_ADD in vs_kernels.cuh)."""
import ctypes as C
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from veloslam_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = 5
pk, t = synth.hdl64_stream_tiled(n, first_packet=0)
b = synth.as_bytes(pk)
calib = synth.calib_hdl64()
poses = synth.ins_trajectory(int((n + 64) * 288e-6 * 100) + 40)
ctx = capi.Context(0, max_batch_packets=n, max_poses=len(poses[0]) + 8, n_slots=2)
ctx.set_calibration(calib)
ctx.set_poses(poses[0], poses[1])
dev = torch.device("cuda", 0)
d_pk = torch.from_numpy(b).to(dev)
d_t = torch.from_numpy(np.ascontiguousarray(t)).to(dev)
torch.cuda.synchronize()
L = capi.load_library()
out = (C.c_ulonglong * 16)()
def run(k):
    ms = []
    for _ in range(k):
        r = ctx.wait(ctx.submit(d_pk, d_t, n=n, stride=1206, n_halo=0, mode=capi.MODE_STREAMING,
                                flags=capi.FLAG_DEVICE_INPUT, t_base_us=int(synth.T0_US)), frames=False)
        ms.append(r.decode_ms)
    return ms
run(3)
if hasattr(L, "vs_debug_fused_prof"):
    L.vs_debug_fused_prof(out, 1)
ms = run(steps)
print("decode_ms", ["%.3f" % m for m in ms])
if hasattr(L, "vs_debug_fused_prof"):
    L.vs_debug_fused_prof(out, 1)
    names = ["scan:wait full", "scan:wait masks", "scan:lookback resolve", "scan:issue wait empty",
             "scan:issue grab+tma", "scan:counts+publish+recs", "dec:wait ready", "dec:mask pass",
             "dec:tile body", "scan:phase A", "scan:post-lookback", "-"]
    tiles = steps * ((n + 7) // 8)
    for i, nm in enumerate(names):
        v = out[i]
        # scan sites: one warp per tile; decode sites: 8 warps per tile
        per = v / tiles / (8 if nm.startswith("dec") else 1)
        print("%-28s %14d cycles  %8.0f cycles/tile/warp" % (nm, v, per))
ctx.close()
