"""Small decode workloads for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from veloslam_b200 import capi, synth
import parity as P

calib = synth.calib_hdl64()
poses = synth.ins_trajectory(30)
for n, zf in ((700, 0.05), (37, 0.9)):
    pk, t = synth.hdl64_packets(n, zero_frac=zf, az0=35000.0)
    b = synth.as_bytes(pk)
    ctx = P.make_ctx(calib, poses, max_batch_packets=1024)
    batches = P.gpu_stream(ctx, b, t, splits=(n // 3,))
    o = P.make_oracle(calib, poses)
    o.trace_enable()
    o.process_packets(b, t)
    P.assert_stream_parity(o, batches, P.TOL_DESKEW, t, calib=calib)
    ctx.close()
pk, t = synth.hdl32_packets(300)
ctx = P.make_ctx(synth.calib_hdl32(), max_batch_packets=1024)
r = ctx.decode(synth.as_bytes(pk), t)
print("ok", r.n_points)
