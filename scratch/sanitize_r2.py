"""Round-2 workloads for compute-sanitizer: the default pipeline, the per-point deskew variant and
the HDLFrame layout kernels (k_layout_rows, k_layout) incl. carried rows, VLP-16 and a stream
with several frame starts per tile."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from veloslam_b200 import capi, synth
import parity as P

calib = synth.calib_hdl64()
poses = synth.ins_trajectory(30)
pk, t = synth.hdl64_packets(700, az0=35000.0)
b = synth.as_bytes(pk)
ctx = P.make_ctx(calib, poses, max_batch_packets=1024)
o = P.make_oracle(calib, poses)
o.process_packets(b, t)
frames = P.gpu_layout_stream(ctx, b, t, splits=(233, 234, 500))
P.assert_layout_parity(o, frames, P.TOL_DESKEW)
r = ctx.wait(ctx.submit(b, t, flags=capi.FLAG_DESKEW_PER_POINT, t_base_us=int(t[0])))
ctx.close()
# many frame starts per tile, skipped blocks
rng = np.random.default_rng(5)
pk, t = synth.hdl64_packets(150)
pk = pk.copy()
pk["blocks"]["azimuth"] = rng.integers(0, 36000, size=pk["blocks"]["azimuth"].shape)
b = synth.as_bytes(pk)
ctx = P.make_ctx(calib, poses, max_batch_packets=1024)
o = P.make_oracle(calib, poses)
o.process_packets(b, t)
P.assert_layout_parity(o, P.gpu_layout_stream(ctx, b, t, splits=(77,)), P.TOL_DESKEW)
ctx.close()
# VLP-16
c = synth.calib_hdl32()
c.n_enabled = 16
pk, t = synth.hdl32_packets(300, az0=10.0)
b = synth.as_bytes(pk)
ctx = P.make_ctx(c, max_batch_packets=1024)
o = P.make_oracle(c)
o.process_packets(b, t)
P.assert_layout_parity(o, P.gpu_layout_stream(ctx, b, t, splits=(111,)), P.TOL_DECODE)
ctx.close()
print("ok", len(frames))
