"""Run only the C++ facade leg of bench.py (development aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from veloslam_b200 import synth
class A: pass
a = A(); a.facade_packets = 1 << 18; a.facade_batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
pk, t = synth.hdl64_stream_tiled(a.facade_packets)
out = bench.run_facade(a, 0, synth.calib_hdl64(), None, synth.as_bytes(pk), t, 1, torch.device("cuda", 0))
print(json.dumps(out, indent=1))
