"""Run the paced UDP online leg of bench.py a few times and print each stream's line (+ the
driver's stderr: VELOSLAM_TRACE_SLOW_MS traces)."""
import json, subprocess, sys, types
sys.path.insert(0, ".")
import torch
import bench
from veloslam_b200 import synth
_run = subprocess.run
def run(cmd, **kw):
    p = _run(cmd, **kw)
    if getattr(p, "stderr", None):
        print("stderr:", p.stderr[-3000:])
    return p
bench.subprocess.run = run
args = types.SimpleNamespace(online_udp_seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 12.0)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    out = bench.run_online_udp(args, 0, 0, synth.calib_hdl64(), 1, torch.device("cuda:0"))
    print(json.dumps(out["per_stream"][0]))
