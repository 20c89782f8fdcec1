import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from veloslam_b200 import calibxml, synth
import facade_util as F, parity as P
F.build()
pk, t = synth.hdl64_packets(1500)
calib = synth.calib_hdl64(); poses = synth.ins_trajectory(80)
b = synth.as_bytes(pk)
d = "/tmp/gd"; os.makedirs(d, exist_ok=True)
b.tofile(d + "/pk.bin"); t.astype("<i8").tofile(d + "/t.bin"); F.write_poses(d + "/poses.bin", *poses)
calibxml.write_db_xml(d + "/db.xml", calib)
o = P.make_oracle(calib, poses); o.process_packets(b, t)
print("oracle frames", [f.n_points for f in o.frames()])
for env in ({"VELOSLAM_GRAPH": "0"}, {}):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([F.DRIVER, "stream", d + "/db.xml", d + "/pk.bin", d + "/t.bin", d + "/poses.bin", d + "/out.bin", "4096", "0"], capture_output=True, text=True, env=e)
    fr = F.read_frames(d + "/out.bin")
    print(env, r.returncode, r.stderr[-300:], [f.n_points for f in fr], [f.timestamp_us - int(t[0]) for f in fr])
