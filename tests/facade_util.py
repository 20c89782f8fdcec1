"""Helpers to run tests/cpp/facade_driver (the C++ facade) and read back what it dumps."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# VS_TEST_DRIVER: another build of the driver (e.g. a ThreadSanitizer one) for the same tests
DRIVER = os.environ.get("VS_TEST_DRIVER") or os.path.join(ROOT, "tests", "cpp", "facade_driver")


def build():
    if not os.environ.get("VS_TEST_DRIVER"):
        from veloslam_b200.build import build_facade
        build_facade()
    return DRIVER


def write_poses(path, pose_t, trv):
    rec = np.zeros(len(pose_t), dtype=[("t", "<i8"), ("v", "<f8", (9,))])
    rec["t"] = pose_t
    rec["v"] = np.asarray(trv).reshape(-1, 9)
    rec.tofile(path)


def run(args, timeout=600):
    return subprocess.run([DRIVER] + [str(a) for a in args], capture_output=True, text=True,
                          timeout=timeout)


class DumpedFrame:
    pass


def read_frames(path):
    raw = open(path, "rb").read()
    (nf,) = struct.unpack_from("<i", raw, 0)
    off = 4
    frames = []
    for _ in range(nf):
        f = DumpedFrame()
        f.timestamp_us, f.skips, f.n_lasers, f.n_packets, valid = struct.unpack_from("<qiiii", raw, off)
        off += 24
        f.carpose_valid = bool(valid)
        f.carpose_TRV = np.frombuffer(raw, "<f8", 9, off).copy()
        off += 72
        f.laser_counts = np.frombuffer(raw, "<i4", f.n_lasers, off).copy()
        off += 4 * f.n_lasers
        n = int(f.laser_counts.sum())
        rec = np.frombuffer(raw, dtype=[("xyzi", "<f4", (4,)), ("az", "<u2"), ("dist", "<f4")],
                            count=n, offset=off)
        off += n * 22
        f.n_points = n
        f.xyzi = rec["xyzi"].copy()
        f.azimuth = rec["az"].copy()
        f.distance = rec["dist"].copy()
        frames.append(f)
    assert off == len(raw)
    return frames


def free_udp_port():
    import socket
    s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_with_udp_feed(args, port, packets_u8, pace_s=40e-6, timeout=60):
    """Start the driver (it prints "ready" once its socket is bound), send every row of
    packets_u8 as one UDP datagram to 127.0.0.1:port, return (returncode, stderr)."""
    import socket
    import time
    p = subprocess.Popen([DRIVER] + [str(a) for a in args], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True)
    line = p.stdout.readline()
    if line.strip() != "ready":
        p.kill()
        return 1, "driver did not come up: " + line + p.stderr.read()
    tx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    for row in packets_u8:
        tx.sendto(row.tobytes(), ("127.0.0.1", port))
        t_end = time.perf_counter() + pace_s
        while time.perf_counter() < t_end:
            pass
    tx.close()
    try:
        _, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        p.kill()
        return 1, "timeout"
    return p.returncode, err
