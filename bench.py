#!/usr/bin/env python
"""bench.py -- decoded+deskewed points/sec of the VeloSLAM ingest hot path on B200.

A "step" is one pass of the hot path (segmentation + per-packet pose + decode/calibrate/
transform/compaction + frame table) over one batch of synthetic HDL-64E S2 packets with a
100 Hz INS timeline (BASELINE.json configs[2], the decode+deskew configuration the metric is
quoted on).  `value` is measured with the packets already resident in HBM; `e2e` goes through
the same C ABI with HOST buffers (pinned), H2D of packets/times and D2H of every point column
inside the timed region.  N > 1: one process per GPU (torchrun), contiguous packet-range shards
of one long recording with a one-rotation halo, no data-path collective (weak scaling); the
frame-index all-gather runs once after the timed loop.

`--impl reference` times the reference's CPU algorithm (oracle/_ref when the reference compiled
here, else the oracle port) on the host cores with all threads it can use.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoded+deskewed points/sec"
UNIT = "points/s"
BYTES_PER_POINT_OUT = 22          # x,y,z f32 + intensity u8 + laser u8 + azimuth u16 + distance u16 + t u32
# What the reference-facing facade (veloslam_b200/cpp/HDLParser.cpp, decodePending) copies back
# to build HDLFrame::points / pointsMeta: every column but t_us (HDLFrame has no per-point time).
BYTES_PER_POINT_E2E = 18
HBM_FALLBACK_GBS = 6650.0         # /opt/skills/guides/B200_PROFILING.md fallback
WORKLOAD = ("HDL-64E S2 decode + rotation segmentation + per-packet deskew against a 100 Hz INS "
            "timeline (BASELINE.json configs[2]; N>1: configs[3] packet-range shards with a "
            "512-packet halo)")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--packets", type=int, default=1 << 20, help="packets per GPU per step")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 16, help="packets per host->device chunk")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-online", action="store_true")
    ap.add_argument("--no-deskew", action="store_true")
    ap.add_argument("--no-single-pass", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--recording-hours", type=float, default=0.0,
                    help="headline = BASELINE.json configs[3]: ONE pass over a recording of this "
                         "many hours of HDL-64E (12 499 200 packets per hour), split by packet range "
                         "over the N ranks (strong scaling), frame-table all-gather + stitch timed")
    ap.add_argument("--recording-leg-hours", type=float, default=1.0,
                    help="size of the configs[3] sub-measurement of a default run (0: skip)")
    ap.add_argument("--no-facade", action="store_true")
    ap.add_argument("--no-hdl32", action="store_true")
    ap.add_argument("--facade-packets", type=int, default=1 << 18,
                    help="packets per pass of the C++ facade run")
    ap.add_argument("--facade-batch", type=int, default=1 << 16)
    ap.add_argument("--online-udp-seconds", type=float, default=60.0,
                    help="length of the paced 10 Hz UDP stream per GPU (configs[4]); 0: skip")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread
    (2 ms period; the timed region is tens of milliseconds), nvidia-smi as a fallback."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, cuda_index):
        self.cuda_index = cuda_index
        self.handle = None
        self.nvml = None
        self.thread = None
        self.stop_flag = False
        self.sm = []
        self.mask = 0
        self.proc = None
        self.path = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            for cand in (uuid, "GPU-" + uuid):
                try:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
                    break
                except Exception:
                    self.handle = None
            if self.handle is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[cuda_index]) if vis else cuda_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.handle = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    self.mask |= n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    self.mask |= n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.handle is not None:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.cuda_index), "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            n = self.nvml
            if self.sm:
                out["sm_mhz"] = float(np.median(self.sm))
                out["samples"] = len(self.sm)
            try:
                out["sm_max_mhz"] = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                pass
            out["reasons"] = sorted(v for k, v in self.REASONS.items() if self.mask & k)
            out["source"] = "nvml"
            return out
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = smax
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        out["source"] = "nvidia-smi"
        return out


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------
def load_cpu_reference():
    """(kind, factory): oracle/_ref (the reference's own sources) when built, else the port."""
    try:
        from oracle import ref as ref_mod
        if ref_mod.available():
            return "reference", ref_mod.RefParser
    except Exception:
        pass
    from oracle.oracle import Oracle
    return "port", Oracle


def cpu_make_parser(factory, calib, poses):
    o = factory()
    o.set_calibration(calib)
    o.add_poses(poses[0], poses[1])
    return o


def cpu_time_threads(factory, calib, poses, pk_bytes, t_us, n_threads):
    """Packet-range shards over n_threads host threads (each its own parser, like one
    reference process per shard); returns (seconds, slots)."""
    n = pk_bytes.shape[0]
    cuts = [(n * i) // n_threads for i in range(n_threads + 1)]
    parsers = [cpu_make_parser(factory, calib, poses) for _ in range(n_threads)]
    threads = []

    def work(i):
        a, b = cuts[i], cuts[i + 1]
        if b > a:
            # the reference's consumer loop (HDLSource.cxx:209-225): frames are taken and the
            # parser's list cleared as they close, after every packet
            parsers[i].consume_packets(pk_bytes[a:b], t_us[a:b])

    t0 = time.perf_counter()
    for i in range(n_threads):
        th = threading.Thread(target=work, args=(i,))
        th.start()
        threads.append(th)
    for th in threads:
        th.join()
    dt = time.perf_counter() - t0
    return dt, n * 384


def count_points(pk_struct):
    return int(np.count_nonzero(pk_struct["blocks"]["returns"]["distance"]))


def run_cpu_baseline(calib, poses, budget_s):
    """Single-thread oracle on a bounded sample of the same workload."""
    from veloslam_b200 import synth
    kind, factory = load_cpu_reference()
    probe_n = 8192
    pk, t = synth.hdl64_stream_tiled(probe_n)
    dt, _ = cpu_time_threads(factory, calib, poses, synth.as_bytes(pk), t, 1)
    rate = probe_n / dt
    n = int(max(probe_n, min(rate * budget_s, 1 << 19)))
    pk, t = synth.hdl64_stream_tiled(n)
    dt, slots = cpu_time_threads(factory, calib, poses, synth.as_bytes(pk), t, 1)
    pts = count_points(pk)
    return {"value": pts / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{n} HDL-64E packets ({slots} return slots, {pts} points emitted) of the "
                      f"bench workload, single thread, {dt:.2f} s",
            "slots_per_s": slots / dt}


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from veloslam_b200 import synth
    kind, factory = load_cpu_reference()
    calib = synth.calib_hdl64()
    n_cpus = os.cpu_count() or 1
    # The reference's per-point push_back / per-packet string copies are allocator- and
    # memory-bound: on boxes with many (hyper-)threads its throughput FALLS beyond some thread
    # count (round 1: 32 threads slower than 16).  The arm therefore probes a few thread counts on
    # a short prefix and runs with the best one -- the strongest CPU figure this box gives.
    # One step = the GPU arm's own batch (args.packets packets of the same synthetic stream) on
    # those threads, unless this box is too slow to finish steps + warmup of that within
    # ~4 minutes: then a shorter prefix of the same stream.
    probe_n = 2048 * n_cpus
    pk, t = synth.hdl64_stream_tiled(probe_n)
    poses = synth.ins_trajectory(int(probe_n * 288e-6 * 100) + 40)
    probe = {}
    cpu_time_threads(factory, calib, poses, synth.as_bytes(pk), t, n_cpus)   # page in, warm the allocator
    for cand in sorted({n_cpus, max(1, n_cpus // 2), max(1, (3 * n_cpus) // 4), min(n_cpus, 16)}):
        cpu_time_threads(factory, calib, poses, synth.as_bytes(pk), t, cand)
        dt, _ = cpu_time_threads(factory, calib, poses, synth.as_bytes(pk), t, cand)
        probe[cand] = probe_n / dt
    n_threads = max(probe, key=probe.get)
    rate = probe[n_threads]
    total_steps = args.steps + args.warmup
    n = int(max(probe_n, min(args.packets, rate * 240.0 / total_steps)))
    pk, t = synth.hdl64_stream_tiled(n)
    poses = synth.ins_trajectory(int(n * 288e-6 * 100) + 40)
    b = synth.as_bytes(pk)
    pts = count_points(pk)
    # the reference's own threading model: ONE consumer thread per stream (HDLSource.cxx:227-235)
    n1 = min(n, 1 << 16)
    dt1, _ = cpu_time_threads(factory, calib, poses, b[:n1], t[:n1], 1)
    one_thread = count_points(pk[:n1]) / dt1
    for _ in range(args.warmup):
        cpu_time_threads(factory, calib, poses, b, t, n_threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_time_threads(factory, calib, poses, b, t, n_threads)
    dt = time.perf_counter() - t0
    value = pts * args.steps / dt
    sample = (f"{n} HDL-64E packets per step ({n * 384} slots, {pts} points; the GPU arm's batch is "
              f"{args.packets} packets), {n_threads} threads over packet-range shards, each running "
              f"the reference's consumer loop (processHDLPacket + getAllFrames/clearAllFrames per packet)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "packets_per_step": n, "same_batch_as_gpu_arm": n == args.packets,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": kind,
                         "sample": sample, "one_thread_per_stream_value": one_thread,
                         "host_cpus": n_cpus,
                         "thread_probe_packets_per_s": {str(k): v for k, v in sorted(probe.items())}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: one pass over a long recording, packet-range sharded (strong scaling)
# ------------------------------------------------------------------------------------------
PACKETS_PER_HOUR = 12_499_200     # 1 h of HDL-64E at 3472 packets/s


def device_stream(first, n, dev, base_packets=16384):
    """Packets [first, first + n) of the infinite synthetic HDL-64E stream of
    synth.hdl64_stream_tiled(), built ON THE DEVICE (a 1-h recording is 15 GB): the returns repeat
    a seeded base block, azimuths and times are those of the global packet index."""
    import torch
    from veloslam_b200 import synth
    base, _ = synth.hdl64_packets(base_packets)
    d_base = torch.from_numpy(synth.as_bytes(base)).to(dev)
    idx = torch.arange(first, first + n, dtype=torch.int64, device=dev)
    d_pk = d_base.index_select(0, idx % base_packets)
    for pair in range(6):
        az = torch.floor(12345.0 + (idx * 6 + pair).to(torch.float64) * synth.HDL64_TICKS_PER_PAIR)
        az = az.to(torch.int64) % 36000
        lo, hi = (az & 0xff).to(torch.uint8), (az >> 8).to(torch.uint8)
        for j in (2 * pair, 2 * pair + 1):
            d_pk[:, 100 * j + 2] = lo
            d_pk[:, 100 * j + 3] = hi
    d_t = synth.T0_US + idx * synth.HDL64_US_PER_PACKET
    del idx, d_base
    return d_pk, d_t


def run_recording(hours, steps, warmup, local, rank, world, dev, calib):
    """ONE pass per step over a `hours`-long recording: rank g decodes its packet range (+ halo),
    then the ranks all-gather their frame tables (NCCL) and every rank stitches the global frame
    index -- all inside the timed region.  Returns a dict (rank 0's view, times = max over ranks)."""
    import torch
    from veloslam_b200 import capi, sharding, synth
    n_total = int(round(PACKETS_PER_HOUR * hours))
    first, halo, end = capi.shard_range(n_total, world, rank, sharding.HALO_HDL64)
    n_sub = end - first + halo
    t_base = int(synth.T0_US)
    poses = synth.ins_trajectory(int(n_total * 288e-6 * 100) + 60)
    ctx = capi.Context(local, max_batch_packets=n_sub, max_poses=len(poses[0]) + 8, n_slots=1)
    ctx.set_calibration(calib)
    ctx.set_poses(poses[0], poses[1])
    d_pk, d_t = device_stream(first - halo, n_sub, dev)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(ctx.stream(0), device=dev)
    # fixed-size exchange buffer: row 0 = [n_rows, 0, ...], then the rows (one collective, no
    # count exchange); a 10 Hz stream closes one frame per ~347 packets
    cap_rows = (n_total + world - 1) // world // 300 + 64
    h_mine = torch.zeros((cap_rows + 1, capi.FRAME_ROW_COLS), dtype=torch.int64).pin_memory()
    h_mine_np = h_mine.numpy()
    d_mine = torch.zeros_like(h_mine, device=dev)
    d_all = torch.zeros((world * (cap_rows + 1), capi.FRAME_ROW_COLS), dtype=torch.int64, device=dev)
    h_all = torch.zeros_like(d_all, device="cpu").pin_memory()
    h_all_np = h_all.numpy().reshape(world, cap_rows + 1, capi.FRAME_ROW_COLS)
    stitcher = capi.Stitcher(world, world * cap_rows)
    ag0 = torch.cuda.Event(enable_timing=True)
    ag1 = torch.cuda.Event(enable_timing=True)

    waits = []

    def one_pass():
        # the frame table never visits the host on its way to the other ranks: k_table_rows writes
        # the exchange rows behind the batch's kernels, the all-gather reads them out of HBM
        tk = ctx.submit(d_pk, d_t, n=n_sub, stride=1206, n_halo=halo, mode=capi.MODE_STREAMING,
                        flags=capi.FLAG_DEVICE_INPUT | capi.FLAG_NO_FRAME_LIST, t_base_us=t_base)
        ctx.frame_table_rows_device(tk, rank, first, d_mine, cap_rows)
        cur = torch.cuda.current_stream()
        cur.wait_stream(ext)
        if world > 1:
            ag0.record()
            torch.distributed.all_gather_into_tensor(d_all, d_mine)
            ag1.record()
            h_all.copy_(d_all, non_blocking=True)
        else:
            h_mine.copy_(d_mine, non_blocking=True)
        wq = time.perf_counter()
        r = ctx.wait(tk, frames=False)          # batch errors (halo, capacity, time range) surface here
        w0 = time.perf_counter()
        waits.append((w0 - wq) * 1e3)
        cur.synchronize()
        w1 = time.perf_counter()
        if world > 1:
            ag_ms = ag0.elapsed_time(ag1)
            counts = h_all_np[:, 0, 0]
            if (counts < 0).any():
                raise SystemExit("bench.py: a rank's frame table exceeds the exchange buffer")
            # table g sits at row g * (cap_rows + 1) + 1 of the gathered buffer: stitched in place
            gf, segs = stitcher(h_all_np, counts, cap_rows + 1, first_row=1)
        else:
            ag_ms = 0.0
            n_rows = int(h_mine_np[0, 0])
            if n_rows < 0:
                raise SystemExit("bench.py: the frame table exceeds the exchange buffer")
            gf, segs = stitcher(h_mine_np, [n_rows], 0, first_row=1)
        w2 = time.perf_counter()
        return r, gf, segs, ag_ms, (w1 - w0) * 1e3, (w2 - w1) * 1e3

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 1)):
        r, gf, segs, _, _, _ = one_pass()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    w0 = time.perf_counter()
    dec, ags, exch, sti = [], [], [], []
    for _ in range(steps):
        r, gf, segs, ag_ms, ex_ms, st_ms = one_pass()
        dec.append(r.gpu_ms)
        ags.append(ag_ms)
        exch.append(ex_ms)
        sti.append(st_ms)
    # the last device op of a pass is the gathered table's copy on torch's stream
    ej = torch.cuda.Event()
    ej.record()
    ext.wait_event(ej)
    e1.record(ext)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3 / steps
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    pts_local = r.n_points

    # the pass that was just timed against the CPU reference: a random run of rotations and the
    # last ones of this rank's range (point offsets beyond 2^32 and frame ids in the thousands
    # at N = 1)
    class _DeviceRows:
        def __getitem__(self, sl):
            return d_pk[sl].cpu().numpy()
    # (the timed passes assemble no frame list on the host: one more pass with it, whose rows
    # must be the device-built rows the timed pass exchanged)
    mine_rows = (h_all_np[rank] if world > 1 else h_mine_np).copy()
    r = ctx.wait(ctx.submit(d_pk, d_t, n=n_sub, stride=1206, n_halo=halo, mode=capi.MODE_STREAMING,
                            flags=capi.FLAG_DEVICE_INPUT, t_base_us=t_base))
    host_rows = capi.frame_table_rows(r.frame_table, rank, first, halo)
    if mine_rows[0, 0] != host_rows.shape[0] or not np.array_equal(mine_rows[1:1 + host_rows.shape[0]], host_rows):
        raise SystemExit("bench.py: the device-built frame-table rows differ from vs_frame_table_rows")
    parity = parity_windows(ctx, r, _DeviceRows(), d_t.cpu().numpy(), t_base, halo, calib, poses,
                            seed=99 + rank)
    if world > 1:
        allp = [None] * world
        torch.distributed.all_gather_object(allp, parity)
        bad = [p_ for p_ in allp if p_["status"] == "FAILED"]
        parity = bad[0] if bad else {"status": "ok" if all(p_["status"] == "ok" for p_ in allp) else allp[0]["status"],
                                     "max_abs_dxyz_m": max(p_.get("max_abs_dxyz_m", 0.0) for p_ in allp),
                                     "points": sum(p_.get("points", 0) for p_ in allp), "ranks": world}
    if parity["status"] == "FAILED":
        raise SystemExit("bench.py: the timed recording pass does NOT match the CPU reference: " + parity["why"])
    if world > 1:
        tt = torch.tensor([ms, wall_ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms, wall_ms = float(tt[0].item()), float(tt[1].item())
        pp = torch.tensor([pts_local], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(pp)
        pts = int(pp.item())
    else:
        pts = pts_local
    out = {"hours": hours, "packets": n_total, "points": pts, "ms_per_pass": ms,
           "wall_ms_per_pass": wall_ms, "points_per_s": pts / (ms * 1e-3),
           "global_frames": int(len(gf)), "frame_segments": int(len(segs)),
           "stitch_timestamp_mismatches": int(gf["timestamp_mismatch"].sum()),
           "points_in_index": int(gf["n_points"].sum()),
           "decode_ms_this_rank": float(np.mean(dec)), "nccl_allgather_ms": float(np.mean(ags)),
           "exchange_ms_host": float(np.mean(exch)), "stitch_ms": float(np.mean(sti)),
           "vs_wait_ms_host": float(np.mean(waits[-steps:])),
           "exchange_bytes_per_rank": int((cap_rows + 1) * capi.FRAME_ROW_COLS * 8),
           "packets_this_rank": int(end - first), "halo": int(halo),
           "points_this_rank": int(pts_local), "launches_per_pass": int(r.n_kernel_launches) + 1,   # + k_table_rows
           "parity_window": parity["status"], "parity": parity, "clocks": clocks,
           "note": "timed region of a pass: vs_submit + vs_wait of the rank's packet range (k_scan, "
                   "k_pose, k_decode, k_frames; recording resident in HBM, generated on the device), "
                   "frame-table rows built on the device (vs_frame_table_rows_device, k_table_rows), ONE "
                   "NCCL all_gather_into_tensor of fixed-size tables out of HBM, D2H, "
                   "vs_stitch_frame_tables on every rank; CUDA events, max over ranks; "
                   "exchange_ms_host = host wait for the gathered rows after vs_wait returned"}
    if out["points_in_index"] != pts:
        raise SystemExit("bench.py: the stitched frame index does not cover every decoded point")
    ctx.close()
    del d_pk, d_t
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------
# parity of what is timed: windows of the timed batch against the CPU reference
# ------------------------------------------------------------------------------------------
def parity_windows(ctx, res, b, t, t_base, halo, calib, poses, seed, rotations=12):
    """Compare two windows of the batch that was just timed -- a random run of `rotations`
    closed frames (~4096 packets) and the last closed frames of the batch -- with the CPU
    reference (oracle/_ref when built, else the oracle port), point for point: laser,
    intensity, azimuth, raw distance and the time column bit-exact, coordinates within the 1e-3 m
    bar of BASELINE.json.  The reference parser is started 600 packets (> one rotation) early
    so that lastAzimuth / firingSkip / the frame origin are those of the whole stream.
    `b`, `t`: the host copy of the submitted packets (halo included)."""
    from oracle.oracle import Oracle
    tab = res.frame_table
    n_closed = res.n_closed
    if n_closed < rotations + 3:
        return {"status": "skipped", "why": "batch holds too few rotations"}
    rng = np.random.default_rng(seed)
    starts = [int(rng.integers(2, n_closed - rotations)), n_closed - rotations]
    out = {"status": "ok", "windows": [], "max_abs_dxyz_m": 0.0, "points": 0, "checker": "oracle port"}
    for i0 in starts:
        i1 = i0 + rotations
        p0, k0 = int(tab["start_packet"][i0]), int(tab["start_block"][i0])
        p1, k1 = int(tab["start_packet"][i1]), int(tab["start_block"][i1])
        fp0, fp1 = int(tab["first_point"][i0]), int(tab["first_point"][i1])
        a = max(0, p0 - 600)
        o = Oracle()
        o.set_calibration(calib)
        o.add_poses(poses[0], poses[1])
        o.trace_enable()
        o.process_packets(b[a:p1 + 1], t[a:p1 + 1])
        tr = o.trace()
        key = (tr["packet"].astype(np.int64) + a) * 12 + tr["block"]
        sel = (key >= p0 * 12 + k0) & (key < p1 * 12 + k1)
        got = res.fetch(fp0, fp1 - fp0)
        n = int(sel.sum())
        if n != fp1 - fp0:
            return {"status": "FAILED", "why": f"window at frame {i0}: {fp1 - fp0} points on the GPU, "
                                               f"{n} from the CPU reference"}
        for k in ("laser", "intensity", "azimuth", "distance"):
            if not np.array_equal(got[k], tr[k][sel]):
                return {"status": "FAILED", "why": f"window at frame {i0}: column {k} differs"}
        want_t = (t[tr["packet"][sel].astype(np.int64) + a] - t_base).astype(np.uint32) + tr["tadj_us"][sel]
        if not np.array_equal(got["t_us"], want_t):
            return {"status": "FAILED", "why": f"window at frame {i0}: column t_us differs"}
        worst = 0.0
        for k in ("x", "y", "z"):
            worst = max(worst, float(np.max(np.abs(got[k].astype(np.float64) - tr[k][sel]))))
        if worst > 1e-3:
            return {"status": "FAILED", "why": f"window at frame {i0}: |dxyz| = {worst} m > 1e-3 m"}
        out["max_abs_dxyz_m"] = max(out["max_abs_dxyz_m"], worst)
        out["points"] += n
        out["windows"].append({"first_frame": i0, "frames": rotations, "packets": [p0, p1], "points": n})
    return out


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
NCU_FULL_CSV = os.path.join(ROOT, "profiles", "r2a_k_decode_ncu_full.csv")
NCU_STEP_CSVS = [os.path.join(ROOT, "profiles", f"r2a_{k}_ncu_full.csv") for k in ("k_scan", "k_pose", "k_decode")]
NCU_FULL_PACKETS = 1 << 20  # packets per launch of the captured kernels


def _ncu_dram_bytes(path):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(path):
        parts = line.strip().split(",")
        if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(parts[2]) * scale[parts[1]]
    return tot


def ncu_traffic(n_per):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_decode launch, and of one whole step
    (k_scan + k_pose + k_decode), from the committed `ncu --set full` captures of the same
    workload and batch size; else None."""
    if n_per != NCU_FULL_PACKETS or not os.path.exists(NCU_FULL_CSV):
        return None, None, None
    try:
        tot = _ncu_dram_bytes(NCU_FULL_CSV)
        step = sum(_ncu_dram_bytes(p) for p in NCU_STEP_CSVS)
    except Exception:
        return None, None, None
    return (tot if tot > 0 else None), os.path.relpath(NCU_FULL_CSV, ROOT), (step if step > 0 else None)


def pin_to_gpu_numa_node(cuda_index):
    """Run this rank's host threads on the CPUs NVML reports as local to its GPU, so that the
    pinned packet ring and result columns are allocated on the memory node next to the GPU's PCIe
    root (host<->device copies then do not cross the socket interconnect).  Best effort."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
        handle = None
        for cand in (uuid, "GPU-" + uuid):
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
                break
            except Exception:
                handle = None
        if handle is None:
            return None
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    from veloslam_b200 import capi, sharding, synth

    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_cpus_all = len(os.sched_getaffinity(0))
    numa_cpus = pin_to_gpu_numa_node(local)

    if args.recording_hours > 0:
        run_recording_headline(args, local, rank, world, dev, host_cpus_all, numa_cpus)
        return

    n_per = args.packets
    halo = sharding.HALO_HDL64 if rank > 0 else 0
    first = rank * n_per
    pk, t = synth.hdl64_stream_tiled(n_per + halo, first_packet=first - halo)
    b = synth.as_bytes(pk)
    n_sub = n_per + halo
    t_base = int(synth.T0_US)
    calib = synth.calib_hdl64()
    span_s = (world * n_per + 64) * 288e-6
    poses = synth.ins_trajectory(int(span_s * 100) + 40)
    n_emitted_expected = None

    ctx = capi.Context(local, max_batch_packets=n_sub, max_poses=len(poses[0]) + 8, n_slots=2)
    ctx.set_calibration(calib)
    ctx.set_poses(poses[0], poses[1])
    d_pk = torch.from_numpy(b).to(dev)
    d_t = torch.from_numpy(np.ascontiguousarray(t)).to(dev)
    torch.cuda.synchronize()

    ext = torch.cuda.ExternalStream(ctx.stream(0), device=dev)
    ext1 = torch.cuda.ExternalStream(ctx.stream(1), device=dev)

    def submit():
        return ctx.submit(d_pk, d_t, n=n_sub, stride=1206, n_halo=halo, mode=capi.MODE_STREAMING,
                          flags=capi.FLAG_DEVICE_INPUT, t_base_us=t_base)

    def step():
        return ctx.wait(submit(), frames=False)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        r = step()
    n_emitted = r.n_points
    launches_per_step = r.n_kernel_launches

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    dec_ms, step_ms = [], []
    # both slot streams start after e0 and are joined before e1
    e0.record(ext)
    ext1.wait_event(e0)
    t_wall0 = time.perf_counter()
    # Steps are independent batches (each rebuilds its state from its own carry / halo), so the
    # host keeps one batch in flight while it finishes the previous one: two result slots.
    pending = submit()
    for i in range(args.steps):
        nxt = submit() if i + 1 < args.steps else None
        r = ctx.wait(pending, frames=False)
        dec_ms.append(r.decode_ms)
        step_ms.append(r.gpu_ms)
        pending = nxt
    ej = torch.cuda.Event()
    ej.record(ext1)
    ext.wait_event(ej)
    e1.record(ext)
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    elapsed_ms = e0.elapsed_time(e1)
    per_rank_ms = [elapsed_ms / args.steps]
    if world > 1:
        tt = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(tt) for _ in range(world)]
        torch.distributed.all_gather(allt, tt)
        per_rank_ms = [float(x.item()) / args.steps for x in allt]
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
        pts = torch.tensor([n_emitted], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(pts)
        total_emitted = int(pts.item())
    else:
        total_emitted = n_emitted
    ms_per_step = elapsed_ms / args.steps
    value = total_emitted / (ms_per_step * 1e-3)

    # ---- the batch that was just timed, against the CPU reference --------------------------------
    parity = None
    if not args.no_parity:
        parity = parity_windows(ctx, r, b, t, t_base, halo, calib, poses, seed=1234 + rank)
        if world > 1:
            allp = [None] * world
            torch.distributed.all_gather_object(allp, parity)
            bad = [p_ for p_ in allp if p_["status"] == "FAILED"]
            parity = bad[0] if bad else {
                "status": "ok" if all(p_["status"] == "ok" for p_ in allp) else allp[0]["status"],
                "max_abs_dxyz_m": max(p_.get("max_abs_dxyz_m", 0.0) for p_ in allp),
                "points": sum(p_.get("points", 0) for p_ in allp),
                "windows": [w for p_ in allp for w in p_.get("windows", [])][:4], "ranks": world}
        if parity["status"] == "FAILED":
            raise SystemExit("bench.py: the timed batch does NOT match the CPU reference: " + parity["why"])

    # ---- roofline of the dominant kernel (k_decode), this rank ------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = HBM_FALLBACK_GBS, "fallback"
    try:
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured"
    except Exception:
        pass
    alg_bytes = 1206 * n_per + BYTES_PER_POINT_OUT * n_emitted
    traffic, traffic_src, traffic_step = ncu_traffic(n_per)
    dec_avg_ms = float(np.mean(dec_ms))
    achieved = alg_bytes / (dec_avg_ms * 1e-3) / 1e9
    step_avg_ms = float(np.mean(step_ms))
    roofline = {"bound": "hbm", "kernel": "k_decode", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": dec_avg_ms,
                # every kernel, memset and gap of the step: algorithmic bytes / ms_per_step
                "frac_whole_step": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                # DRAM bytes of k_scan + k_pose + k_decode (ncu) over the algorithmic bytes: the
                # packets are read twice and the scan's records are written and re-read
                "traffic_whole_step": traffic_step,
                "traffic_whole_step_over_algorithmic": (traffic_step / alg_bytes) if traffic_step else None,
                "note": "k_decode timed with CUDA events on its launch stream while the other "
                        "result slot's k_scan/k_pose may overlap it (two batches in flight)"}

    # ---- per-point deskew extension, same batch, off the headline number -------
    deskew = None
    if not args.no_deskew:
        off = np.zeros((12, 32), dtype=np.uint16)
        for j in range(12):   # S2-like timing: block pairs 48 us apart, lasers 1.5 us apart
            off[j] = np.round((j // 2) * 48.0 + np.arange(32) * 1.5).astype(np.uint16)
        ctx.set_firing_offsets(off)

        def submit_dsk():
            return ctx.submit(d_pk, d_t, n=n_sub, stride=1206, n_halo=halo, mode=capi.MODE_STREAMING,
                              flags=capi.FLAG_DEVICE_INPUT | capi.FLAG_DESKEW_PER_POINT,
                              t_base_us=t_base)
        for _ in range(3):
            ctx.wait(submit_dsk(), frames=False)
        barrier()
        d0 = torch.cuda.Event(enable_timing=True)
        d1 = torch.cuda.Event(enable_timing=True)
        d0.record(ext)
        ext1.wait_event(d0)
        dk, kd = [], 8
        pend = submit_dsk()               # two batches in flight, like the headline loop
        for i in range(kd):
            nxt = submit_dsk() if i + 1 < kd else None
            dk.append(ctx.wait(pend, frames=False).decode_ms)
            pend = nxt
        dj = torch.cuda.Event()
        dj.record(ext1)
        ext.wait_event(dj)
        d1.record(ext)
        barrier()
        ms = d0.elapsed_time(d1) / kd
        deskew = {"points_per_s_per_gpu": n_emitted / (ms * 1e-3), "ms_per_step": ms,
                  "k_decode_ms": float(np.mean(dk)),
                  "note": "VS_FLAG_DESKEW_PER_POINT: pose per point (slerp weights linearised per "
                          "packet, quaternion rotation, translation lerp), re-based to the frame "
                          "origin; not a reference behaviour; two batches in flight like the headline"}
        ctx.set_calibration(calib)   # resets the firing table

    # ---- frame index exchange (off the timed loop) ------------------------------------------
    rr = step()
    tab = sharding.local_table(rr.frame_table, rank, first, halo)
    if world > 1:
        tables = sharding.all_gather_tables(tab)
    else:
        tables = [tab]
    frames = sharding.stitch(tables)
    n_global_frames = len(frames)
    mism = sum(1 for f in frames if "timestamp_mismatch" in f)

    # ---- e2e through the C ABI with host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, ctx_args=(local, calib, poses), b=b[halo:], t=t[halo:], t_base=t_base,
                      world=world, dev=dev)
    e2e_frames = e2e_xyzi = None
    if not args.no_e2e:
        e2e_xyzi = run_e2e(args, ctx_args=(local, calib, poses), b=b[halo:], t=t[halo:], t_base=t_base,
                           world=world, dev=dev, xyzi_only=True)
        e2e_frames = run_e2e(args, ctx_args=(local, calib, poses), b=b[halo:], t=t[halo:], t_base=t_base,
                             world=world, dev=dev, layout=True)
    pcie = None
    if not args.no_e2e:
        pcie = run_pcie_probe(world, dev)
    e2e_facade = None
    if not args.no_facade and not args.no_e2e:
        e2e_facade = run_facade(args, local, calib, poses, b[halo:], t[halo:], world, dev)
    hdl32 = None
    if not args.no_hdl32 and rank == 0:
        hdl32 = run_hdl32(local, dev)
    online = None
    if not args.no_online:
        online = run_online(local, calib, poses, b[halo:], t[halo:], t_base)
        if world > 1:
            allo = [None] * world
            torch.distributed.all_gather_object(allo, online)
            online = {"per_rank_p99_ms": [o["index_only"]["p99_ms"] for o in allo],
                      "per_rank_p50_ms": [o["index_only"]["p50_ms"] for o in allo],
                      "aggregate_points_per_s": sum(o["index_only"]["points_per_s"] for o in allo),
                      "with_points_p99_ms": [o["with_points"]["p99_ms"] for o in allo],
                      "streams": world, "note": allo[0]["note"]}
    online_udp = None
    if args.online_udp_seconds > 0 and not args.no_online:
        online_udp = run_online_udp(args, local, rank, calib, world, dev)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = run_cpu_baseline(calib, poses, args.cpu_seconds)

    # ---- the opt-in single-pass decode kernel on the same batch, off the headline number -------
    single = None
    if world == 1 and not args.no_single_pass:
        try:
            os.environ["VELOSLAM_SINGLE_PASS"] = "1"   # read by vs_create
            ctx2 = capi.Context(local, max_batch_packets=n_sub, max_poses=len(poses[0]) + 8, n_slots=2)
            ctx2.set_calibration(calib)
            ctx2.set_poses(poses[0], poses[1])

            def submit2():
                return ctx2.submit(d_pk, d_t, n=n_sub, stride=1206, n_halo=halo,
                                   mode=capi.MODE_STREAMING, flags=capi.FLAG_DEVICE_INPUT,
                                   t_base_us=t_base)
            for _ in range(3):
                r2 = ctx2.wait(submit2(), frames=False)
            torch.cuda.synchronize()
            k2, dk2 = 8, []
            w0 = time.perf_counter()
            pend = submit2()
            for i in range(k2):
                nxt = submit2() if i + 1 < k2 else None
                dk2.append(ctx2.wait(pend, frames=False).decode_ms)
                pend = nxt
            torch.cuda.synchronize()
            ms2 = (time.perf_counter() - w0) / k2 * 1e3
            single = {"points_per_s_per_gpu": r2.n_points / (ms2 * 1e-3), "ms_per_step": ms2,
                      "k_decode_ms": float(np.mean(dk2)), "points": r2.n_points,
                      "same_point_count_as_two_pass": bool(r2.n_points == n_emitted),
                      "note": "VELOSLAM_SINGLE_PASS=1: k_pose_pre + k_decode<.,0,FUSED> (segmentation "
                              "scans inside the decode kernel, packets read once); host wall clock, "
                              "two batches in flight; slower than the default two-pass pipeline "
                              "(DESIGN.md 4)"}
            ctx2.close()
        except Exception as e:  # informational only: never breaks the headline line
            single = {"error": str(e)}
        finally:
            os.environ.pop("VELOSLAM_SINGLE_PASS", None)


    # ---- configs[3] as written: one pass over a 1-h recording, strong scaling, gather timed -----
    recording = None
    if args.recording_leg_hours > 0:
        ctx.close()
        ctx = None
        del d_pk, d_t
        torch.cuda.empty_cache()
        try:
            recording = run_recording(args.recording_leg_hours, 3, 1, local, rank, world, dev, calib)
        except SystemExit:
            raise
        except Exception as e:  # reported, never silently dropped
            recording = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "packets_per_gpu_per_step": n_per, "slots_per_gpu_per_step": n_per * 384,
                "points_emitted_per_gpu_per_step": n_emitted, "mode": "streaming (reference parity)",
                "input_bytes_per_gpu": n_sub * 1206,
                "l2": "inputs (1.26 GB) and outputs (8.4 GB) per step exceed the 126 MB L2",
                "slots_per_s": world * n_per * 384 / (ms_per_step * 1e-3),
                "global_frames": n_global_frames, "stitch_timestamp_mismatches": mism,
                "host_affinity": (f"{numa_cpus} of {host_cpus_all} CPUs (NVML: local to the GPU)"
                                  if numa_cpus else "unchanged"),
                "wall_ms_per_step": wall / args.steps * 1e3,
                "per_rank_ms_per_step": [round(x, 4) for x in per_rank_ms],
            },
            "roofline": roofline, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        if parity is not None:
            line["parity_window"] = parity["status"]
            line["parity"] = parity
        if e2e is not None:
            line["e2e"] = e2e
        if e2e_xyzi is not None:
            line["e2e_xyzi_13B"] = e2e_xyzi
        if e2e_frames is not None:
            line["e2e_frames"] = e2e_frames
        if pcie is not None:
            line["pcie_probe"] = pcie
        if e2e_facade is not None:
            line["e2e_facade"] = e2e_facade
        if hdl32 is not None:
            line["hdl32"] = hdl32
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if online is not None:
            line["online"] = online
        if online_udp is not None:
            line["online_udp"] = online_udp
        if deskew is not None:
            line["deskew_per_point"] = deskew
        if single is not None:
            line["single_pass_variant"] = single
        if recording is not None:
            line["recording"] = recording
        print(json.dumps(line), flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def run_recording_headline(args, local, rank, world, dev, host_cpus_all, numa_cpus):
    """`--recording-hours H`: the headline IS BASELINE.json configs[3] -- total work fixed, split
    over the ranks (strong scaling), collective and stitch inside the timed region."""
    import torch
    from veloslam_b200 import synth
    calib = synth.calib_hdl64()
    rec = run_recording(args.recording_hours, args.steps, max(args.warmup, 3), local, rank, world,
                        dev, calib)
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = HBM_FALLBACK_GBS, "fallback"
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured"
        except Exception:
            pass
        alg = 1206 * rec["packets_this_rank"] + BYTES_PER_POINT_OUT * rec["points_this_rank"]
        line = {
            "metric": METRIC, "value": rec["points_per_s"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_pass"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (generated on the device)",
            "config": {"workload": f"offline replay of {args.recording_hours:g} h of synthetic HDL-64E "
                                   f"({rec['packets']} packets, {rec['points']} points) decode + deskew, "
                                   f"packet-range sharded across {world} B200 with a 512-packet halo "
                                   "(BASELINE.json configs[3]); one pass per step, frame-table "
                                   "all-gather + stitch inside the timed region",
                       "l2": "inputs and outputs per pass exceed the 126 MB L2 by orders of magnitude",
                       "host_affinity": (f"{numa_cpus} of {host_cpus_all} CPUs" if numa_cpus else "unchanged")},
            "roofline": {"bound": "hbm", "kernel": "whole pass of rank 0 (k_scan + k_pose + k_decode + "
                                                   "k_frames)", "achieved": alg / (rec["decode_ms_this_rank"] * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                         "frac": alg / (rec["decode_ms_this_rank"] * 1e-3) / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg, "traffic": None},
            "gpu_launches": rec["launches_per_pass"] * args.steps, "clocks": rec["clocks"],
            "parity_window": rec["parity_window"], "recording": rec,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def run_e2e(args, ctx_args, b, t, t_base, world, dev, layout=False, xyzi_only=False):
    """Host packets in, host points out, with two result slots so copies overlap the kernels.
    layout=False: vs_submit / vs_wait / vs_fetch_points, the stream-order columns (18 B/point).
    layout=True: vs_submit / vs_wait / vs_layout_frames / vs_fetch_layout, every frame as the
    reference's HDLFrame holds it (laser-major PointXYZI + PointMeta records, 28 B/point) -- what
    the C++ facade hands to its callers without touching a point on the CPU."""
    if layout:
        return run_e2e_frames(args, ctx_args, b, t, t_base, world, dev)
    import torch
    from veloslam_b200 import capi
    local, calib, poses = ctx_args
    chunk = min(args.e2e_chunk, b.shape[0])
    n_chunks = b.shape[0] // chunk
    ctx = capi.Context(local, max_batch_packets=chunk, max_poses=len(poses[0]) + 8, n_slots=2)
    ctx.set_calibration(calib)
    ctx.set_poses(poses[0], poses[1])
    # pinned host input (the caller's ring) and pinned host output columns, one set per slot
    h_pk = torch.from_numpy(b[:n_chunks * chunk]).pin_memory()
    h_t = torch.from_numpy(np.ascontiguousarray(t[:n_chunks * chunk])).pin_memory()
    cap = chunk * 384
    outs = []
    for _ in range(2):
        cols = [torch.empty(cap, dtype=dt).pin_memory() for dt in
                (torch.float32, torch.float32, torch.float32, torch.uint8, torch.uint8,
                 torch.int16, torch.int16, torch.int32)]
        outs.append(cols)

    bytes_per_point = 13 if xyzi_only else BYTES_PER_POINT_E2E

    def ptrs_of(slot):
        # x, y, z, intensity, laser, azimuth, distance; t_us stays on the device (NULL).
        # xyzi_only: the 13 B/point a consumer of coordinates + intensity needs (laser, azimuth
        # and raw distance are recomputable from the packets the host already holds)
        if xyzi_only:
            return [x.data_ptr() for x in outs[slot][:4]] + [None] * 4
        return [x.data_ptr() for x in outs[slot][:7]] + [None]

    def one_pass():
        carry = capi.carry_init()
        pending = None
        total = 0
        h2d = d2h = 0
        for c in range(n_chunks):
            a = c * chunk
            # the carry of chunk c is known only after chunk c-1 finished (a 72-byte state);
            # its kernels are short next to the copies, so wait c-1 first, then overlap the
            # D2H of c-1 with the H2D + kernels of c
            if pending is not None:
                rprev = ctx.wait(pending[0], frames=False)
                carry = rprev.carry_out
            tk = ctx.submit(h_pk[a:a + chunk], h_t[a:a + chunk], n=chunk, stride=1206,
                            mode=capi.MODE_STREAMING, flags=0, t_base_us=t_base, carry=carry)
            h2d += chunk * (1206 + 8)
            if pending is not None:
                ctx.fetch_into(pending[0], 0, rprev.n_points, ptrs_of(pending[1]))
                total += rprev.n_points
                d2h += rprev.n_points * bytes_per_point
            pending = (tk, c % 2)
        rprev = ctx.wait(pending[0], frames=False)
        ctx.fetch_into(pending[0], 0, rprev.n_points, ptrs_of(pending[1]))
        total += rprev.n_points
        d2h += rprev.n_points * bytes_per_point
        return total, h2d, d2h

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    one_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        total, h2d, d2h = one_pass()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        dt = float(tt.item())
        pts = torch.tensor([total], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(pts)
        total = int(pts.item())
    ctx.close()
    if xyzi_only:
        return {"value": total * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "chunk_packets": chunk,
                "bytes_per_point": 13,
                "note": "as e2e, fetching x, y, z, intensity only (vs_fetch_points with NULL for the "
                        "other columns): 13 B/point over PCIe"}
    return {"value": total * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
            "chunk_packets": chunk, "note": "host pinned packets -> vs_submit/vs_wait -> "
            "vs_fetch_points into pinned host columns: the 7 columns the HDLFrame facade "
            "consumes (x, y, z, intensity, laser, azimuth, distance = 18 B/point; t_us has no "
            "counterpart in HDLFrame and stays on the device); PCIe D2H bound"}


def run_e2e_frames(args, ctx_args, b, t, t_base, world, dev):
    import torch
    from veloslam_b200 import capi
    local, calib, poses = ctx_args
    chunk = min(args.e2e_chunk, b.shape[0])
    n_chunks = b.shape[0] // chunk
    ctx = capi.Context(local, max_batch_packets=chunk, max_poses=len(poses[0]) + 8, n_slots=2)
    ctx.set_calibration(calib)
    ctx.set_poses(poses[0], poses[1])
    h_pk = torch.from_numpy(b[:n_chunks * chunk]).pin_memory()
    h_t = torch.from_numpy(np.ascontiguousarray(t[:n_chunks * chunk])).pin_memory()
    cap = chunk * 384 + 400 * 384     # + room for the carried rows of an open rotation
    outs = [(torch.empty(cap * 16, dtype=torch.uint8).pin_memory(),
             torch.empty(cap * 12, dtype=torch.uint8).pin_memory()) for _ in range(2)]

    def one_pass():
        carry = capi.carry_init()
        carried = np.zeros(64, np.uint32)
        total = h2d = d2h = 0
        prev = prev2 = None       # (ticket, slot) submitted / laid out and being fetched
        for c in range(n_chunks + 2):
            if prev is not None:
                rp = ctx.wait(prev[0], frames=False)          # kernels of c-1 done long ago
                carry = rp.carry_out
            if prev2 is not None:
                ctx.sync(prev2[0])                            # copies of c-2 had a whole iteration
                prev2 = None
            cur = None
            if c < n_chunks:
                a = c * chunk
                tk = ctx.submit(h_pk[a:a + chunk], h_t[a:a + chunk], n=chunk, stride=1206,
                                mode=capi.MODE_STREAMING, flags=0, t_base_us=t_base, carry=carry)
                h2d += chunk * (1206 + 8)
                cur = (tk, c % 2)
            if prev is not None:
                lay, rows = ctx.layout_frames(prev[0], carried if carried.any() else None, 16, True)
                n_slots = int(lay.n_slots)
                assert n_slots <= cap
                if n_slots:
                    ctx.fetch_layout(prev[0], 0, n_slots, outs[prev[1]][0].data_ptr(),
                                     outs[prev[1]][1].data_ptr())
                carried = rows["row_count"][-1].astype(np.uint32)
                total += rp.n_points
                d2h += n_slots * 28
                prev2 = prev
            prev = cur
        return total, h2d, d2h

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    one_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        total, h2d, d2h = one_pass()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    per_rank_gbs = d2h * args.e2e_steps / dt / 1e9
    gbs = [per_rank_gbs]
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        allg = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        torch.distributed.all_gather(allg, torch.tensor([per_rank_gbs], dtype=torch.float64, device=dev))
        gbs = [float(x.item()) for x in allg]
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        dt = float(tt.item())
        pts = torch.tensor([total], dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(pts)
        total = int(pts.item())
    ctx.close()
    return {"value": total * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "chunk_packets": chunk,
            "bytes_per_point": 28, "per_rank_d2h_gb_per_s": [round(g, 2) for g in gbs],
            "note": "host pinned packets -> vs_submit/vs_wait -> vs_layout_frames (k_layout on the "
                    "device) -> vs_fetch_layout into pinned host memory: every frame laser-major as "
                    "HDLFrame::points[row] / pointsMeta[row] hold it (16-B PointXYZI + 12-B PointMeta "
                    "per point); PCIe D2H bound"}


def run_facade(args, local, calib, poses, b, t, world, dev):
    """The C++ facade driven like the reference's consumer (HDLSource.cxx:209-225):
    HDLParser::processHDLPacket per packet, getAllFrames / clearAllFrames, through
    tests/cpp/facade_driver (one process per GPU).  Pipelined throughput mode, frames handed out
    as zero-copy HDLFrames."""
    import torch
    from veloslam_b200 import calibxml
    from veloslam_b200.build import DRIVER_EXE, build_facade
    build_facade()
    n = min(args.facade_packets, b.shape[0])
    passes = 3
    tmp = tempfile.mkdtemp(prefix="vs_facade_")
    out = {}
    try:
        b[:n].tofile(os.path.join(tmp, "pk.bin"))
        np.ascontiguousarray(t[:n]).astype("<i8").tofile(os.path.join(tmp, "t.bin"))
        span_s = (passes + 2) * n * 288e-6
        from veloslam_b200 import synth
        pt, trv = synth.ins_trajectory(int(span_s * 100) + 40)
        rec = np.zeros(len(pt), dtype=[("t", "<i8"), ("v", "<f8", (9,))])
        rec["t"] = pt
        rec["v"] = trv
        rec.tofile(os.path.join(tmp, "poses.bin"))
        calibxml.write_db_xml(os.path.join(tmp, "db.xml"), calib)
        for name, store, meta in (("frames_xyzi_meta", 0, 1), ("frames_xyzi_only", 0, 0),
                                  ("frames_with_raw_packets", 1, 1)):
            if world > 1:
                torch.distributed.barrier()
            cmd = [DRIVER_EXE, "bench", os.path.join(tmp, "db.xml"), os.path.join(tmp, "pk.bin"),
                   os.path.join(tmp, "t.bin"), os.path.join(tmp, "poses.bin"), str(args.facade_batch),
                   str(passes), str(store), str(meta), "1", str(local)]
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            if p.returncode != 0:
                out[name] = {"error": (p.stderr or p.stdout)[-400:]}
                continue
            j = json.loads(p.stdout.strip().splitlines()[-1])
            pts, sec = j["points"], j["seconds"]
            if world > 1:
                tt = torch.tensor([sec], dtype=torch.float64, device=dev)
                torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
                sec = float(tt.item())
                pp = torch.tensor([pts], dtype=torch.int64, device=dev)
                torch.distributed.all_reduce(pp)
                pts = int(pp.item())
            out[name] = {"value": pts / sec, "unit": UNIT, "points": pts, "frames": j["frames"],
                         "seconds": sec, "packets": j["packets"],
                         "pinned_pool_bytes": j["pinned_pool_bytes"], "host_seconds": j.get("host"),
                         # a bare per-packet memcpy loop over the same packets (no parser, no GPU):
                         # what any per-packet API costs on this host
                         "memcpy_floor_points_per_s": (j["points"] / j["passes"] / j["memcpy_floor_seconds_per_pass"]
                                                       if j.get("memcpy_floor_seconds_per_pass") else None)}
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    out["batch_packets"] = args.facade_batch
    out["packets_per_pass"] = n
    out["note"] = ("tests/cpp/facade_driver bench: HDLParser::processHDLPacket per packet + "
                   "getAllFrames/clearAllFrames (the reference's HDLSource.cxx:209-225 consumer "
                   "loop), setPipelined(true); frames arrive as HDLFrames whose points[row] vectors "
                   "adopt page-locked arenas the GPU-built layout was copied into (28 B/point with "
                   "PointMeta, 16 without); the third figure also copies every raw packet into "
                   "HDLFrame::packets as the reference does")
    return out


def run_hdl32(local, dev):
    """HDL-32E (configs[0]'s sensor: per-return azimuth / time adjustment, k_decode<1>), device
    resident, 262 144 packets per batch, no poses."""
    import torch
    from veloslam_b200 import capi, synth
    n = 1 << 18
    pk, t = synth.hdl32_packets(n)
    ctx = capi.Context(local, max_batch_packets=n, max_poses=8, n_slots=2)
    ctx.set_calibration(synth.calib_hdl32())
    d_pk = torch.from_numpy(synth.as_bytes(pk)).to(dev)
    d_t = torch.from_numpy(np.ascontiguousarray(t)).to(dev)

    def sub():
        return ctx.submit(d_pk, d_t, n=n, stride=1206, n_halo=0, mode=capi.MODE_STREAMING,
                          flags=capi.FLAG_DEVICE_INPUT, t_base_us=int(t[0]))
    for _ in range(3):
        r = ctx.wait(sub(), frames=False)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(ctx.stream(0), device=dev)
    ext1 = torch.cuda.ExternalStream(ctx.stream(1), device=dev)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    ext1.wait_event(e0)
    k, dec = 20, []
    pend = sub()
    for i in range(k):
        nxt = sub() if i + 1 < k else None
        r = ctx.wait(pend, frames=False)
        dec.append(r.decode_ms)
        pend = nxt
    ej = torch.cuda.Event()
    ej.record(ext1)
    ext.wait_event(ej)
    e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    ctx.close()
    return {"points_per_s_per_gpu": r.n_points / (ms * 1e-3), "ms_per_step": ms,
            "k_decode_ms": float(np.mean(dec)), "packets": n, "points": r.n_points,
            "note": "HDL-32E stream, 32-laser calibration (LUT branch), no poses; two batches in flight"}


def run_pcie_probe(world, dev, seconds=0.4):
    """What the box's host side can move: every rank copies a 256 MiB pinned buffer device -> host
    (then host -> device) back to back for `seconds`, all ranks at once, nothing else running.
    The ceiling the end-to-end figures are bound by (per rank and in aggregate)."""
    import torch
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    out = {}
    for name, src, dst in (("d2h", d, h), ("h2d", h, d)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        k = 0
        while time.perf_counter() - t0 < seconds:
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            k += 4
        gbs = k * n / (time.perf_counter() - t0) / 1e9
        if world > 1:
            allg = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            torch.distributed.all_gather(allg, torch.tensor([gbs], dtype=torch.float64, device=dev))
            per = [float(x.item()) for x in allg]
        else:
            per = [gbs]
        out[name + "_gb_per_s_per_rank"] = [round(x, 1) for x in per]
        out[name + "_gb_per_s_total"] = round(sum(per), 1)
    out["note"] = ("pinned 256 MiB copies, all ranks concurrently, no kernels: the host-side ceiling "
                   "of this box for the e2e figures (18 or 28 bytes leave the GPU per point)")
    return out


def run_online_udp(args, local, rank, calib, world, dev):
    """BASELINE.json configs[4]: one paced HDL-64E stream PER GPU over loopback UDP at the
    sensor's own rate (3472 packets/s, 10 rotations/s) through the C++ facade's online path
    (HDLManager::startOnline -> HDLSource receive ring -> TimeSolver -> HDLParser on the GPU ->
    HDLManager::addFrame), one tests/cpp/facade_driver process per rank, all ranks at once.
    Per rotation: time from the closing packet entering the parser to the frame sitting in the
    manager (GPU round trip + zero-copy HDLFrame), and the same from the packet's arrival."""
    import torch
    from veloslam_b200 import calibxml, synth
    from veloslam_b200.build import DRIVER_EXE, build_facade
    build_facade()
    seconds = args.online_udp_seconds
    n = int(seconds * 3472) + 400
    pk, t = synth.hdl64_stream_tiled(n)
    tmp = tempfile.mkdtemp(prefix="vs_online_")
    try:
        synth.as_bytes(pk).tofile(os.path.join(tmp, "pk.bin"))
        pt, trv = synth.ins_trajectory(int(n * 288e-6 * 100) + 60)
        rec = np.zeros(len(pt), dtype=[("t", "<i8"), ("v", "<f8", (9,))])
        rec["t"] = pt
        rec["v"] = trv
        rec.tofile(os.path.join(tmp, "poses.bin"))
        calibxml.write_db_xml(os.path.join(tmp, "db.xml"), calib)
        if world > 1:
            torch.distributed.barrier()
        cmd = [DRIVER_EXE, "online", os.path.join(tmp, "db.xml"), os.path.join(tmp, "pk.bin"),
               os.path.join(tmp, "poses.bin"), str(seconds), str(local), str(23680 + 16 * rank)]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=int(seconds) + 120)
        if p.returncode != 0:
            mine = {"error": (p.stderr or p.stdout)[-300:]}
        else:
            mine = json.loads(p.stdout.strip().splitlines()[-1])
    finally:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    if world > 1:
        allo = [None] * world
        torch.distributed.all_gather_object(allo, mine)
    else:
        allo = [mine]
    ok = [o for o in allo if "error" not in o]
    out = {"streams": world, "seconds": seconds, "paced_packets_per_s_per_stream": 3472,
           "per_stream": allo,
           "note": "one paced UDP stream per GPU through HDLManager::startOnline (C++ facade); "
                   "rotation latency = closing packet handed to HDLParser -> HDLFrame (zero-copy, "
                   "laid out on the GPU) stored by HDLManager::addFrame; from_arrival adds the wait "
                   "in the receive ring; a 10 Hz sensor leaves 100 ms per rotation"}
    if ok:
        out["rotation_p50_ms_worst_stream"] = max(o["rotation_p50_ms"] for o in ok)
        out["rotation_p99_ms_worst_stream"] = max(o["rotation_p99_ms"] for o in ok)
        out["from_arrival_p99_ms_worst_stream"] = max(o["from_arrival_p99_ms"] for o in ok)
        out["dropped_packets"] = sum(o["dropped"] for o in ok)
        out["frames"] = sum(o["frames"] for o in ok)
    return out


def run_online(local, calib, poses, b, t, t_base, rotations=300):
    """BASELINE.json configs[4] on this rank's GPU: one stream, rotation-sized batches (347
    packets, 133 k points), host packets in -> frame index table on the host, back to back.
    Latency = wall time of vs_submit + vs_wait (H2D of the packets, the four kernels, D2H of the
    batch header and frame rows); the second figure adds the D2H of the facade's 7 columns."""
    import torch
    from veloslam_b200 import capi
    rot = 347
    n_rot = min(rotations, b.shape[0] // rot)
    ctx = capi.Context(local, max_batch_packets=512, max_poses=len(poses[0]) + 8, n_slots=1)
    ctx.set_calibration(calib)
    ctx.set_poses(poses[0], poses[1])
    h_pk = torch.from_numpy(b[:n_rot * rot]).pin_memory()
    h_t = torch.from_numpy(np.ascontiguousarray(t[:n_rot * rot])).pin_memory()
    cols = [torch.empty(512 * 384, dtype=dt).pin_memory() for dt in
            (torch.float32, torch.float32, torch.float32, torch.uint8, torch.uint8, torch.int16,
             torch.int16)]
    ptrs = [c.data_ptr() for c in cols] + [None]
    out = {}
    for with_points in (False, True):
        carry = capi.carry_init()
        lat = []
        pts = 0
        t_all0 = time.perf_counter()
        for r in range(n_rot):
            a = r * rot
            t0 = time.perf_counter()
            tk = ctx.submit(h_pk[a:a + rot], h_t[a:a + rot], n=rot, stride=1206,
                            mode=capi.MODE_STREAMING, flags=0, t_base_us=t_base, carry=carry)
            res = ctx.wait(tk, frames=False)
            if with_points and res.n_points:
                ctx.fetch_into(tk, 0, res.n_points, ptrs)
            lat.append(time.perf_counter() - t0)
            carry = res.carry_out
            pts += res.n_points
        dt_all = time.perf_counter() - t_all0
        lat = np.array(lat[10:]) * 1e3  # first rotations warm the path up
        out["with_points" if with_points else "index_only"] = {
            "p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
            "max_ms": float(lat.max()), "points_per_s": pts / dt_all}
    ctx.close()
    # the same loop with the C ABI called from C++ (tests/cpp/facade_driver latency): what the
    # figures above cost without the interpreter between vs_submit and vs_wait
    try:
        from veloslam_b200 import calibxml
        from veloslam_b200.build import DRIVER_EXE, build_facade
        build_facade()
        tmp = tempfile.mkdtemp(prefix="vs_lat_")
        try:
            b[:n_rot * rot].tofile(os.path.join(tmp, "pk.bin"))
            np.ascontiguousarray(t[:n_rot * rot]).astype("<i8").tofile(os.path.join(tmp, "t.bin"))
            rec = np.zeros(len(poses[0]), dtype=[("t", "<i8"), ("v", "<f8", (9,))])
            rec["t"] = poses[0]
            rec["v"] = poses[1]
            rec.tofile(os.path.join(tmp, "poses.bin"))
            calibxml.write_db_xml(os.path.join(tmp, "db.xml"), calib)
            for name, wp in (("index_only_cpp", 0), ("with_points_cpp", 1)):
                cmd = [DRIVER_EXE, "latency", os.path.join(tmp, "db.xml"), os.path.join(tmp, "pk.bin"),
                       os.path.join(tmp, "t.bin"), os.path.join(tmp, "poses.bin"), str(rot), str(local),
                       str(wp)]
                pr = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
                out[name] = (json.loads(pr.stdout.strip().splitlines()[-1]) if pr.returncode == 0
                             else {"error": (pr.stderr or pr.stdout)[-300:]})
        finally:
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
    except Exception as e:  # the driver is test infrastructure: its absence does not fail the bench
        out["index_only_cpp"] = {"error": repr(e)[:300]}
    out["rotations"] = n_rot
    out["packets_per_rotation"] = rot
    out["note"] = ("one HDL-64E stream on this GPU, rotation-sized batches back to back (a 10 Hz "
                   "sensor leaves 100 ms per rotation); host wall clock around vs_submit+vs_wait")
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
