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
            parsers[i].process_packets(pk_bytes[a:b], t_us[a:b])

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
    n_threads = os.cpu_count() or 1
    # bounded sample per step: ~ cpu-seconds / (steps + warmup) of work on all threads
    probe_n = 4096 * n_threads
    pk, t = synth.hdl64_stream_tiled(probe_n)
    poses = synth.ins_trajectory(int(probe_n * 288e-6 * 100) + 40)
    dt, _ = cpu_time_threads(factory, calib, poses, synth.as_bytes(pk), t, n_threads)
    rate = probe_n / dt
    total_steps = args.steps + args.warmup
    per_step_s = min(10.0, max(1.0, 120.0 / total_steps))
    n = int(max(probe_n, min(rate * per_step_s, 1 << 21)))
    pk, t = synth.hdl64_stream_tiled(n)
    poses = synth.ins_trajectory(int(n * 288e-6 * 100) + 40)
    b = synth.as_bytes(pk)
    pts = count_points(pk)
    for _ in range(args.warmup):
        cpu_time_threads(factory, calib, poses, b, t, n_threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_time_threads(factory, calib, poses, b, t, n_threads)
    dt = time.perf_counter() - t0
    value = pts * args.steps / dt
    sample = (f"{n} HDL-64E packets per step ({n * 384} slots, {pts} points), "
              f"{n_threads} threads over packet-range shards")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "HDL-64E S2 decode + per-packet deskew against a 100 Hz INS "
                               "timeline (BASELINE.json configs[2]), CPU reference path",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
NCU_FULL_CSV = os.path.join(ROOT, "profiles", "r1z_k_decode_ncu_full.csv")
NCU_FULL_PACKETS = 1 << 20  # packets per launch of the captured k_decode


def ncu_traffic(n_per):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_decode launch from the committed
    `ncu --set full` capture (same workload and batch size), else None."""
    if n_per != NCU_FULL_PACKETS or not os.path.exists(NCU_FULL_CSV):
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    try:
        for line in open(NCU_FULL_CSV):
            name, unit, val = line.strip().split(",")[:3]
            if name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(val) * scale[unit]
    except Exception:
        return None, None
    return (tot if tot > 0 else None), os.path.relpath(NCU_FULL_CSV, ROOT)


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

    # ---- roofline of the dominant kernel (k_decode), this rank ------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = HBM_FALLBACK_GBS, "fallback"
    try:
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured"
    except Exception:
        pass
    alg_bytes = 1206 * n_per + BYTES_PER_POINT_OUT * n_emitted
    traffic, traffic_src = ncu_traffic(n_per)
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
                "note": "k_decode timed with CUDA events on its launch stream while the other "
                        "result slot's k_scan/k_pose may overlap it (two batches in flight)"}

    # ---- per-point deskew extension, same batch, off the headline number -------
    deskew = None
    if not args.no_deskew:
        off = np.zeros((12, 32), dtype=np.uint16)
        for j in range(12):   # S2-like timing: block pairs 48 us apart, lasers 1.5 us apart
            off[j] = np.round((j // 2) * 48.0 + np.arange(32) * 1.5).astype(np.uint16)
        ctx.set_firing_offsets(off)

        def step_dsk():
            return ctx.wait(ctx.submit(d_pk, d_t, n=n_sub, stride=1206, n_halo=halo,
                                       mode=capi.MODE_STREAMING,
                                       flags=capi.FLAG_DEVICE_INPUT | capi.FLAG_DESKEW_PER_POINT,
                                       t_base_us=t_base), frames=False)
        for _ in range(3):
            step_dsk()
        barrier()
        d0 = torch.cuda.Event(enable_timing=True)
        d1 = torch.cuda.Event(enable_timing=True)
        d0.record(ext)
        dk = []
        for _ in range(5):
            dk.append(step_dsk().decode_ms)
        d1.record(ext)
        barrier()
        ms = d0.elapsed_time(d1) / 5
        deskew = {"points_per_s_per_gpu": n_emitted / (ms * 1e-3), "ms_per_step": ms,
                  "k_decode_ms": float(np.mean(dk)),
                  "note": "VS_FLAG_DESKEW_PER_POINT: slerp/lerp pose per point, re-based to the "
                          "frame origin (not a reference behaviour; one batch in flight)"}
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


    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "HDL-64E S2 decode + rotation segmentation + per-packet deskew against "
                            "a 100 Hz INS timeline (BASELINE.json configs[2]; N>1: configs[3] "
                            "packet-range shards with a 512-packet halo)",
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
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if online is not None:
            line["online"] = online
        if deskew is not None:
            line["deskew_per_point"] = deskew
        if single is not None:
            line["single_pass_variant"] = single
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def run_e2e(args, ctx_args, b, t, t_base, world, dev):
    """Host packets in, host point columns out, through vs_submit / vs_wait / vs_fetch_points
    with two result slots so copies overlap the kernels."""
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

    def ptrs_of(slot):
        # x, y, z, intensity, laser, azimuth, distance; t_us stays on the device (NULL)
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
                d2h += rprev.n_points * BYTES_PER_POINT_E2E
            pending = (tk, c % 2)
        rprev = ctx.wait(pending[0], frames=False)
        ctx.fetch_into(pending[0], 0, rprev.n_points, ptrs_of(pending[1]))
        total += rprev.n_points
        d2h += rprev.n_points * BYTES_PER_POINT_E2E
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
    return {"value": total * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
            "chunk_packets": chunk, "note": "host pinned packets -> vs_submit/vs_wait -> "
            "vs_fetch_points into pinned host columns: the 7 columns the HDLFrame facade "
            "consumes (x, y, z, intensity, laser, azimuth, distance = 18 B/point; t_us has no "
            "counterpart in HDLFrame and stays on the device); PCIe D2H bound"}


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
