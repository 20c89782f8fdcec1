"""ctypes binding of the C ABI declared in include/veloslam_b200.h.

This is harness code for tests and bench.py: the product is the shared library.  It fails
loudly when the CUDA library is missing -- there is no CPU fallback and this module never
touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB

INT64_MIN = -(2 ** 63)
VS_TIME_NONE = INT64_MIN
MODE_STREAMING, MODE_OFFLINE = 0, 1
FLAG_DEVICE_INPUT, FLAG_PCAP_TIMES, FLAG_DESKEW_PER_POINT, FLAG_NO_FRAME_LIST = 1, 2, 4, 8
STATUS = {0: "VS_OK", 1: "VS_ERR_INVALID_ARG", 2: "VS_ERR_NOT_CALIBRATED", 3: "VS_ERR_CUDA",
          4: "VS_ERR_CAPACITY", 5: "VS_ERR_NO_DEVICE", 6: "VS_ERR_HALO", 7: "VS_ERR_STATE"}

EXPORTS = ["vs_version", "vs_create", "vs_destroy", "vs_last_error", "vs_set_calibration",
           "vs_set_filters", "vs_set_poses", "vs_interpolate", "vs_carry_init", "vs_submit",
           "vs_wait", "vs_fetch_points", "vs_read_frame_information", "vs_host_alloc",
           "vs_host_free", "vs_stream", "vs_slot_stream", "vs_device_alloc", "vs_device_free",
           "vs_device_upload", "vs_solve_packet_times", "vs_poses_from_ins", "vs_set_firing_offsets",
           "vs_layout_frames", "vs_fetch_layout", "vs_sync", "vs_shard_range", "vs_frame_table_rows",
           "vs_frame_table_rows_device", "vs_stitch_frame_tables"]


class LaserCorr(C.Structure):
    _fields_ = [("rot_correction_deg", C.c_double), ("vert_correction_deg", C.c_double),
                ("dist_correction_cm", C.c_double), ("vert_offset_correction_cm", C.c_double),
                ("horiz_offset_correction_cm", C.c_double)]


class TimeSolver(C.Structure):
    """vs_time_solver: state of TimeSolver::calcTimestamp(uint32_t)."""
    _fields_ = [("base_us", C.c_int64), ("last_report", C.c_uint32), ("inited", C.c_int32)]


# vs_ins_pva (NovAtel INSPVA, reference type_defs.h:39-58) as a numpy record
INS_PVA_DTYPE = np.dtype([("message_id", "<u2"), ("week_number", "<u2"), ("milliseconds", "<u4"),
                          ("week_number_pos", "<u4"), ("pad0", "<u4"), ("seconds_pos", "<f8"),
                          ("llh", "<f8", (3,)), ("v", "<f8", (3,)), ("eulr", "<f8", (3,)),
                          ("ins_status", "<i4"), ("pad1", "<i4")])


class Filters(C.Structure):
    _fields_ = [("laser_mask", C.c_uint64), ("points_skip", C.c_int32),
                ("crop_returns", C.c_int32), ("crop_inside", C.c_int32), ("reserved", C.c_int32),
                ("crop_region", C.c_double * 6)]


class Carry(C.Structure):
    _fields_ = [("last_azimuth", C.c_int32), ("firing_skip", C.c_int32),
                ("frame_meta_inited", C.c_int32), ("is_hdl64", C.c_int32),
                ("origin_T", C.c_double * 3), ("frame_timestamp_us", C.c_int64),
                ("frame_skips", C.c_int32), ("frame_carpose_valid", C.c_int32),
                ("frame_carpose", C.c_double * 9), ("frames_closed", C.c_int64),
                ("points_emitted", C.c_int64), ("packets_seen", C.c_int64)]


class Frame(C.Structure):
    _fields_ = [("first_point", C.c_int64), ("n_points", C.c_int64), ("timestamp_us", C.c_int64),
                ("start_packet", C.c_int32), ("start_block", C.c_int32),
                ("meta_packet", C.c_int32), ("skips", C.c_int32), ("closed", C.c_int32),
                ("hdl64_order", C.c_int32), ("carpose_valid", C.c_int32), ("reserved", C.c_int32),
                ("carpose", C.c_double * 9), ("laser_counts", C.c_uint32 * 64)]


class Result(C.Structure):
    _fields_ = [("n_packets", C.c_int64), ("n_points", C.c_int64), ("n_frames", C.c_int32),
                ("n_closed", C.c_int32), ("x", C.c_void_p), ("y", C.c_void_p), ("z", C.c_void_p),
                ("intensity", C.c_void_p), ("laser", C.c_void_p), ("azimuth", C.c_void_p),
                ("distance", C.c_void_p), ("t_us", C.c_void_p), ("frames", C.POINTER(Frame)),
                ("carry_out", Carry), ("t_base_us", C.c_int64), ("first_upper_block", C.c_int64),
                ("gpu_ms", C.c_float), ("decode_ms", C.c_float), ("n_kernel_launches", C.c_int32),
                ("reserved", C.c_int32)]


class FrameRows(C.Structure):
    """vs_frame_rows: where the rows of one frame sit in the HDLFrame layout arrays."""
    _fields_ = [("first_slot", C.c_int64), ("n_slots", C.c_int64),
                ("row_start", C.c_uint32 * 64), ("row_count", C.c_uint32 * 64),
                ("row_carried", C.c_uint32 * 64), ("row_laser", C.c_int32 * 64)]


class Layout(C.Structure):
    _fields_ = [("xyzi", C.c_void_p), ("meta", C.c_void_p), ("n_slots", C.c_int64),
                ("rows", C.POINTER(FrameRows)), ("n_frames", C.c_int32),
                ("xyzi_stride", C.c_int32), ("n_kernel_launches", C.c_int32),
                ("reserved", C.c_int32)]


class FrameSegment(C.Structure):
    _fields_ = [("rank", C.c_int32), ("reserved", C.c_int32), ("first_point", C.c_int64),
                ("n_points", C.c_int64)]


class GlobalFrame(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("start_packet", C.c_int64), ("timestamp_us", C.c_int64),
                ("start_block", C.c_int32), ("skips", C.c_int32), ("closed", C.c_int32),
                ("hdl64_order", C.c_int32), ("first_segment", C.c_int32), ("n_segments", C.c_int32),
                ("timestamp_mismatch", C.c_int32), ("reserved", C.c_int32)]


FRAME_ROW_COLS = 10

# PointMeta as the reference lays it out (type_defs.h:168-176), 12 bytes
POINT_META_DTYPE = np.dtype({"names": ["azimuth", "distance", "intensityFlag", "distanceFlag", "flags"],
                             "formats": ["<u2", "<f4", "u1", "u1", "u1"],
                             "offsets": [0, 4, 8, 9, 10], "itemsize": 12})


class VeloError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library():
    """dlopen libveloslam_b200.so; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise RuntimeError(f"{LIB} is missing: build it with __graft_entry__.build() "
                           "(python -m veloslam_b200.build); there is no CPU fallback")
    L = C.CDLL(LIB)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.vs_version.restype = C.c_char_p
    L.vs_version.argtypes = []
    L.vs_create.restype = C.c_int
    L.vs_create.argtypes = [C.c_int, i64, i64, C.c_int, C.POINTER(vp)]
    L.vs_destroy.restype = None
    L.vs_destroy.argtypes = [vp]
    L.vs_last_error.restype = C.c_char_p
    L.vs_last_error.argtypes = [vp]
    L.vs_set_calibration.restype = C.c_int
    L.vs_set_calibration.argtypes = [vp, C.POINTER(LaserCorr), C.c_int, C.c_int]
    L.vs_set_filters.restype = C.c_int
    L.vs_set_filters.argtypes = [vp, C.POINTER(Filters)]
    L.vs_set_poses.restype = C.c_int
    L.vs_set_poses.argtypes = [vp, vp, vp, i64]
    L.vs_interpolate.restype = C.c_int
    L.vs_interpolate.argtypes = [vp, i64, C.POINTER(C.c_double), C.POINTER(i32), C.POINTER(i32)]
    L.vs_carry_init.restype = None
    L.vs_carry_init.argtypes = [C.POINTER(Carry)]
    L.vs_submit.restype = C.c_int
    L.vs_submit.argtypes = [vp, vp, i64, vp, i64, i64, C.c_int, C.c_uint32, i64,
                            C.POINTER(Carry), C.POINTER(u64)]
    L.vs_wait.restype = C.c_int
    L.vs_wait.argtypes = [vp, u64, C.POINTER(Result)]
    L.vs_fetch_points.restype = C.c_int
    L.vs_fetch_points.argtypes = [vp, u64, i64, i64] + [vp] * 8
    L.vs_read_frame_information.restype = C.c_int
    L.vs_read_frame_information.argtypes = [vp, vp, i64, vp, i64, C.c_uint32, vp, vp, vp, i32,
                                            C.POINTER(i32)]
    L.vs_device_alloc.restype = C.c_int
    L.vs_device_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    L.vs_device_free.restype = None
    L.vs_device_free.argtypes = [vp, vp]
    L.vs_device_upload.restype = C.c_int
    L.vs_device_upload.argtypes = [vp, vp, vp, u64]
    L.vs_solve_packet_times.restype = C.c_int
    L.vs_solve_packet_times.argtypes = [vp, vp, i64, i64, C.c_uint32, i64, C.POINTER(TimeSolver), vp]
    L.vs_poses_from_ins.restype = C.c_int
    L.vs_poses_from_ins.argtypes = [vp, vp, i64, C.POINTER(C.c_double), vp, vp, vp]
    L.vs_set_firing_offsets.restype = C.c_int
    L.vs_set_firing_offsets.argtypes = [vp, vp]
    L.vs_stream.restype = vp
    L.vs_stream.argtypes = [vp]
    L.vs_slot_stream.restype = vp
    L.vs_slot_stream.argtypes = [vp, C.c_int]
    L.vs_layout_frames.restype = C.c_int
    L.vs_layout_frames.argtypes = [vp, u64, vp, C.c_int, C.c_int, C.POINTER(Layout)]
    L.vs_fetch_layout.restype = C.c_int
    L.vs_fetch_layout.argtypes = [vp, u64, i64, i64, vp, vp]
    L.vs_sync.restype = C.c_int
    L.vs_sync.argtypes = [vp, u64, C.POINTER(C.c_float)]
    L.vs_shard_range.restype = C.c_int
    L.vs_shard_range.argtypes = [i64, i32, i32, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.vs_frame_table_rows.restype = C.c_int
    L.vs_frame_table_rows.argtypes = [vp, i32, i32, i64, i64, vp]
    L.vs_frame_table_rows_device.restype = C.c_int
    L.vs_frame_table_rows_device.argtypes = [vp, C.c_uint64, i32, i64, vp, i64]
    L.vs_stitch_frame_tables.restype = C.c_int
    L.vs_stitch_frame_tables.argtypes = [vp, vp, i32, i64, vp, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    _lib = L
    return L


def shard_range(n_packets, world, rank, halo):
    """vs_shard_range: (first, n_halo, end) of rank's packet range."""
    f, h, e = C.c_int64(), C.c_int64(), C.c_int64()
    rc = load_library().vs_shard_range(n_packets, world, rank, halo, C.byref(f), C.byref(h), C.byref(e))
    if rc != 0:
        raise VeloError(rc, "vs_shard_range: bad arguments")
    return int(f.value), int(h.value), int(e.value)


def frame_table_rows(frame_table, rank, first_packet, n_halo):
    """vs_frame_table_rows over a vs_frame structured array: (n, 10) int64 exchange rows."""
    tab = np.ascontiguousarray(frame_table)
    assert tab.dtype.itemsize == C.sizeof(Frame)
    rows = np.zeros((tab.shape[0], FRAME_ROW_COLS), dtype=np.int64)
    rc = load_library().vs_frame_table_rows(_ptr(tab) if tab.shape[0] else None, tab.shape[0], rank,
                                            first_packet, n_halo, _ptr(rows) if tab.shape[0] else None)
    if rc != 0:
        raise VeloError(rc, "vs_frame_table_rows: bad arguments")
    return rows


def frame_table_rows_into(result, rank, first_packet, n_halo, out_rows):
    """vs_frame_table_rows straight from a BatchResult's context-owned vs_frame array into a
    caller-owned (cap, 10) int64 buffer (numpy array or pinned torch tensor): no intermediate
    copies.  Call before the slot's next submit.  Returns the number of rows."""
    ptr, n = result._raw_frames
    cap = int(out_rows.shape[0])
    if n > cap:
        raise VeloError(4, "frame_table_rows_into: buffer too small")
    rc = load_library().vs_frame_table_rows(C.cast(ptr, C.c_void_p), n, rank, first_packet, n_halo,
                                            _ptr(out_rows))
    if rc != 0:
        raise VeloError(rc, "vs_frame_table_rows: bad arguments")
    return int(n)


def stitch_frame_tables(tables):
    """vs_stitch_frame_tables over per-rank (n_g, 10) int64 tables (rank order): structured
    arrays (global frames, segments)."""
    world = len(tables)
    counts = np.array([t.shape[0] for t in tables], dtype=np.int32)
    total = int(counts.sum())
    rows = np.ascontiguousarray(np.concatenate([np.asarray(t, np.int64).reshape(-1, FRAME_ROW_COLS)
                                                for t in tables], axis=0)) if total else \
        np.zeros((0, FRAME_ROW_COLS), np.int64)
    frames = np.zeros(max(total, 1), dtype=np.dtype(GlobalFrame))
    segs = np.zeros(max(total, 1), dtype=np.dtype(FrameSegment))
    nf, ns = C.c_int32(), C.c_int32()
    rc = load_library().vs_stitch_frame_tables(_ptr(rows) if total else None, _ptr(counts), world, 0,
                                               _ptr(frames), frames.shape[0], _ptr(segs), segs.shape[0],
                                               C.byref(nf), C.byref(ns))
    if rc != 0:
        raise VeloError(rc, "vs_stitch_frame_tables failed")
    return frames[:nf.value], segs[:ns.value]


class Stitcher:
    """vs_stitch_frame_tables over the buffer a fixed-size all-gather leaves -- table g at row
    g * stride_rows -- into preallocated outputs: no concatenation, no per-call allocation."""

    def __init__(self, world, cap_rows_total):
        self.world = world
        self.frames = np.zeros(max(cap_rows_total, 1), dtype=np.dtype(GlobalFrame))
        self.segs = np.zeros(max(cap_rows_total, 1), dtype=np.dtype(FrameSegment))
        self.counts = np.zeros(world, dtype=np.int32)
        self._L = load_library()

    def __call__(self, rows_buffer, counts, stride_rows, first_row=0):
        """rows_buffer: int64 array (any shape) whose table g begins first_row + g * stride_rows
        rows in; counts: rows per rank."""
        self.counts[:] = counts
        nf, ns = C.c_int32(), C.c_int32()
        base = _ptr(rows_buffer) + first_row * FRAME_ROW_COLS * 8
        rc = self._L.vs_stitch_frame_tables(base, _ptr(self.counts), self.world, stride_rows,
                                            _ptr(self.frames), self.frames.shape[0], _ptr(self.segs),
                                            self.segs.shape[0], C.byref(nf), C.byref(ns))
        if rc != 0:
            raise VeloError(rc, "vs_stitch_frame_tables failed")
        return self.frames[:nf.value], self.segs[:ns.value]


def carry_init():
    c = Carry()
    load_library().vs_carry_init(C.byref(c))
    return c


def _ptr(a):
    """Address of a numpy array, a torch tensor (host or device) or a raw int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class FrameView:
    """Python copy of one vs_frame."""

    def __init__(self, f):
        self.first_point = int(f.first_point)
        self.n_points = int(f.n_points)
        self.timestamp_us = int(f.timestamp_us)
        self.start_packet = int(f.start_packet)
        self.start_block = int(f.start_block)
        self.meta_packet = int(f.meta_packet)
        self.skips = int(f.skips)
        self.closed = bool(f.closed)
        self.hdl64_order = bool(f.hdl64_order)
        self.carpose_valid = bool(f.carpose_valid)
        self.carpose = np.array(list(f.carpose), dtype=np.float64)
        self.laser_counts = np.array(list(f.laser_counts), dtype=np.int64)


FRAME_TABLE_DTYPE = np.dtype(Frame)   # vs_frame rows as a numpy structured array


class BatchResult:
    def __init__(self, ctx, ticket, r, frames=True):
        self._ctx = ctx
        self.ticket = ticket
        self.n_packets = int(r.n_packets)
        self.n_points = int(r.n_points)
        self.n_frames = int(r.n_frames)
        self.n_closed = int(r.n_closed)
        # (FLAG_NO_FRAME_LIST: the library assembles no list, r.frames is NULL)
        self.frames = [FrameView(r.frames[i]) for i in range(r.n_frames)] if frames and r.frames else None
        # the same rows as one structured array (copied out of the context-owned buffer)
        self._raw_frames = (r.frames, r.n_frames if r.frames else 0)
        self._frame_table = None
        self.carry_out = Carry.from_buffer_copy(r.carry_out)
        self.t_base_us = int(r.t_base_us)
        self.first_upper_block = int(r.first_upper_block)
        self.gpu_ms = float(r.gpu_ms)
        self.decode_ms = float(r.decode_ms)
        self.n_kernel_launches = int(r.n_kernel_launches)
        self.graph_launches = int(r.reserved)   # batches this slot issued as one CUDA graph
        self.device_ptrs = {k: getattr(r, k) for k in
                            ("x", "y", "z", "intensity", "laser", "azimuth", "distance", "t_us")}

    @property
    def frame_table(self):
        """vs_frame rows as one structured array (copied out of the context-owned buffer; call
        before the slot's next submit)."""
        if self._frame_table is None:
            ptr, n = self._raw_frames
            self._frame_table = np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n > 0 else \
                np.zeros(0, dtype=FRAME_TABLE_DTYPE)
        return self._frame_table

    def fetch(self, first=0, count=None, columns=None):
        """Copy point columns to host numpy arrays."""
        return self._ctx.fetch_points(self.ticket, first,
                                      self.n_points - first if count is None else count, columns)


COLUMNS = (("x", np.float32), ("y", np.float32), ("z", np.float32), ("intensity", np.uint8),
           ("laser", np.uint8), ("azimuth", np.uint16), ("distance", np.uint16),
           ("t_us", np.uint32))


class Context:
    """One parser + pose snapshot on one GPU (vs_ctx)."""

    def __init__(self, device=0, max_batch_packets=1 << 16, max_poses=1 << 16, n_slots=1):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.vs_create(device, max_batch_packets, max_poses, n_slots, C.byref(h))
        if rc != 0:
            raise VeloError(rc, self._L.vs_last_error(None).decode() or "vs_create failed")
        self._h = h
        self.device = device
        self.max_batch_packets = max_batch_packets

    def close(self):
        if getattr(self, "_h", None):
            self._L.vs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VeloError(rc, self._L.vs_last_error(self._h).decode())

    # -- configuration --------------------------------------------------------------
    def set_calibration(self, calib):
        n = calib.n_rows
        arr = (LaserCorr * max(n, 1))()
        for i in range(n):
            arr[i] = LaserCorr(calib.rot_deg[i], calib.vert_deg[i], calib.dist_cm[i],
                               calib.voff_cm[i], calib.hoff_cm[i])
        self._check(self._L.vs_set_calibration(self._h, arr, n, calib.n_enabled))

    def set_filters(self, laser_selection=None, points_skip=0, crop_returns=0, crop_inside=0,
                    crop_region=(0, 0, 0, 0, 0, 0)):
        mask = (1 << 64) - 1
        if laser_selection is not None:
            mask = 0
            for i, v in enumerate(laser_selection):
                if v:
                    mask |= 1 << i
        f = Filters(mask, points_skip, int(bool(crop_returns)), int(bool(crop_inside)), 0,
                    (C.c_double * 6)(*[float(v) for v in crop_region]))
        self._check(self._L.vs_set_filters(self._h, C.byref(f)))

    def set_firing_offsets(self, off_us):
        """12 x 32 firing offsets (us) for sensors without a built-in table (HDL-64E)."""
        o = np.ascontiguousarray(off_us, dtype=np.uint16).reshape(12, 32)
        self._check(self._L.vs_set_firing_offsets(self._h, _ptr(o)))

    def set_poses(self, t_us, trv):
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        v = np.ascontiguousarray(trv, dtype=np.float64).reshape(-1, 9) if len(t) else \
            np.zeros((0, 9))
        assert v.shape[0] == t.shape[0]
        self._check(self._L.vs_set_poses(self._h, _ptr(t) if len(t) else None,
                                         _ptr(v) if len(t) else None, len(t)))

    def interpolate(self, t_us):
        out = (C.c_double * 9)()
        found, valid = C.c_int32(), C.c_int32()
        self._check(self._L.vs_interpolate(self._h, int(t_us), out, C.byref(found),
                                           C.byref(valid)))
        return bool(found.value), np.array(list(out)), bool(valid.value)

    # -- online front end in batch form ----------------------------------
    def solve_packet_times(self, pkts, now_us, state=None, n=None, stride=None, flags=0, out=None):
        """TimeSolver::calcTimestamp(uint32_t) for a packet array; returns (times, state).
        Host arrays by default; with FLAG_DEVICE_INPUT pkts / out are device tensors."""
        if state is None:
            state = TimeSolver()
        if n is None:
            n = int(pkts.shape[0])
        if stride is None:
            stride = int(pkts.shape[1]) if pkts.ndim == 2 else 1206
        if out is None:
            out = np.empty(n, dtype=np.int64)
        self._check(self._L.vs_solve_packet_times(self._h, _ptr(pkts), stride, n, flags, int(now_us),
                                                  C.byref(state), _ptr(out)))
        return out, state

    def poses_from_ins(self, recs, origin_xyz, arrival_us):
        """INSPVA records -> (t_us, trv n x 9): calcTransform + calcTimestamp(InsPVA) on the GPU."""
        recs = np.ascontiguousarray(recs)
        assert recs.dtype.itemsize == INS_PVA_DTYPE.itemsize
        n = len(recs)
        arr = np.ascontiguousarray(np.broadcast_to(np.asarray(arrival_us, np.int64), (n,)))
        t = np.empty(n, dtype=np.int64)
        trv = np.empty((n, 9), dtype=np.float64)
        org = (C.c_double * 3)(*[float(v) for v in origin_xyz])
        self._check(self._L.vs_poses_from_ins(self._h, _ptr(recs), n, org, _ptr(arr), _ptr(t), _ptr(trv)))
        return t, trv

    # -- batches ----------------------------------------------------------------------
    def submit(self, pkts, pkt_time_us, n=None, stride=None, n_halo=0, mode=MODE_STREAMING,
               flags=0, t_base_us=VS_TIME_NONE, carry=None):
        """pkts: (n, stride) uint8 numpy array / torch tensor, or a raw address with n & stride."""
        if n is None:
            n = int(pkts.shape[0])
        if stride is None:
            stride = int(pkts.shape[1]) if pkts.ndim == 2 else 1206
        tk = C.c_uint64()
        cptr = C.byref(carry) if carry is not None else None
        self._check(self._L.vs_submit(self._h, _ptr(pkts), stride, _ptr(pkt_time_us), n, n_halo,
                                      mode, flags, t_base_us, cptr, C.byref(tk)))
        self._keepalive = (pkts, pkt_time_us)
        return tk.value

    def wait(self, ticket, frames=True):
        r = Result()
        self._check(self._L.vs_wait(self._h, ticket, C.byref(r)))
        return BatchResult(self, ticket, r, frames)

    def decode(self, pkts, pkt_time_us, **kw):
        return self.wait(self.submit(pkts, pkt_time_us, **kw))

    def fetch_points(self, ticket, first, count, columns=None):
        names = [c for c, _ in COLUMNS] if columns is None else list(columns)
        out = {}
        ptrs = []
        for name, dt in COLUMNS:
            if name in names:
                out[name] = np.empty(max(count, 0), dtype=dt)
                ptrs.append(out[name].ctypes.data if count > 0 else None)
            else:
                ptrs.append(None)
        self._check(self._L.vs_fetch_points(self._h, ticket, first, count, *ptrs))
        return out

    def fetch_into(self, ticket, first, count, host_ptrs):
        """vs_fetch_points into caller-owned (e.g. pinned) buffers: host_ptrs = 8 addresses."""
        self._check(self._L.vs_fetch_points(self._h, ticket, first, count, *host_ptrs))

    # -- HDLFrame layout on the device ------------------------------------------------------
    def layout_frames(self, ticket, carried_counts=None, xyzi_stride=16, with_meta=True):
        """vs_layout_frames: returns (Layout struct, rows as a structured numpy array copy)."""
        lay = Layout()
        car = None
        if carried_counts is not None:
            car = np.ascontiguousarray(carried_counts, dtype=np.uint32)
            assert car.shape == (64,)
        self._check(self._L.vs_layout_frames(self._h, ticket, _ptr(car), xyzi_stride,
                                             1 if with_meta else 0, C.byref(lay)))
        rows = np.ctypeslib.as_array(lay.rows, shape=(lay.n_frames,)).copy() if lay.n_frames else \
            np.zeros(0, dtype=np.dtype(FrameRows))
        return lay, rows

    def fetch_layout(self, ticket, first_slot, n_slots, xyzi_host=None, meta_host=None):
        """Enqueue the D2H copy of layout slots into host buffers (addresses / arrays)."""
        self._check(self._L.vs_fetch_layout(self._h, ticket, first_slot, n_slots, _ptr(xyzi_host),
                                            _ptr(meta_host)))

    def frame_table_rows_device(self, ticket, rank, first_packet, d_rows, cap_rows):
        """vs_frame_table_rows_device: enqueue the kernel that writes the batch's exchange rows
        (row 0 = [n_rows, 0...]) into the device buffer d_rows ((cap_rows + 1, 10) int64: a torch
        cuda tensor or an address) behind the batch's kernels.  No synchronisation."""
        self._check(self._L.vs_frame_table_rows_device(self._h, ticket, rank, first_packet, _ptr(d_rows),
                                                       cap_rows))

    def sync(self, ticket):
        ms = C.c_float()
        self._check(self._L.vs_sync(self._h, ticket, C.byref(ms)))
        return float(ms.value)

    def fetch_frames_layout(self, ticket, carried_counts=None, with_meta=True):
        """Convenience for tests: whole layout of a finished batch as numpy arrays
        (xyzi (n_slots, 4) float32, meta POINT_META_DTYPE or None, rows)."""
        lay, rows = self.layout_frames(ticket, carried_counts, 16, with_meta)
        n = int(lay.n_slots)
        xyzi = np.zeros((n, 4), dtype=np.float32)
        meta = np.zeros(n, dtype=POINT_META_DTYPE) if with_meta else None
        if n:
            self.fetch_layout(ticket, 0, n, xyzi, meta)
        self.sync(ticket)
        return xyzi, meta, rows

    def read_frame_information(self, pkts, pkt_time_us, flags=0):
        n, stride = int(pkts.shape[0]), int(pkts.shape[1])
        cap = 12 * n + 1
        sp = np.zeros(cap, np.int32)
        sk = np.zeros(cap, np.int32)
        ts = np.zeros(cap, np.int64)
        nf = C.c_int32()
        self._check(self._L.vs_read_frame_information(
            self._h, _ptr(pkts), stride, _ptr(pkt_time_us), n, flags, _ptr(sp), _ptr(sk),
            _ptr(ts), cap, C.byref(nf)))
        return sp[:nf.value].copy(), sk[:nf.value].copy(), ts[:nf.value].copy()

    def stream(self, slot=0):
        return self._L.vs_slot_stream(self._h, slot)
