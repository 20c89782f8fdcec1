"""ctypes binding of the CPU oracle (oracle/libvelo_oracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from veloslam_b200/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB, build_oracle

INT64_MIN = -(2 ** 63)


class FrameInfo(C.Structure):
    _fields_ = [("timestamp_us", C.c_int64), ("skips", C.c_int32), ("n_lasers", C.c_int32),
                ("n_points", C.c_int32), ("n_packets", C.c_int32), ("is_hdl64_order", C.c_int32),
                ("pad", C.c_int32), ("carpose_TRV", C.c_double * 9),
                ("carpose_seconds_pos", C.c_double)]


_lib = None


class TimeSolverState(C.Structure):
    _fields_ = [("base_us", C.c_int64), ("last_report", C.c_uint32), ("inited", C.c_int32)]


class InsPVA(C.Structure):
    """NovAtel INSPVA record, the reference's layout (type_defs.h:39-58)."""
    _fields_ = [("message_id", C.c_uint16), ("week_number", C.c_uint16),
                ("milliseconds", C.c_uint32), ("week_number_pos", C.c_uint32), ("pad0", C.c_uint32),
                ("seconds_pos", C.c_double), ("LLH", C.c_double * 3), ("V", C.c_double * 3),
                ("Eulr", C.c_double * 3), ("ins_status", C.c_int32), ("pad1", C.c_int32)]


INS_DTYPE = np.dtype([("message_id", "<u2"), ("week_number", "<u2"), ("milliseconds", "<u4"),
                      ("week_number_pos", "<u4"), ("pad0", "<u4"), ("seconds_pos", "<f8"),
                      ("LLH", "<f8", (3,)), ("V", "<f8", (3,)), ("Eulr", "<f8", (3,)),
                      ("ins_status", "<i4"), ("pad1", "<i4")])
assert INS_DTYPE.itemsize == C.sizeof(InsPVA) == 104


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        build_oracle()
    L = C.CDLL(LIB)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
    u8p = C.POINTER(C.c_uint8)
    sig = {
        "vo_create": (vp, []),
        "vo_destroy": (None, [vp]),
        "vo_set_calibration": (None, [vp, dp, dp, dp, dp, dp, C.c_int, C.c_int]),
        "vo_set_laser_selection": (None, [vp, C.POINTER(i32)]),
        "vo_set_points_skip": (None, [vp, i32]),
        "vo_set_crop": (None, [vp, i32, i32, dp]),
        "vo_clear_poses": (None, [vp]),
        "vo_add_pose": (None, [vp, i64, dp, dp, dp]),
        "vo_num_poses": (i32, [vp]),
        "vo_interpolate": (i32, [vp, i64, dp, dp]),
        "vo_pose_matrix": (None, [dp, dp]),
        "vo_unload": (None, [vp]),
        "vo_set_firing_skip": (None, [vp, i32]),
        "vo_get_state": (None, [vp, C.POINTER(i32)]),
        "vo_process_packet": (None, [vp, u8p, C.c_uint32, i64]),
        "vo_process_packets": (None, [vp, u8p, i64, i64, C.POINTER(i64)]),
        "vo_consume_packets": (None, [vp, u8p, i64, i64, C.POINTER(i64), C.POINTER(i64)]),
        "vo_split_frame": (None, [vp]),
        "vo_num_frames": (i32, [vp]),
        "vo_clear_frames": (None, [vp]),
        "vo_frame_get_info": (i32, [vp, i32, C.POINTER(FrameInfo)]),
        "vo_frame_laser_counts": (i32, [vp, i32, C.POINTER(i32)]),
        "vo_frame_points": (i32, [vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_uint16),
                                  C.POINTER(C.c_float)]),
        "vo_open_frame_points": (i64, [vp]),
        "vo_trace_enable": (None, [vp, i32]),
        "vo_trace_size": (i64, [vp]),
        "vo_trace_fetch": (None, [vp, C.POINTER(i32), u8p, u8p, u8p, C.POINTER(i32),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                  u8p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16),
                                  C.POINTER(C.c_uint32)]),
        "vo_read_frame_information": (i32, [u8p, i64, i64, C.POINTER(i64), C.POINTER(i32),
                                            C.POINTER(i32), C.POINTER(i64), i32]),
        "vo_get_frame": (i32, [vp, u8p, i64, i64, C.POINTER(i64), i64, i32]),
        "vo_llh2xyz": (None, [dp, dp]),
        "vo_xyz2llh": (None, [dp, dp]),
        "vo_llh2enu": (None, [dp, dp, dp]),
        "vo_ts_init": (None, [C.POINTER(TimeSolverState)]),
        "vo_ts_hdl": (i64, [C.POINTER(TimeSolverState), C.c_uint32, i64]),
        "vo_ts_ins": (i64, [C.POINTER(InsPVA), i64]),
        "vo_ins_pose": (None, [C.POINTER(InsPVA), dp, dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class OracleFrame:
    """One closed frame as the reference's HDLFrame holds it (laser-major)."""

    def __init__(self, info, counts, xyzi, azimuth, distance):
        self.timestamp_us = info.timestamp_us
        self.skips = info.skips
        self.n_lasers = info.n_lasers
        self.n_points = info.n_points
        self.n_packets = info.n_packets
        self.is_hdl64_order = bool(info.is_hdl64_order)
        self.carpose_TRV = np.array(list(info.carpose_TRV), dtype=np.float64)
        self.carpose_valid = info.carpose_seconds_pos != -1
        self.laser_counts = counts
        self.xyzi = xyzi
        self.azimuth = azimuth
        self.distance = distance


class Oracle:
    """The reference's HDLParser + TransformManager, restated on the CPU."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.vo_create())

    def __del__(self):
        try:
            if self._h:
                self._L.vo_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- configuration ------------------------------------------------------------
    def set_calibration(self, calib):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in
             (calib.rot_deg, calib.vert_deg, calib.dist_cm, calib.voff_cm, calib.hoff_cm)]
        self._L.vo_set_calibration(self._h, *[_p(x, C.c_double) for x in a], calib.n_rows,
                                   calib.n_enabled)

    def set_laser_selection(self, sel):
        s = np.ascontiguousarray(sel, dtype=np.int32)
        assert s.shape == (64,)
        self._L.vo_set_laser_selection(self._h, _p(s, C.c_int32))

    def set_points_skip(self, n):
        self._L.vo_set_points_skip(self._h, int(n))

    def set_crop(self, crop_returns, crop_inside, region):
        r = np.ascontiguousarray(region, dtype=np.float64)
        self._L.vo_set_crop(self._h, int(crop_returns), int(crop_inside), _p(r, C.c_double))

    # -- pose timeline ------------------------------------------------------------
    def clear_poses(self):
        self._L.vo_clear_poses(self._h)

    def add_poses(self, t_us, trv):
        trv = np.ascontiguousarray(trv, dtype=np.float64).reshape(-1, 9)
        for t, row in zip(np.asarray(t_us, dtype=np.int64), trv):
            T, R, V = (np.ascontiguousarray(row[0:3]), np.ascontiguousarray(row[3:6]),
                       np.ascontiguousarray(row[6:9]))
            self._L.vo_add_pose(self._h, int(t), _p(T, C.c_double), _p(R, C.c_double),
                                _p(V, C.c_double))

    def num_poses(self):
        return self._L.vo_num_poses(self._h)

    def interpolate(self, t_us):
        out = np.zeros(9, dtype=np.float64)
        sp = C.c_double(0)
        ok = self._L.vo_interpolate(self._h, int(t_us), _p(out, C.c_double), C.byref(sp))
        return bool(ok), out, sp.value

    @staticmethod
    def pose_matrix(trv):
        trv = np.ascontiguousarray(trv, dtype=np.float64)
        out = np.zeros(12, dtype=np.float64)
        lib().vo_pose_matrix(_p(trv, C.c_double), _p(out, C.c_double))
        return out.reshape(3, 4)

    # -- decode -------------------------------------------------------------------
    def unload(self):
        self._L.vo_unload(self._h)

    def set_firing_skip(self, s):
        self._L.vo_set_firing_skip(self._h, int(s))

    def state(self):
        s = np.zeros(4, dtype=np.int32)
        self._L.vo_get_state(self._h, _p(s, C.c_int32))
        return {"last_azimuth": int(s[0]), "firing_skip": int(s[1]),
                "frame_meta_inited": bool(s[2]), "is_hdl64": bool(s[3])}

    def trace_enable(self, on=True):
        self._L.vo_trace_enable(self._h, int(on))

    def process_packet(self, data, t_us, length=None):
        d = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        n = d.shape[0] if length is None else length
        self._L.vo_process_packet(self._h, _p(d, C.c_uint8), n, int(t_us))

    def process_packets(self, pkts_u8, t_us):
        d = np.ascontiguousarray(pkts_u8, dtype=np.uint8)
        assert d.ndim == 2 and d.shape[1] >= 1206
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        assert t.shape[0] == d.shape[0]
        self._L.vo_process_packets(self._h, _p(d, C.c_uint8), d.shape[0], d.shape[1],
                                   _p(t, C.c_int64))

    def consume_packets(self, pkts_u8, t_us):
        """processHDLPacket under the reference's consumer loop (HDLSource.cxx:209-225): closed
        frames are taken and the list cleared after every packet.  Returns (frames, points)."""
        d = np.ascontiguousarray(pkts_u8, dtype=np.uint8)
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        out = np.zeros(2, dtype=np.int64)
        self._L.vo_consume_packets(self._h, _p(d, C.c_uint8), d.shape[0], d.shape[1],
                                   _p(t, C.c_int64), _p(out, C.c_int64))
        return int(out[0]), int(out[1])

    def split_frame(self):
        self._L.vo_split_frame(self._h)

    def num_frames(self):
        return self._L.vo_num_frames(self._h)

    def clear_frames(self):
        self._L.vo_clear_frames(self._h)

    def open_frame_points(self):
        return int(self._L.vo_open_frame_points(self._h))

    def frame_summary(self, f):
        """(timestamp_us, skips, n_points, is_hdl64_order, per-row counts) of closed frame f
        without copying its points (long streams are checked frame list by frame list)."""
        info = FrameInfo()
        if not self._L.vo_frame_get_info(self._h, f, C.byref(info)):
            raise IndexError(f)
        counts = np.zeros(max(info.n_lasers, 1), dtype=np.int32)
        self._L.vo_frame_laser_counts(self._h, f, _p(counts, C.c_int32))
        return (int(info.timestamp_us), int(info.skips), int(info.n_points),
                bool(info.is_hdl64_order), counts[:info.n_lasers].copy())

    def frame(self, f):
        info = FrameInfo()
        if not self._L.vo_frame_get_info(self._h, f, C.byref(info)):
            raise IndexError(f)
        counts = np.zeros(max(info.n_lasers, 1), dtype=np.int32)
        self._L.vo_frame_laser_counts(self._h, f, _p(counts, C.c_int32))
        n = info.n_points
        xyzi = np.zeros((max(n, 1), 4), dtype=np.float32)
        az = np.zeros(max(n, 1), dtype=np.uint16)
        dist = np.zeros(max(n, 1), dtype=np.float32)
        self._L.vo_frame_points(self._h, f, _p(xyzi, C.c_float), _p(az, C.c_uint16),
                                _p(dist, C.c_float))
        return OracleFrame(info, counts[:info.n_lasers], xyzi[:n], az[:n], dist[:n])

    def frames(self):
        return [self.frame(i) for i in range(self.num_frames())]

    def trace(self):
        n = int(self._L.vo_trace_size(self._h))
        m = max(n, 1)
        out = {
            "packet": np.zeros(m, np.int32), "block": np.zeros(m, np.uint8),
            "dsr": np.zeros(m, np.uint8), "laser": np.zeros(m, np.uint8),
            "frame": np.zeros(m, np.int32), "x": np.zeros(m, np.float32),
            "y": np.zeros(m, np.float32), "z": np.zeros(m, np.float32),
            "intensity": np.zeros(m, np.uint8), "azimuth": np.zeros(m, np.uint16),
            "distance": np.zeros(m, np.uint16), "tadj_us": np.zeros(m, np.uint32),
        }
        self._L.vo_trace_fetch(
            self._h, _p(out["packet"], C.c_int32), _p(out["block"], C.c_uint8),
            _p(out["dsr"], C.c_uint8), _p(out["laser"], C.c_uint8), _p(out["frame"], C.c_int32),
            _p(out["x"], C.c_float), _p(out["y"], C.c_float), _p(out["z"], C.c_float),
            _p(out["intensity"], C.c_uint8), _p(out["azimuth"], C.c_uint16),
            _p(out["distance"], C.c_uint16), _p(out["tadj_us"], C.c_uint32))
        return {k: v[:n] for k, v in out.items()}

    # -- offline path ---------------------------------------------------------------
    @staticmethod
    def read_frame_information(pkts_u8, t_us):
        d = np.ascontiguousarray(pkts_u8, dtype=np.uint8)
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        cap = d.shape[0] * 12 + 1
        sp = np.zeros(cap, np.int32)
        sk = np.zeros(cap, np.int32)
        ts = np.zeros(cap, np.int64)
        n = lib().vo_read_frame_information(_p(d, C.c_uint8), d.shape[0], d.shape[1],
                                            _p(t, C.c_int64), _p(sp, C.c_int32),
                                            _p(sk, C.c_int32), _p(ts, C.c_int64), cap)
        return sp[:n].copy(), sk[:n].copy(), ts[:n].copy()

    def get_frame(self, pkts_u8, t_us, start_packet, skip, first=False):
        """HDLParser::getFrame.  The reference hands back frames.back() (HDLParser.cxx:531): when
        the packet that closes the frame holds several wraps that is the LAST frame closed by
        that packet, not the one asked for; first=True returns the first closed frame instead
        (identical whenever a packet holds at most one wrap, i.e. for real sensor data)."""
        d = np.ascontiguousarray(pkts_u8, dtype=np.uint8)
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        ok = self._L.vo_get_frame(self._h, _p(d, C.c_uint8), d.shape[0], d.shape[1],
                                  _p(t, C.c_int64), int(start_packet), int(skip))
        if not ok:
            return None
        return self.frame(0 if first else self.num_frames() - 1)


# ---- geodesy, TimeSolver, INS record -> pose (module-level: no parser state) ----
def llh2enu(llh, orgxyz):
    a = (C.c_double * 3)(*llh)
    o = (C.c_double * 3)(*orgxyz)
    e = (C.c_double * 3)()
    lib().vo_llh2enu(a, o, e)
    return np.array(e[:])


def llh2xyz(llh):
    a = (C.c_double * 3)(*llh)
    x = (C.c_double * 3)()
    lib().vo_llh2xyz(a, x)
    return np.array(x[:])


def xyz2llh(xyz):
    a = (C.c_double * 3)(*xyz)
    x = (C.c_double * 3)()
    lib().vo_xyz2llh(a, x)
    return np.array(x[:])


class TimeSolver:
    """TimeSolver::calcTimestamp(uint32_t) restated with the clock as an argument."""

    def __init__(self):
        self.state = TimeSolverState()
        lib().vo_ts_init(C.byref(self.state))

    def hdl(self, gps, now_us):
        return int(lib().vo_ts_hdl(C.byref(self.state), int(gps), int(now_us)))

    def hdl_many(self, gps, now_us):
        return np.array([self.hdl(g, now_us) for g in gps], dtype=np.int64)


def ins_times(recs, now_us):
    """TimeSolver::calcTimestamp(InsPVA const*) per record; now_us: scalar or per record."""
    recs = np.ascontiguousarray(recs, dtype=INS_DTYPE)
    now = np.broadcast_to(np.asarray(now_us, np.int64), (len(recs),))
    base = recs.ctypes.data
    return np.array([lib().vo_ts_ins(C.cast(base + i * INS_DTYPE.itemsize, C.POINTER(InsPVA)),
                                     int(now[i])) for i in range(len(recs))], dtype=np.int64)


def ins_poses(recs, orgxyz):
    """calcTransform (INSSource.cxx:300-326) per record -> n x 9 (T ENU, R, V)."""
    recs = np.ascontiguousarray(recs, dtype=INS_DTYPE)
    o = (C.c_double * 3)(*orgxyz)
    out = np.zeros((len(recs), 9))
    base = recs.ctypes.data
    for i in range(len(recs)):
        lib().vo_ins_pose(C.cast(base + i * INS_DTYPE.itemsize, C.POINTER(InsPVA)), o,
                          out[i].ctypes.data_as(C.POINTER(C.c_double)))
    return out
