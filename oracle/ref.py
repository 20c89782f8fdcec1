"""ctypes binding of oracle/_ref/libvelo_ref.so: the REFERENCE's own HDLParser /
TransformManager / vtkPacketFileWriter compiled verbatim against oracle/ref_shim.

TEST INFRASTRUCTURE ONLY.  Same method names as oracle.oracle.Oracle so the two can be
compared output-for-output (tests/test_oracle_vs_ref.py) and either can serve as the CPU
reference arm of bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

from .build_ref import LIB, build_ref
from .oracle import FrameInfo, OracleFrame, _p

_lib = None


def available():
    """True when the reference library exists (built here, or shipped with the snapshot)."""
    try:
        return build_ref() is not None and os.path.exists(LIB)
    except Exception:
        return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref is not built (the reference sources are not present)")
    L = C.CDLL(LIB)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double)
    u8p = C.POINTER(C.c_uint8)
    sig = {
        "vr_create": (vp, []),
        "vr_destroy": (None, [vp]),
        "vr_set_corrections_file": (None, [vp, C.c_char_p]),
        "vr_num_channels": (i32, [vp]),
        "vr_set_laser_selection": (None, [vp, C.POINTER(i32)]),
        "vr_set_points_skip": (None, [vp, i32]),
        "vr_set_crop": (None, [vp, i32, i32, dp]),
        "vr_clear_poses": (None, [vp]),
        "vr_add_pose": (None, [vp, i64, dp, dp, dp]),
        "vr_num_poses": (i32, [vp]),
        "vr_interpolate": (i32, [vp, i64, dp, dp]),
        "vr_pose_matrix": (None, [dp, dp]),
        "vr_unload": (None, [vp]),
        "vr_get_state": (None, [vp, C.POINTER(i32)]),
        "vr_process_packets": (None, [vp, u8p, i64, i64, C.POINTER(i64), C.POINTER(i32)]),
        "vr_consume_packets": (None, [vp, u8p, i64, i64, C.POINTER(i64), C.POINTER(i64)]),
        "vr_split_frame": (None, [vp]),
        "vr_num_frames": (i32, [vp]),
        "vr_clear_frames": (None, [vp]),
        "vr_open_frame_points": (i64, [vp]),
        "vr_frame_get_info": (i32, [vp, i32, C.POINTER(FrameInfo)]),
        "vr_frame_laser_counts": (i32, [vp, i32, C.POINTER(i32)]),
        "vr_frame_points": (i32, [vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_uint16),
                                  C.POINTER(C.c_float)]),
        "vr_read_frame_information": (i32, [vp, C.c_char_p, C.POINTER(i64), C.POINTER(i32),
                                            C.POINTER(i64), i32]),
        "vr_get_frame": (i32, [vp, C.c_char_p, i64, i32]),
        "vr_num_snapshot_frames": (i32, [vp]),
        "vr_write_pcap": (i32, [C.c_char_p, u8p, i64, i64, C.POINTER(i64)]),
        "vr_set_fake_now": (None, [i64]),
        "vr_ts_create": (vp, []),
        "vr_ts_destroy": (None, [vp]),
        "vr_ts_hdl": (i64, [vp, C.c_uint32]),
        "vr_ts_ins": (i64, [vp, vp]),
        "vr_llh2enu": (None, [dp, dp, dp]),
        "vr_llh2xyz": (None, [dp, dp]),
        "vr_sizeof_inspva": (i32, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class RefParser:
    """The reference's HDLParser + TransformManager (real code), driven like the oracle."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p(self._L.vr_create())
        self._tmp = []

    def __del__(self):
        try:
            if self._h:
                self._L.vr_destroy(self._h)
                self._h = None
            for t in self._tmp:
                os.unlink(t)
        except Exception:
            pass

    def set_calibration(self, calib):
        """Goes through the reference's own XML loader (HDLParser::setCorrectionsFile)."""
        from veloslam_b200.calibxml import write_db_xml
        f = tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False)
        f.close()
        write_db_xml(f.name, calib)
        self._tmp.append(f.name)
        self._L.vr_set_corrections_file(self._h, f.name.encode())

    def num_channels(self):
        return self._L.vr_num_channels(self._h)

    def set_laser_selection(self, sel):
        s = np.ascontiguousarray(sel, dtype=np.int32)
        self._L.vr_set_laser_selection(self._h, _p(s, C.c_int32))

    def set_points_skip(self, n):
        self._L.vr_set_points_skip(self._h, int(n))

    def set_crop(self, crop_returns, crop_inside, region):
        r = np.ascontiguousarray(region, dtype=np.float64)
        self._L.vr_set_crop(self._h, int(crop_returns), int(crop_inside), _p(r, C.c_double))

    def clear_poses(self):
        self._L.vr_clear_poses(self._h)

    def add_poses(self, t_us, trv):
        trv = np.ascontiguousarray(trv, dtype=np.float64).reshape(-1, 9)
        for t, row in zip(np.asarray(t_us, dtype=np.int64), trv):
            T, R, V = (np.ascontiguousarray(row[0:3]), np.ascontiguousarray(row[3:6]),
                       np.ascontiguousarray(row[6:9]))
            self._L.vr_add_pose(self._h, int(t), _p(T, C.c_double), _p(R, C.c_double),
                                _p(V, C.c_double))

    def num_poses(self):
        return self._L.vr_num_poses(self._h)

    def interpolate(self, t_us):
        out = np.zeros(9, dtype=np.float64)
        sp = C.c_double(0)
        ok = self._L.vr_interpolate(self._h, int(t_us), _p(out, C.c_double), C.byref(sp))
        return bool(ok), out, sp.value

    @staticmethod
    def pose_matrix(trv):
        trv = np.ascontiguousarray(trv, dtype=np.float64)
        out = np.zeros(12, dtype=np.float64)
        lib().vr_pose_matrix(_p(trv, C.c_double), _p(out, C.c_double))
        return out.reshape(3, 4)

    def unload(self):
        self._L.vr_unload(self._h)

    def state(self):
        s = np.zeros(4, dtype=np.int32)
        self._L.vr_get_state(self._h, _p(s, C.c_int32))
        return {"last_azimuth": int(s[0]), "firing_skip": int(s[1]),
                "frame_meta_inited": bool(s[2]), "is_hdl64": bool(s[3])}

    def process_packets(self, pkts_u8, t_us, lengths=None):
        d = np.ascontiguousarray(pkts_u8, dtype=np.uint8)
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        ln = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        self._L.vr_process_packets(self._h, _p(d, C.c_uint8), d.shape[0], d.shape[1],
                                   _p(t, C.c_int64), None if ln is None else _p(ln, C.c_int32))

    def consume_packets(self, pkts_u8, t_us):
        """The reference's consumer loop (HDLSource.cxx:209-225): after every packet, frames that
        closed are taken (getAllFrames().back()) and the parser's list is cleared.  Returns
        (frames taken, points in them)."""
        d = np.ascontiguousarray(pkts_u8, dtype=np.uint8)
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        out = np.zeros(2, dtype=np.int64)
        self._L.vr_consume_packets(self._h, _p(d, C.c_uint8), d.shape[0], d.shape[1],
                                   _p(t, C.c_int64), _p(out, C.c_int64))
        return int(out[0]), int(out[1])

    def split_frame(self):
        self._L.vr_split_frame(self._h)

    def num_frames(self):
        return self._L.vr_num_frames(self._h)

    def clear_frames(self):
        self._L.vr_clear_frames(self._h)

    def open_frame_points(self):
        return int(self._L.vr_open_frame_points(self._h))

    def _frame(self, f):
        info = FrameInfo()
        if not self._L.vr_frame_get_info(self._h, f, C.byref(info)):
            raise IndexError(f)
        counts = np.zeros(max(info.n_lasers, 1), dtype=np.int32)
        self._L.vr_frame_laser_counts(self._h, f, _p(counts, C.c_int32))
        n = info.n_points
        xyzi = np.zeros((max(n, 1), 4), dtype=np.float32)
        az = np.zeros(max(n, 1), dtype=np.uint16)
        dist = np.zeros(max(n, 1), dtype=np.float32)
        self._L.vr_frame_points(self._h, f, _p(xyzi, C.c_float), _p(az, C.c_uint16),
                                _p(dist, C.c_float))
        return OracleFrame(info, counts[:info.n_lasers], xyzi[:n], az[:n], dist[:n])

    def frames(self):
        return [self._frame(i) for i in range(self.num_frames())]

    def read_frame_information(self, pcap_path):
        cap = 1 << 16
        pos = np.zeros(cap, np.int64)
        sk = np.zeros(cap, np.int32)
        ts = np.zeros(cap, np.int64)
        n = self._L.vr_read_frame_information(self._h, pcap_path.encode(), _p(pos, C.c_int64),
                                              _p(sk, C.c_int32), _p(ts, C.c_int64), cap)
        return pos[:n].copy(), sk[:n].copy(), ts[:n].copy()

    def get_frame(self, pcap_path, file_pos, skip):
        ok = self._L.vr_get_frame(self._h, pcap_path.encode(), int(file_pos), int(skip))
        return self._frame(0) if ok else None

    @staticmethod
    def write_pcap(path, pkts_u8, t_us):
        d = np.ascontiguousarray(pkts_u8, dtype=np.uint8)
        t = np.ascontiguousarray(t_us, dtype=np.int64)
        return bool(lib().vr_write_pcap(path.encode(), _p(d, C.c_uint8), d.shape[0], d.shape[1],
                                        _p(t, C.c_int64)))


# ---- the reference's TimeSolver.cxx / CoordiTran.cpp behind a settable clock ----
def set_fake_now(us):
    """Every clock the reference reads (microsec_clock / day_clock in the shim) returns this."""
    lib().vr_set_fake_now(int(us))


class RefTimeSolver:
    def __init__(self, now_us):
        set_fake_now(now_us)          # the constructor reads the clock too
        self._h = lib().vr_ts_create()

    def __del__(self):
        try:
            lib().vr_ts_destroy(self._h)
        except Exception:
            pass

    def hdl(self, gps, now_us):
        set_fake_now(now_us)
        return int(lib().vr_ts_hdl(self._h, int(gps)))

    def ins(self, rec_addr, now_us):
        set_fake_now(now_us)
        return int(lib().vr_ts_ins(self._h, rec_addr))


def llh2enu(llh, orgxyz):
    a = (C.c_double * 3)(*llh)
    o = (C.c_double * 3)(*orgxyz)
    e = (C.c_double * 3)()
    lib().vr_llh2enu(a, o, e)
    return np.array(e[:])


def llh2xyz(llh):
    a = (C.c_double * 3)(*llh)
    x = (C.c_double * 3)()
    lib().vr_llh2xyz(a, x)
    return np.array(x[:])


def sizeof_inspva():
    return int(lib().vr_sizeof_inspva())
