"""Per-point deskew extension -- numpy statement of the semantics.

TEST INFRASTRUCTURE ONLY (only tests/ may import this).  PARITY UNPINNED: the reference does not
implement per-point motion compensation (it applies ONE pose per packet, Euler-angle lerp, and
re-bases only the translation: SURVEY.md F1/F2/F3), so there is nothing in /root/reference to
pin this against.  It states, in float64 numpy, what VS_FLAG_DESKEW_PER_POINT computes; the GPU
path is checked against it and against the reference-pinned per-packet path where the two must
coincide (tests/test_gpu_deskew.py).

Semantics
  * point time  tau = t_packet + off(block, dsr): the offsets that already define the t_us column
    (HDL-32 / VLP-16: the reference's own firing table, HDLParser.cxx:133-137, 946-962; HDL-64:
    zero unless the caller supplied the sensor-manual table with vs_set_firing_offsets);
  * pose bracket (a, b) = the one the reference picks for the PACKET time
    (clamp(lower_bound(t_packet), 1, N-1), TimeLine.h:384-468); r = (tau - t_a) / (t_b - t_a), not
    clamped (a packet that straddles a pose sample extrapolates its bracket);
  * rotation of a pose sample = PoseTransform::getMatrix's Ry(R0) Rx(R1) Rz(R2), degrees
    (type_defs.h:134-146), as a unit quaternion; q(tau) = slerp(q_a, q_b, r) along the shorter
    arc; T(tau) = T_a + (T_b - T_a) r;
  * frame origin (q_o, T_o) = the same pose function at the time of the frame's origin packet
    (the packet whose pose the reference subtracts, HDLParser.cxx:1004-1007, 1057-1062);
  * p' = R_o^T (R(tau) p + T(tau) - T_o): rotation AND translation re-based to the origin.
"""
import numpy as np


def euler_matrix(R_deg):
    r = np.asarray(R_deg, dtype=np.float64) * np.pi / 180.0
    c, s = np.cos(r), np.sin(r)
    ry = np.array([[c[0], 0, s[0]], [0, 1, 0], [-s[0], 0, c[0]]])
    rx = np.array([[1, 0, 0], [0, c[1], -s[1]], [0, s[1], c[1]]])
    rz = np.array([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]])
    return ry @ rx @ rz


def mat_to_quat(m):
    """Unit quaternion (w, x, y, z) of a rotation matrix, largest-pivot form."""
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
    return np.array(q)


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_rotate(q, p):
    """Rotate points p (n x 3) by unit quaternions q (4,) or (n x 4)."""
    q = np.asarray(q, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    qv = q[..., 1:]
    t = 2.0 * np.cross(qv, p)
    return p + q[..., :1] * t + np.cross(qv, t)


def slerp(qa, qb, r):
    """qa, qb: (4,); r: scalar or (n,).  Shorter arc, r not clamped."""
    d = float(np.dot(qa, qb))
    if d < 0:
        qb, d = -qb, -d
    theta = np.arccos(min(1.0, d))
    r = np.asarray(r, dtype=np.float64)
    if theta < 1e-8:
        w0, w1 = 1.0 - r, r
    else:
        w0 = np.sin((1.0 - r) * theta) / np.sin(theta)
        w1 = np.sin(r * theta) / np.sin(theta)
    return w0[..., None] * qa + w1[..., None] * qb if r.ndim else w0 * qa + w1 * qb


class PoseTimeline:
    def __init__(self, pose_t, pose_trv):
        self.t = np.asarray(pose_t, dtype=np.int64)
        self.trv = np.asarray(pose_trv, dtype=np.float64).reshape(-1, 9)
        self.q = np.array([mat_to_quat(euler_matrix(v[3:6])) for v in self.trv])

    def bracket(self, t_us):
        i = int(np.searchsorted(self.t, t_us, side="left"))
        return min(max(i, 1), len(self.t) - 1)

    def pose(self, i, tau):
        """(q, T) at times tau (array) inside/outside bracket (i-1, i)."""
        ta, tb = self.t[i - 1], self.t[i]
        r = (np.asarray(tau, dtype=np.float64) - float(ta)) / float(tb - ta)
        q = slerp(self.q[i - 1], self.q[i], r)
        Ta, Tb = self.trv[i - 1, :3], self.trv[i, :3]
        T = Ta + (Tb - Ta) * (r[..., None] if np.ndim(r) else r)
        return q, T


def deskew_points(xyz_sensor, pkt_of_point, off_us, pkt_time_us, origin_time_of_pkt, pose_t, pose_trv):
    """xyz_sensor: n x 3 sensor-frame points; pkt_of_point / off_us: per point; pkt_time_us and
    origin_time_of_pkt: per packet.  Returns n x 3 float64 in the frame-origin coordinates."""
    tl = PoseTimeline(pose_t, pose_trv)
    out = np.empty((len(xyz_sensor), 3))
    pkt_of_point = np.asarray(pkt_of_point)
    order = np.argsort(pkt_of_point, kind="stable")
    bounds = np.flatnonzero(np.diff(pkt_of_point[order])) + 1
    for idx in np.split(order, bounds):
        P = int(pkt_of_point[idx[0]])
        t_pkt = int(pkt_time_us[P])
        i = tl.bracket(t_pkt)
        tau = t_pkt + np.asarray(off_us)[idx].astype(np.float64)
        q, T = tl.pose(i, tau)
        t_o = int(origin_time_of_pkt[P])
        qo, To = tl.pose(tl.bracket(t_o), np.float64(t_o))
        world = quat_rotate(q, xyz_sensor[idx]) + T - To
        out[idx] = quat_rotate(quat_conj(qo), world)
    return out
