"""Independent evaluator of the per-point deskew extension in 50-digit
arithmetic (mpmath) and in a DIFFERENT formulation from the CUDA path and from
oracle/deskew_port.py: no quaternions.  The pose rotation between two INS samples is the SO(3)
geodesic  R(r) = R_a exp(r log(R_a^T R_b))  (axis-angle of the relative rotation, Rodrigues'
formula), which is what slerp along the shorter arc of the unit quaternions traces; the rest is
the statement in include/veloslam_b200.h: T(r) = T_a + r (T_b - T_a), bracket of the PACKET time,
r = (tau - t_a) / (t_b - t_a) unclamped, p' = R_o^T (R(tau) p + T(tau) - T_o) with (R_o, T_o) the
same pose function at the frame origin's time.

Test infrastructure only.  Slow (pure Python, arbitrary precision): use on samples of points."""
import bisect

import mpmath as mp

mp.mp.dps = 50


def _euler_matrix(R_deg):
    """PoseTransform::getMatrix: Ry(R0) Rx(R1) Rz(R2), degrees (type_defs.h:134-146)."""
    a, b, c = [mp.mpf(float(v)) * mp.pi / 180 for v in R_deg]
    ry = mp.matrix([[mp.cos(a), 0, mp.sin(a)], [0, 1, 0], [-mp.sin(a), 0, mp.cos(a)]])
    rx = mp.matrix([[1, 0, 0], [0, mp.cos(b), -mp.sin(b)], [0, mp.sin(b), mp.cos(b)]])
    rz = mp.matrix([[mp.cos(c), -mp.sin(c), 0], [mp.sin(c), mp.cos(c), 0], [0, 0, 1]])
    return ry * rx * rz


def _geodesic(Ra, Rb, r):
    """R_a exp(r log(R_a^T R_b)) via the axis and angle of the relative rotation."""
    D = Ra.T * Rb
    c = (D[0, 0] + D[1, 1] + D[2, 2] - 1) / 2
    c = max(mp.mpf(-1), min(mp.mpf(1), c))
    phi = mp.acos(c)
    if phi < mp.mpf(10) ** -30:
        return Ra
    s = mp.sin(phi)
    ax = [(D[2, 1] - D[1, 2]) / (2 * s), (D[0, 2] - D[2, 0]) / (2 * s), (D[1, 0] - D[0, 1]) / (2 * s)]
    K = mp.matrix([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    th = r * phi
    return Ra * (mp.eye(3) + mp.sin(th) * K + (1 - mp.cos(th)) * (K * K))


class Timeline:
    def __init__(self, pose_t, pose_trv):
        self.t = [int(v) for v in pose_t]
        self.trv = [[float(x) for x in row] for row in pose_trv]
        self._R = {}

    def R(self, i):
        if i not in self._R:
            self._R[i] = _euler_matrix(self.trv[i][3:6])
        return self._R[i]

    def bracket(self, t_us):
        i = bisect.bisect_left(self.t, int(t_us))
        return min(max(i, 1), len(self.t) - 1)

    def pose(self, i, tau):
        ta, tb = self.t[i - 1], self.t[i]
        r = (mp.mpf(tau) - ta) / (tb - ta)
        R = _geodesic(self.R(i - 1), self.R(i), r)
        Ta = mp.matrix(self.trv[i - 1][:3])
        Tb = mp.matrix(self.trv[i][:3])
        return R, Ta + (Tb - Ta) * r


def deskew_point(tl, p_sensor, t_packet_us, off_us, t_origin_us):
    """One point: returns [x, y, z] as Python floats (rounded from 50 digits)."""
    i = tl.bracket(t_packet_us)
    R, T = tl.pose(i, int(t_packet_us) + int(off_us))
    Ro, To = tl.pose(tl.bracket(t_origin_us), int(t_origin_us))
    p = mp.matrix([mp.mpf(float(v)) for v in p_sensor])
    out = Ro.T * (R * p + T - To)
    return [float(out[k]) for k in range(3)]
