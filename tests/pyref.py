"""Second, independent restatement (pure Python + numpy scalars, small cases only) of the association part of
the reference: transformToStart (lins/include/StateEstimator.hpp:1066-1080), the 1-NN + ring walks of
findCorrespondingSurfFeatures (:844-915) and findCorrespondingCornerFeatures (:970-1029).

TEST INFRASTRUCTURE: used to pin the C++ oracle on hand-checkable scenes; never imported by product code.
"""
import math

import numpy as np

F = np.float32


def quat2axis(q):  # q = (x, y, z, w); math_utils.h:75-88
    x, y, z, w = q
    mag = math.sqrt(x * x + y * y + z * z)
    v = np.array([x, y, z], float)
    if mag >= 1e-10:
        ang = 2.0 * math.atan2(mag, w)
        while ang >= math.pi:
            ang -= 2 * math.pi
        while ang < -math.pi:
            ang += 2 * math.pi
        v = v / mag * ang
    return v


def axis2quat(v):  # -> (w, x, y, z); math_utils.h:43-73
    th = math.sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]))
    if th < 1e-10:
        return (1.0, 0.0, 0.0, 0.0)
    a = v / th
    m = math.sin(th / 2.0)
    return (math.cos(th / 2.0), a[0] * m, a[1] * m, a[2] * m)


def qrot(q, v):  # Eigen _transformVector
    w, x, y, z = q
    qv = np.array([x, y, z])
    uv = np.cross(qv, v)
    uv = uv + uv
    return v + w * uv + np.cross(qv, uv)


def transform_to_start(p, intensity, rn, q_xyzw, scan_period=0.1):
    fi = F(intensity) - F(int(F(intensity)))
    s = (1.0 / scan_period) * float(fi)
    phi = quat2axis(q_xyzw)
    r = axis2quat(s * phi)
    P1 = qrot(r, np.asarray(p, float)) + s * np.asarray(rn, float)
    return np.array([F(P1[0]), F(P1[1]), F(P1[2])], dtype=np.float32)


def sqd(a, b):
    dx, dy, dz = F(a[0]) - F(b[0]), F(a[1]) - F(b[1]), F(a[2]) - F(b[2])
    return F(F(F(dx * dx) + F(dy * dy)) + F(dz * dz))


def nn1(sel, tgt):
    best, bi = F(np.inf), -1
    for j in range(len(tgt)):
        if not np.all(np.isfinite(tgt[j, :3])):
            continue
        d = sqd(sel, tgt[j])
        if d < best:
            best, bi = d, j
    return bi, best


def assoc_surf(sel, tgt, n_query, near=25.0):
    """tgt: (T,4) float32 x,y,z,intensity. Returns (i1,i2,i3)."""
    c, d = nn1(sel, tgt)
    i1 = i2 = i3 = -1
    if c >= 0 and float(d) < near:
        i1 = c
        cr = int(tgt[c, 3])
        m2 = m3 = F(near)
        for j in range(c + 1, min(n_query, len(tgt))):
            if int(tgt[j, 3]) > cr + 2.5:
                break
            dd = sqd(tgt[j], sel)
            if int(tgt[j, 3]) <= cr:
                if dd < m2:
                    m2, i2 = dd, j
            else:
                if dd < m3:
                    m3, i3 = dd, j
        for j in range(c - 1, -1, -1):
            if int(tgt[j, 3]) < cr - 2.5:
                break
            dd = sqd(tgt[j], sel)
            if int(tgt[j, 3]) >= cr:
                if dd < m2:
                    m2, i2 = dd, j
            else:
                if dd < m3:
                    m3, i3 = dd, j
    return i1, i2, i3


def assoc_corner(sel, tgt, n_query, near=25.0):
    c, d = nn1(sel, tgt)
    i1 = i2 = -1
    if c >= 0 and float(d) < near:
        i1 = c
        cr = int(tgt[c, 3])
        m2 = F(near)
        for j in range(c + 1, min(n_query, len(tgt))):
            if int(tgt[j, 3]) > cr + 2.5:
                break
            dd = sqd(tgt[j], sel)
            if int(tgt[j, 3]) > cr:
                if dd < m2:
                    m2, i2 = dd, j
        for j in range(c - 1, -1, -1):
            if int(tgt[j, 3]) < cr - 2.5:
                break
            dd = sqd(tgt[j], sel)
            if int(tgt[j, 3]) < cr:
                if dd < m2:
                    m2, i2 = dd, j
    return i1, i2
