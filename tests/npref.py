"""Independent numpy / scipy restatement of rows A5-A11 of SURVEY.md §8 — TEST INFRASTRUCTURE.

Written from the reference source (lins/include/StateEstimator.hpp:465-600, :917-951, :1031-1060;
lins/include/KalmanFilter.hpp:71-94; lins/include/math_utils.h:39-88, :304-321), NOT from oracle/: rotations come from
scipy.spatial.transform.Rotation, the gain from LAPACK's Cholesky (scipy.linalg.cho_factor / cho_solve on the M x M
innovation covariance, the reference's form A), everything else is plain numpy.  tests/test_oracle_pin_cpu.py runs
this next to the C++ oracle: the two share no code, so agreement pins the oracle's linear algebra to third-party
libraries (the reference's Eigen is not installable here).
State layout: rn(0:3) vn(3:6) q xyzw(6:10) ba(10:13) bw(13:16) gn(16:19); error state pos0 vel3 att6 acc9 gyr12 gra15.
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve
from scipy.spatial.transform import Rotation

POS, VEL, ATT, ACC, GYR, GRA = 0, 3, 6, 9, 12, 15


def skew(v):  # math_utils.h:197-204
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def wrap_pi(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def quat2axis(q_xyzw):  # math_utils.h:75-88 (rotation vector, angle wrapped to [-pi, pi))
    v = np.asarray(q_xyzw[:3], float)
    n = np.linalg.norm(v)
    if n < 1e-10:
        return v.copy()
    return v / n * wrap_pi(2.0 * np.arctan2(n, q_xyzw[3]))


def axis2quat(vec):  # math_utils.h:60-73 -> xyzw
    th = np.linalg.norm(vec)
    if th < 1e-10:
        return np.array([0.0, 0.0, 0.0, 1.0])
    return Rotation.from_rotvec(vec).as_quat()  # (sin(th/2) a, cos(th/2))


def rinvleft(axis):  # math_utils.h:304-321
    th = np.linalg.norm(axis)
    if th < 1e-10:
        return np.eye(3)
    a = axis / th
    h = th / 2.0
    s = h / np.tan(h)
    return s * np.eye(3) + (1.0 - s) * np.outer(a, a) - h * skew(a)


def qmul(a, b):  # xyzw Hamilton product (Eigen Quaterniond operator*)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def box_plus(state, dx):  # KalmanFilter.hpp:71-81
    out = np.array(state, float)
    out[0:3] += dx[POS:POS + 3]
    out[3:6] += dx[VEL:VEL + 3]
    out[10:13] += dx[ACC:ACC + 3]
    out[13:16] += dx[GYR:GYR + 3]
    out[16:19] += dx[GRA:GRA + 3]
    q = qmul(state[6:10], axis2quat(dx[ATT:ATT + 3]))
    out[6:10] = q / np.linalg.norm(q)
    return out


def box_minus(a, b):  # a (-) b, KalmanFilter.hpp:84-94
    dx = np.zeros(18)
    dx[POS:POS + 3] = a[0:3] - b[0:3]
    dx[VEL:VEL + 3] = a[3:6] - b[3:6]
    dx[ACC:ACC + 3] = a[10:13] - b[10:13]
    dx[GYR:GYR + 3] = a[13:16] - b[13:16]
    dx[GRA:GRA + 3] = a[16:19] - b[16:19]
    binv = np.array([-b[6], -b[7], -b[8], b[9]])
    dx[ATT:ATT + 3] = quat2axis(qmul(binv, a[6:10]))
    return dx


def plane_coeff(sel, t1, t2, t3, weighted):
    """StateEstimator.hpp:922-949.  sel, t*: float32 xyz.  Returns (accepted, coeff float32[4])."""
    P0, P1, P2, P3 = (np.asarray(v, np.float32).astype(np.float64) for v in (sel, t1, t2, t3))
    M = np.cross(P1 - P2, P1 - P3)
    r = float(np.dot(P0 - P1, M))
    m = float(np.linalg.norm(M))
    with np.errstate(all="ignore"):
        res = np.float32(r / m)
        jac = M / m
        s = np.float32(1.0)
        if weighted:
            f = np.asarray(sel, np.float32)
            rng2 = np.float32(np.float32(f[0] * f[0]) + np.float32(f[1] * f[1])) + np.float32(f[2] * f[2])
            s = np.float32(1.0 - 1.8 * float(np.abs(res)) / float(np.sqrt(np.sqrt(np.float32(rng2)))))
        ok = bool(s > np.float32(0.1)) and bool(res != 0)
    c = np.array([float(s) * jac[0], float(s) * jac[1], float(s) * jac[2], float(s) * float(res)]).astype(np.float32)
    return ok, c


def line_coeff(sel, t1, t2, weighted):
    """StateEstimator.hpp:1035-1058."""
    P0, P1, P2 = (np.asarray(v, np.float32).astype(np.float64) for v in (sel, t1, t2))
    P = np.cross(P0 - P1, P0 - P2)
    with np.errstate(all="ignore"):
        r = np.float32(np.linalg.norm(P))
        d12 = np.float32(np.linalg.norm(P1 - P2))
        res = np.float32(r / d12)
        jac = P @ skew(P2 - P1) / (float(d12) * float(r))
        s = np.float32(1.0)
        if weighted:
            s = np.float32(1.0 - 1.8 * float(np.abs(res)))
        ok = bool(s > np.float32(0.1)) and bool(res != 0)
    c = np.array([float(s) * jac[0], float(s) * jac[1], float(s) * jac[2], float(s) * float(res)]).astype(np.float32)
    return ok, c


def measurement(lin, keypoints, coeffs, lidar_scale):
    """StateEstimator.hpp:512-532: residual (M) and Hk (M x 18)."""
    M = len(keypoints)
    H = np.zeros((M, 18))
    r = lidar_scale * coeffs[:, 3].astype(np.float64)
    axis = quat2axis(lin[6:10])
    R = Rotation.from_quat(lin[6:10]).as_matrix()
    Ri = rinvleft(-axis)
    for i in range(M):
        c = coeffs[i, :3].astype(np.float64)
        H[i, ATT:ATT + 3] = c @ (-R @ skew(keypoints[i].astype(np.float64))) @ Ri
        H[i, POS:POS + 3] = c
    return r, H


def ieskf(prior, P0, iterations, lidar_std=0.01, lidar_scale=1.0, num_iter=30):
    """performIESKF (StateEstimator.hpp:465-600) with the associations supplied per iteration.

    iterations(k, lin_state) -> (keypoints (M,3) float32 = ORIGINAL query points, coeffs (M,4) float32), i.e. the output
    of findCorresponding* at linState_.  Returns (state, P, iters, converged, diverged).
    """
    Pk = np.array(P0, float)
    lin = np.array(prior, float)
    residual_norm = 1e6
    conv = div = False
    K = H = None
    M = 0
    it = 0
    while it < num_iter and not conv and not div:
        kp, cf = iterations(it, lin)
        M = len(kp)
        r, H = measurement(lin, kp, cf, lidar_scale)
        Rk = lidar_std ** 2 * np.eye(M)
        Py = H @ Pk @ H.T + Rk
        Pyinv = cho_solve(cho_factor(Py, lower=True), np.eye(M)) if M else np.zeros((0, 0))
        K = Pk @ H.T @ Pyinv
        d = box_minus(prior, lin)
        upd = -K @ (r + H @ d) + d
        it += 1
        if np.isnan(upd).any():
            div = True
            break
        if np.linalg.norm(r) > residual_norm * 10:
            div = True
            break
        lin = box_plus(lin, upd)
        if np.linalg.norm(upd) <= 1e-2:
            conv = True
        residual_norm = np.linalg.norm(r)
    if div:
        return np.array(prior, float), Pk, it, conv, div
    IKH = np.eye(18) - K @ H
    Pn = IKH @ Pk @ IKH.T + K @ (lidar_std ** 2 * np.eye(M)) @ K.T
    Pn = 0.5 * (Pn + Pn.T)
    return lin, Pn, it, conv, div
