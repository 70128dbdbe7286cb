"""Golden vectors for row F2 (scan-to-map refinement), generated with the REAL OpenCV.

An independent Python restatement of lins/src/lidar_mapping_node.cpp:1351-1652 in which every OpenCV call the
reference makes is made for real — cv2.eigen, cv2.solve(DECOMP_QR), cv2.transpose / cv2.gemm for matAtA and matAtB,
cv2.invert — and everything else is scalar numpy with the reference's float / double promotions.  The 5-NN is a
brute-force exact search (f32 ((dx*dx)+dy*dy)+dz*dz, stable argsort = lowest index among ties).

Run in the build container (cv2 importable):   python tests/golden/make_map_golden.py
writes tests/golden/map_unit.npz (inputs + outputs); the GPU box never needs cv2 or /root/reference.
"""
import importlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
f32, f64 = np.float32, np.float64


def xyz(cloud):
    return np.stack([cloud["x"], cloud["y"], cloud["z"]], 1).astype(f32)


def associate_to_map(p, T):
    cR, sR, cP, sP, cY, sY = (f32(np.cos(T[0])), f32(np.sin(T[0])), f32(np.cos(T[1])), f32(np.sin(T[1])), f32(np.cos(T[2])), f32(np.sin(T[2])))
    x1 = cY * p[:, 0] - sY * p[:, 1]
    y1 = sY * p[:, 0] + cY * p[:, 1]
    z1 = p[:, 2]
    y2 = cR * y1 - sR * z1
    z2 = sR * y1 + cR * z1
    return np.stack([cP * x1 + sP * z2 + T[3], y2 + T[4], -sP * x1 + cP * z2 + T[5]], 1).astype(f32)


def knn5(mp, q):
    d = q[None, :] - mp
    dist = ((d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    order = np.argsort(dist, kind="stable")[:5]
    return order.astype(np.int32), dist[order]


def corner_fit(mp, sel, ind, dist):
    coeff = np.zeros(4, f32)
    if not dist[4] < 1.0:
        return coeff, False
    P = mp[ind]
    c = f32(0) * P[0]
    for j in range(5):
        c = c + P[j]
    c = c / f32(5)
    a = np.zeros(6, f32)
    for j in range(5):
        ax, ay, az = P[j] - c
        a = a + np.array([ax * ax, ax * ay, ax * az, ay * ay, ay * az, az * az], f32)
    a = a / f32(5)
    A = np.array([[a[0], a[1], a[2]], [a[1], a[3], a[4]], [a[2], a[4], a[5]]], f32)
    _, D, V = cv2.eigen(A)
    D = D.ravel()
    if D[0] > f32(3) * D[1]:
        x0, y0, z0 = sel
        x1, y1, z1 = (f32(f64(c[i]) + 0.1 * f64(V[0, i])) for i in range(3))
        x2, y2, z2 = (f32(f64(c[i]) - 0.1 * f64(V[0, i])) for i in range(3))
        m11 = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)
        m12 = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)
        m13 = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)
        a012 = np.sqrt(m11 * m11 + m12 * m12 + m13 * m13)
        l12 = np.sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2))
        la = ((y1 - y2) * m11 + (z1 - z2) * m12) / a012 / l12
        lb = -((x1 - x2) * m11 - (z1 - z2) * m13) / a012 / l12
        lc = -((x1 - x2) * m12 + (y1 - y2) * m13) / a012 / l12
        ld2 = a012 / l12
        s = f32(1 - 0.9 * f64(abs(ld2)))
        coeff = np.array([s * la, s * lb, s * lc, s * ld2], f32)
        return coeff, bool(f64(s) > 0.1)
    return coeff, False


def surf_fit(mp, sel, ind, dist):
    coeff = np.zeros(4, f32)
    if not dist[4] < 1.0:
        return coeff, False
    P = mp[ind]
    ok, X = cv2.solve(P.astype(f32), -np.ones((5, 1), f32), flags=cv2.DECOMP_QR)
    pa, pb, pc = X.ravel().astype(f32)
    pd = f32(1)
    with np.errstate(all="ignore"):
        ps = np.sqrt(pa * pa + pb * pb + pc * pc)
        pa, pb, pc, pd = pa / ps, pb / ps, pc / ps, pd / ps
        for j in range(5):
            if f64(abs(pa * P[j, 0] + pb * P[j, 1] + pc * P[j, 2] + pd)) > 0.2:
                return coeff, False
        pd2 = pa * sel[0] + pb * sel[1] + pc * sel[2] + pd
        s = f32(1 - 0.9 * f64(abs(pd2)) / f64(np.sqrt(np.sqrt(sel[0] * sel[0] + sel[1] * sel[1] + sel[2] * sel[2]))))
        coeff = np.array([s * pa, s * pb, s * pc, s * pd2], f32)
        return coeff, bool(f64(s) > 0.1)


def one_pass(unit_xyz, T):
    out = {}
    ori, coeffs = [], []
    for name, q, mp, fit in (("corner", unit_xyz["corner_last"], unit_xyz["corner_map"], corner_fit),
                             ("surf", unit_xyz["surf_last"], unit_xyz["surf_map"], surf_fit)):
        sel = associate_to_map(q, T)
        knn = np.zeros((len(q), 5), np.int32); co = np.zeros((len(q), 4), f32); mask = np.zeros(len(q), np.uint8)
        for i in range(len(q)):
            ind, dist = knn5(mp, sel[i])
            c, ok = fit(mp, sel[i], ind, dist)
            knn[i], co[i], mask[i] = ind, c, ok
            if ok:
                ori.append(q[i]); coeffs.append(c)
        out[name + "_knn"], out[name + "_coeff"], out[name + "_mask"] = knn, co, mask
    return out, np.array(ori, f32).reshape(-1, 3), np.array(coeffs, f32).reshape(-1, 4)


def lm(ori, co, it, T, state):
    srx, crx, sry, cry, srz, crz = (f32(np.sin(T[0])), f32(np.cos(T[0])), f32(np.sin(T[1])), f32(np.cos(T[1])), f32(np.sin(T[2])), f32(np.cos(T[2])))
    n = len(ori)
    if n < 50:
        return False, 0.0, 0.0
    x, y, z = ori[:, 0], ori[:, 1], ori[:, 2]
    cx, cy, cz = co[:, 0], co[:, 1], co[:, 2]
    arx = (crx * sry * srz * x + crx * crz * sry * y - srx * sry * z) * cx + (-srx * srz * x - crz * srx * y - crx * z) * cy + \
          (crx * cry * srz * x + crx * cry * crz * y - cry * srx * z) * cz
    ary = ((cry * srx * srz - crz * sry) * x + (sry * srz + cry * crz * srx) * y + crx * cry * z) * cx + \
          ((-cry * crz - srx * sry * srz) * x + (cry * srz - crz * srx * sry) * y - crx * sry * z) * cz
    arz = ((crz * srx * sry - cry * srz) * x + (-cry * crz - srx * sry * srz) * y) * cx + (crx * crz * x - crx * srz * y) * cy + \
          ((sry * srz + cry * crz * srx) * x + (crz * sry - cry * srx * srz) * y) * cz
    A = np.stack([arx, ary, arz, cx, cy, cz], 1).astype(f32)
    B = (-co[:, 3:4]).astype(f32)
    At = cv2.transpose(A)
    AtA = cv2.gemm(At, A, 1.0, None, 0.0)
    AtB = cv2.gemm(At, B, 1.0, None, 0.0)
    _, X = cv2.solve(AtA, AtB, flags=cv2.DECOMP_QR)
    if it == 0:
        _, E, V = cv2.eigen(AtA)
        E = E.ravel()
        V2 = V.copy()
        state["deg"] = False
        for i in range(5, -1, -1):
            if E[i] < 100:
                V2[i, :] = 0; state["deg"] = True
            else:
                break
        state["P"] = cv2.gemm(cv2.invert(V, flags=cv2.DECOMP_LU)[1], V2, 1.0, None, 0.0)
    if state["deg"]:
        X = cv2.gemm(state["P"], X, 1.0, None, 0.0)
    X = X.ravel().astype(f32)
    T += X
    r = lambda a: f64(f32(a * f32(57.29578)))  # noqa: E731
    dR = f32(np.sqrt(r(X[0]) ** 2 + r(X[1]) ** 2 + r(X[2]) ** 2))
    dT = f32(np.sqrt(f64(X[3] * f32(100)) ** 2 + f64(X[4] * f32(100)) ** 2 + f64(X[5] * f32(100)) ** 2))
    state.setdefault("AtA", []).append(AtA.copy()); state.setdefault("AtB", []).append(AtB.ravel().copy())
    return bool(dR < 0.05 and dT < 0.05), float(dR), float(dT)


def scan2map(unit_xyz, T0):
    T = np.array(T0, f32).copy()
    state, rep = {}, dict(iters=0, converged=0, n_sel=[], delta_r=[], delta_t=[])
    first = None
    for it in range(10):
        out, ori, co = one_pass(unit_xyz, T)
        if first is None:
            first = out
        rep["iters"] = it + 1; rep["n_sel"].append(len(ori))
        conv, dR, dT = lm(ori, co, it, T, state)
        rep["delta_r"].append(dR); rep["delta_t"].append(dT)
        if conv:
            rep["converged"] = 1
            break
    rep["degenerate"] = int(state.get("deg", False))
    return T, rep, first, state


def main():
    synth = importlib.import_module("lins---lidar-inertial-slam_b200.synth")
    u = synth.generate_map_unit("config3", seed=21, n_keyframes=5, sigma_t=0.08, sigma_r=0.008)
    ux = dict(corner_map=xyz(u.corner_map), surf_map=xyz(u.surf_map), corner_last=xyz(u.corner_last), surf_last=xyz(u.surf_last))
    T, rep, first, state = scan2map(ux, u.guess)
    print("map", len(u.corner_map), len(u.surf_map), "queries", len(u.corner_last), len(u.surf_last))
    print("report", rep, "\nerr before", np.abs(u.guess - u.truth), "\nerr after ", np.abs(T - u.truth))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "map_unit.npz"),
                        corner_map=u.corner_map, surf_map=u.surf_map, corner_last=u.corner_last, surf_last=u.surf_last,
                        truth=u.truth, guess=u.guess, T_out=T, iters=rep["iters"], converged=rep["converged"], degenerate=rep["degenerate"],
                        n_sel=np.array(rep["n_sel"], np.int32), delta_r=np.array(rep["delta_r"], f32), delta_t=np.array(rep["delta_t"], f32),
                        AtA=np.array(state["AtA"], f32), AtB=np.array(state["AtB"], f32), **{"it0_" + k: v for k, v in first.items()})


if __name__ == "__main__":
    main()
