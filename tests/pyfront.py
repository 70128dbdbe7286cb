"""Independent Python restatement of the LINS CPU front end — TEST INFRASTRUCTURE.

Written from the reference source, not from csrc/host/: image projection
(lins/src/image_projection_node.cpp:191-415: findStartEndAngle, projectPointCloud, groundRemoval, cloudSegmentation,
labelComponents) and the four-stage feature extraction (lins/include/StateEstimator.hpp:619-827: undistortPcl,
calculateSmoothness, markOccludedPoints, extractFeatures) with the pcl::VoxelGrid<PointXYZI> (leaf 0.2 m) it calls.
tests/test_frontend_cpu.py runs it next to the product's C++ restatements on simulated sweeps.

Float semantics follow the C++ overloads the reference resolves to: `float` locals, atan2 / sqrt of floats evaluated in
float, `* 180 / M_PI` in double and stored back to float.  VLP-16 constants: lins/include/parameters.h:82-92.
"""
import ctypes
import ctypes.util
import math

import numpy as np

F = np.float32
FLT_MAX = np.finfo(np.float32).max


class Lidar:
    def __init__(self, line_num=16, scan_num=1800, ang_res_x=0.2, ang_res_y=2.0, ang_bottom=15.0 + 0.1, ground_scan_ind=5, scan_period=0.1):
        self.line_num, self.scan_num, self.ground_scan_ind, self.scan_period = line_num, scan_num, ground_scan_ind, scan_period
        self.ang_res_x, self.ang_res_y, self.ang_bottom = F(ang_res_x), F(ang_res_y), F(ang_bottom)


# glibc's float functions through ctypes: numpy's float32 ufuncs may run SIMD variants that differ from libm by an ulp
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("atan2f", "sinf", "cosf"):
    getattr(_libm, _n).restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.sinf.argtypes = [ctypes.c_float]
_libm.cosf.argtypes = [ctypes.c_float]


def _atan2f(y, x):
    return F(_libm.atan2f(float(F(y)), float(F(x))))


def _sinf(a):
    return F(_libm.sinf(float(F(a))))


def _cosf(a):
    return F(_libm.cosf(float(F(a))))


def image_projection(xyz, lm=None):
    """xyz: (n, 3) float32 raw points in firing order.  Returns dict(seg (m, 4) float32 [x, y, z, intensity], outlier (k, 4),
    start_ring, end_ring, ori = (start, end, diff) float32, ground (m,) uint8, col (m,) uint32, range (m,) float32)."""
    lm = lm or Lidar()
    L, S = lm.line_num, lm.scan_num
    xyz = np.asarray(xyz, F)
    n = len(xyz)
    # findStartEndAngle (:191-203) — note the [size - 2].x slip; cloud_info fields are float32
    start = F(-_atan2f(xyz[0, 1], xyz[0, 0]))  # atan2(float, float): the float overload, then negated
    end = F(float(-_atan2f(xyz[n - 1, 1], xyz[n - 2, 0])) + 2 * math.pi)
    if float(end) - float(start) > 3 * math.pi:
        end = F(float(end) - 2 * math.pi)
    elif float(end) - float(start) < math.pi:
        end = F(float(end) + 2 * math.pi)
    diff = F(end - start)
    # projectPointCloud (:205-243)
    rangeMat = np.full((L, S), FLT_MAX, F)
    full = np.full((L * S, 4), np.nan, F)
    full[:, 3] = -1.0  # nanPoint.intensity
    for i in range(n):
        x, y, z = xyz[i]
        vert = F(float(_atan2f(z, np.sqrt(F(F(x * x) + F(y * y))))) * 180 / math.pi)
        rowf = F(F(vert + lm.ang_bottom) / lm.ang_res_y)
        if not (rowf >= 0) or rowf >= L:  # size_t rowIdn: negative values wrap to huge ones and are skipped
            continue
        row = int(rowf)
        hor = F(float(_atan2f(x, y)) * 180 / math.pi)
        # C's round() rounds half away from zero (numpy rounds half to even)
        v = (float(hor) - 90.0) / float(lm.ang_res_x)
        cold = -(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1)) + S // 2
        if cold < 0:
            continue
        col = int(cold)
        if col >= S:
            col -= S
        if col < 0 or col >= S:
            continue
        rng = np.sqrt(F(F(F(x * x) + F(y * y)) + F(z * z)))
        rangeMat[row, col] = rng
        inten = F(float(F(row)) + float(F(col)) / 10000.0)
        full[col + row * S] = (x, y, z, inten)
    # groundRemoval (:245-291)
    groundMat = np.zeros((L, S), np.int8)
    for j in range(S):
        for i in range(lm.ground_scan_ind):
            lo, up = j + i * S, j + (i + 1) * S
            if full[lo, 3] == -1 or full[up, 3] == -1:
                groundMat[i, j] = -1
                continue
            dx, dy, dz = F(full[up, 0] - full[lo, 0]), F(full[up, 1] - full[lo, 1]), F(full[up, 2] - full[lo, 2])
            ang = F(float(_atan2f(dz, np.sqrt(F(F(dx * dx) + F(dy * dy))))) * 180 / math.pi)
            if abs(float(ang) - 0.0) <= 10:
                groundMat[i, j] = 1
                groundMat[i + 1, j] = 1
    labelMat = np.zeros((L, S), np.int32)
    labelMat[(groundMat == 1) | (rangeMat == FLT_MAX)] = -1
    # cloudSegmentation (:293-339) with labelComponents (:341-413)
    alphaX, alphaY = F(float(lm.ang_res_x) / 180.0 * math.pi), F(float(lm.ang_res_y) / 180.0 * math.pi)
    theta = F(1.0472)
    # std::pair<uint8_t, uint8_t> neighbours (:71, :128-140): (-1, 0) and (0, -1) are stored as 255
    neigh = [(255, 0), (0, 1), (0, 255), (1, 0)]
    label_count = 1
    for r0 in range(L):
        for c0 in range(S):
            if labelMat[r0, c0] != 0:
                continue
            queue = [(r0, c0)]
            pushed = [(r0, c0)]
            line_flag = [False] * L
            qs = 0
            while qs < len(queue):
                fr, fc = queue[qs]
                qs += 1
                labelMat[fr, fc] = label_count
                for dr, dc in neigh:
                    tr, tc = fr + dr, fc + dc
                    if tr < 0 or tr >= L:
                        continue
                    if tc < 0:
                        tc = S - 1
                    if tc >= S:
                        tc = 0
                    if labelMat[tr, tc] != 0:
                        continue
                    d1 = max(rangeMat[fr, fc], rangeMat[tr, tc])
                    d2 = min(rangeMat[fr, fc], rangeMat[tr, tc])
                    alpha = alphaX if dr == 0 else alphaY
                    # float d2 * sin(float) etc.: sin / cos / atan2 of floats resolve to the float overloads
                    ang = _atan2f(F(d2 * _sinf(alpha)), F(d1 - F(d2 * _cosf(alpha))))
                    if ang > theta:
                        queue.append((tr, tc))
                        labelMat[tr, tc] = label_count
                        line_flag[tr] = True
                        pushed.append((tr, tc))
            feasible = len(pushed) >= 30
            if not feasible and len(pushed) >= 5:
                feasible = sum(line_flag) >= 3
            if feasible:
                label_count += 1
            else:
                for r, c in pushed:
                    labelMat[r, c] = 999999
    seg, outlier, ground, colind, rng_out = [], [], [], [], []
    start_ring, end_ring = np.zeros(L, np.int32), np.zeros(L, np.int32)
    for i in range(L):
        start_ring[i] = len(seg) - 1 + 5
        for j in range(S):
            if labelMat[i, j] > 0 or groundMat[i, j] == 1:
                if labelMat[i, j] == 999999:
                    if i > lm.ground_scan_ind and j % 5 == 0:
                        outlier.append(full[j + i * S].copy())
                    continue
                if groundMat[i, j] == 1 and (j % 5 != 0 and j > 5 and j < S - 5):
                    continue
                ground.append(1 if groundMat[i, j] == 1 else 0)
                colind.append(j)
                rng_out.append(rangeMat[i, j])
                seg.append(full[j + i * S].copy())
        end_ring[i] = len(seg) - 1 - 5
    A = lambda v, t, shape: np.asarray(v, t).reshape(shape)  # noqa: E731
    return dict(seg=A(seg, F, (-1, 4)), outlier=A(outlier, F, (-1, 4)), start_ring=start_ring, end_ring=end_ring, ori=(start, end, diff),
                ground=A(ground, np.uint8, -1), col=A(colind, np.uint32, -1), range=A(rng_out, F, -1))


def voxel_grid(pts, leaf=0.2):
    """pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.7 / 1.8 filters/impl/voxel_grid.hpp): voxel index from
    floor(p * inverse_leaf) - min_b, points sorted by voxel index, one centroid of ALL fields (f32 accumulation in input
    order) per voxel, output in ascending voxel index."""
    pts = np.asarray(pts, F).reshape(-1, 4)
    fin = np.isfinite(pts[:, :3]).all(1)
    if not fin.any():
        return np.zeros((0, 4), F)
    inv = F(1.0) / F(leaf)
    p = pts[fin]
    mn, mx = p[:, :3].min(0), p[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    mul = np.array([1, div[0], div[0] * div[1]], np.int64)
    ijk = np.floor(p[:, :3] * inv).astype(np.int64) - min_b
    idx = (ijk * mul).sum(1)
    order = np.argsort(idx, kind="stable")
    out = []
    k = 0
    while k < len(order):
        j = k
        acc = np.zeros(4, F)
        while j < len(order) and idx[order[j]] == idx[order[k]]:
            acc = (acc + p[order[j]]).astype(F)
            j += 1
        out.append((acc / F(j - k)).astype(F))
        k = j
    return np.asarray(out, F).reshape(-1, 4)


def extract_features(seg, info, lm=None, edge_threshold=0.5, surf_threshold=0.5):
    """StateEstimator.hpp:619-827 on the segmented cloud + cloud_info of image_projection()."""
    lm = lm or Lidar()
    seg = np.asarray(seg, F)
    n = len(seg)
    start_o, end_o, diff_o = (float(v) for v in info["ori"])
    cap = max(lm.line_num * lm.scan_num, n + 16)

    def padded(a, t):  # the message arrays are LINE_NUM * SCAN_NUM long (zeros behind the segmented points)
        out = np.zeros(cap, t)
        out[: len(a)] = a
        return out

    rng, colind, ground = padded(info["range"], F), padded(info["col"], np.int64), padded(info["ground"], np.uint8)
    # undistortPcl (:619-654); IMU_LIDAR_EXTRINSIC_ANGLE = 0: rotatePoint is the identity up to a f64 round trip
    und = seg.copy()
    half = False
    for i in range(n):
        ori = -float(_atan2f(seg[i, 1], seg[i, 0]))  # point.y, point.x are floats: atan2(float, float) -> float, stored in a double
        if not half:
            if ori < start_o - math.pi / 2:
                ori += 2 * math.pi
            elif ori > start_o + math.pi * 3 / 2:
                ori -= 2 * math.pi
            if ori - start_o > math.pi:
                half = True
        else:
            ori += 2 * math.pi
            if ori < end_o - math.pi * 3 / 2:
                ori += 2 * math.pi
            elif ori > end_o + math.pi / 2:
                ori -= 2 * math.pi
        rel = (ori - start_o) / diff_o
        und[i, 3] = F(int(seg[i, 3]) + lm.scan_period * rel)
    curv = np.zeros(cap)
    picked = np.zeros(cap, np.int64)
    label = np.zeros(cap, np.int64)
    smooth_val = np.zeros(cap)
    smooth_ind = np.zeros(cap, np.int64)
    r = rng.astype(np.float32)
    for i in range(5, n - 5):  # calculateSmoothness (:656-678): float sums (left to right), squared into a double
        d = F(r[i - 5] + r[i - 4])
        d = F(d + r[i - 3]); d = F(d + r[i - 2]); d = F(d + r[i - 1]); d = F(d - F(r[i] * F(10)))
        d = F(d + r[i + 1]); d = F(d + r[i + 2]); d = F(d + r[i + 3]); d = F(d + r[i + 4]); d = F(d + r[i + 5])
        curv[i] = float(d) * float(d)
        picked[i] = 0
        label[i] = 0
        smooth_val[i] = curv[i]
        smooth_ind[i] = i
    for i in range(5, n - 6):  # markOccludedPoints (:680-713)
        d1, d2 = r[i], r[i + 1]
        cd = abs(int(colind[i + 1]) - int(colind[i]))
        if cd < 10:
            if float(F(d1 - d2)) > 0.3:
                picked[i - 5 : i + 1] = 1
            elif float(F(d2 - d1)) > 0.3:
                picked[i + 1 : i + 7] = 1
        f1, f2 = abs(F(r[i - 1] - r[i])), abs(F(r[i + 1] - r[i]))
        if float(f1) > 0.02 * float(r[i]) and float(f2) > 0.02 * float(r[i]):
            picked[i] = 1
    sharp, less_sharp, flat, less_flat = [], [], [], []
    ties = 0

    def suppress(ind):
        # (the reference indexes ind + l unchecked; with the never-written default entries — ind 0 — that reaches index -1:
        # undefined there, both restatements stop at the array ends)
        picked[ind] = 1
        for l in range(1, 6):
            if ind + l >= cap:
                break
            if abs(int(colind[ind + l]) - int(colind[ind + l - 1])) > 10:
                break
            picked[ind + l] = 1
        for l in range(-1, -6, -1):
            if ind + l < 0:
                break
            if abs(int(colind[ind + l]) - int(colind[ind + l + 1])) > 10:
                break
            picked[ind + l] = 1

    for i in range(lm.line_num):  # extractFeatures (:719-827)
        ring_less = []
        for j in range(6):
            sp = (int(info["start_ring"][i]) * (6 - j) + int(info["end_ring"][i]) * j) // 6
            ep = (int(info["start_ring"][i]) * (5 - j) + int(info["end_ring"][i]) * (j + 1)) // 6 - 1
            if sp >= ep:
                continue
            seg_vals = smooth_val[sp:ep]
            ties += len(seg_vals) - len(np.unique(seg_vals))
            o = np.argsort(seg_vals, kind="stable")  # std::sort [sp, ep): element ep keeps its place (and is still visited below)
            smooth_val[sp:ep] = seg_vals[o]
            smooth_ind[sp:ep] = smooth_ind[sp:ep][o]
            largest = 0
            for k in range(ep, sp - 1, -1):
                ind = int(smooth_ind[k])
                if picked[ind] == 0 and curv[ind] > edge_threshold and ground[ind] == 0:
                    largest += 1
                    if largest <= 2:
                        label[ind] = 2
                        sharp.append(und[ind]); less_sharp.append(und[ind])
                    elif largest <= 20:
                        label[ind] = 1
                        less_sharp.append(und[ind])
                    else:
                        break
                    suppress(ind)
            smallest = 0
            for k in range(sp, ep + 1):
                ind = int(smooth_ind[k])
                if picked[ind] == 0 and curv[ind] < surf_threshold and ground[ind] == 1:
                    label[ind] = -1
                    flat.append(und[ind])
                    smallest += 1
                    if smallest >= 4:
                        break
                    suppress(ind)
            for k in range(sp, ep + 1):
                if label[k] <= 0:
                    ring_less.append(und[k])
        ds = voxel_grid(np.asarray(ring_less, F).reshape(-1, 4))
        less_flat.extend(list(ds))
    A = lambda v: np.asarray(v, F).reshape(-1, 4)  # noqa: E731
    return dict(undist=und, sharp=A(sharp), less_sharp=A(less_sharp), flat=A(flat), less_flat=A(less_flat), sort_ties=ties)
