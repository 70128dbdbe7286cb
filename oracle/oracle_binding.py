"""ctypes binding of oracle/liblins_oracle.so — TEST INFRASTRUCTURE (CPU oracle), not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY UNPINNED: see oracle/lins_oracle.hpp.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_DIR)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_defs = importlib.import_module("lins---lidar-inertial-slam_b200.ctypes_defs")
LinsParams, LinsReport, LinsBatchDesc = _defs.LinsParams, _defs.LinsReport, _defs.LinsBatchDesc
as_points, ptr, SCAN_RESULT_DTYPE = _defs.as_points, _defs.ptr, _defs.SCAN_RESULT_DTYPE

LIB_PATH = os.path.join(_DIR, "liblins_oracle.so")
FORM_A, FORM_B = 0, 1
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.lins_oracle_create.restype = vp
        L.lins_oracle_create.argtypes = [C.POINTER(LinsParams), C.c_int]
        L.lins_oracle_destroy.argtypes = [vp]
        L.lins_oracle_set_map.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.lins_oracle_ieskf.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.POINTER(LinsReport)]
        L.lins_oracle_ieskf_trace.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp, vp, C.POINTER(LinsReport), vp, vp, vp, vp, vp]
        L.lins_oracle_associate.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int] + [vp] * 8
        L.lins_oracle_estimate_transform.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.lins_oracle_update_map.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.POINTER(C.c_int)]
        L.lins_oracle_nn.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.lins_oracle_ieskf_batch.argtypes = [C.POINTER(LinsParams), C.POINTER(LinsBatchDesc), C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, vp, vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.lins_oracle_sym_eig6.argtypes = [vp, vp, vp]
        L.lins_oracle_qr_solve6.argtypes = [vp, vp, vp]
        L.lins_oracle_gain_form_a.argtypes = [vp, C.c_int, vp, C.c_double, vp]
        L.lins_oracle_boxplus.argtypes = [vp, vp, vp]
        L.lins_oracle_boxminus.argtypes = [vp, vp, vp]
        L.lins_oracle_transform.argtypes = [C.POINTER(LinsParams), vp, C.c_int, vp, C.c_int, vp]
        L.lins_oracle_measurement_rows.argtypes = [C.POINTER(LinsParams), vp, vp, vp, C.c_int, vp, vp]
        # row F2 (mapping refinement)
        L.lins_map_oracle_create.restype = vp
        L.lins_map_oracle_destroy.argtypes = [vp]
        L.lins_map_oracle_set_map.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.lins_map_oracle_associate.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp] + [vp] * 6
        L.lins_map_oracle_scan2map.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.POINTER(_defs.LinsMapReport)]
        L.lins_map_oracle_lm_solve.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.POINTER(C.c_int)]
        L.lins_map_oracle_eigen.argtypes = [vp, C.c_int, vp, vp]
        L.lins_map_oracle_qr_solve.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.lins_map_oracle_lu_invert.argtypes = [vp, C.c_int, vp]
        L.lins_map_oracle_gemm.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


class Oracle:
    def __init__(self, params=None, use_kdtree=True):
        self.L = lib()
        self.params = params or LinsParams.shipped()
        self.h = self.L.lins_oracle_create(C.byref(self.params), int(use_kdtree))

    def close(self):
        if getattr(self, "h", None):
            self.L.lins_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, surf_less_flat, corner_less_sharp):
        s, c = as_points(surf_less_flat), as_points(corner_less_sharp)
        self.L.lins_oracle_set_map(self.h, ptr(s), len(s), ptr(c), len(c))

    def ieskf(self, surf_flat, corner_sharp, state, cov, form=FORM_B):
        s, c = as_points(surf_flat), as_points(corner_sharp)
        st = np.ascontiguousarray(state, dtype=np.float64).reshape(19)
        cv = np.ascontiguousarray(cov, dtype=np.float64).reshape(324)
        so, co, rep = np.zeros(19), np.zeros(324), LinsReport()
        self.L.lins_oracle_ieskf(self.h, ptr(s), len(s), ptr(c), len(c), ptr(st), ptr(cv), form, ptr(so), ptr(co), C.byref(rep))
        return so, co, rep

    def ieskf_trace(self, surf_flat, corner_sharp, state, cov, form=FORM_B):
        s, c = as_points(surf_flat), as_points(corner_sharp)
        ns, nc, K = len(s), len(c), self.params.num_iter
        st = np.ascontiguousarray(state, dtype=np.float64).reshape(19)
        cv = np.ascontiguousarray(cov, dtype=np.float64).reshape(324)
        so, co, rep = np.zeros(19), np.zeros(324), LinsReport()
        tr = dict(surf_ind=np.full((K, ns, 3), -2, np.int32), corner_ind=np.full((K, nc, 2), -2, np.int32),
                  surf_mask=np.zeros((K, ns), np.uint8), corner_mask=np.zeros((K, nc), np.uint8), lin_state=np.zeros((K, 19)))
        self.L.lins_oracle_ieskf_trace(self.h, ptr(s), ns, ptr(c), nc, ptr(st), ptr(cv), form, ptr(so), ptr(co), C.byref(rep),
                                       ptr(tr["surf_ind"]), ptr(tr["corner_ind"]), ptr(tr["surf_mask"]), ptr(tr["corner_mask"]),
                                       ptr(tr["lin_state"]))
        for k in tr:
            tr[k] = tr[k][: rep.iters]
        return so, co, rep, tr

    def associate(self, surf_flat, corner_sharp, lin_state, it):
        s, c = as_points(surf_flat), as_points(corner_sharp)
        ns, nc = len(s), len(c)
        st = np.ascontiguousarray(lin_state, dtype=np.float64).reshape(19)
        out = dict(
            surf_ind=np.full((ns, 3), -2, np.int32), corner_ind=np.full((nc, 2), -2, np.int32),
            surf_coeff=np.zeros((ns, 4), np.float32), corner_coeff=np.zeros((nc, 4), np.float32),
            surf_mask=np.zeros(ns, np.uint8), corner_mask=np.zeros(nc, np.uint8),
            surf_sel=np.zeros((ns, 3), np.float32), corner_sel=np.zeros((nc, 3), np.float32),
        )
        self.L.lins_oracle_associate(self.h, ptr(s), ns, ptr(c), nc, ptr(st), int(it), ptr(out["surf_ind"]), ptr(out["corner_ind"]),
                                     ptr(out["surf_coeff"]), ptr(out["corner_coeff"]), ptr(out["surf_mask"]), ptr(out["corner_mask"]),
                                     ptr(out["surf_sel"]), ptr(out["corner_sel"]))
        return out

    def estimate_transform(self, surf_flat, corner_sharp, t, q_xyzw):
        s, c = as_points(surf_flat), as_points(corner_sharp)
        pose = np.ascontiguousarray(np.concatenate([np.asarray(t, float), np.asarray(q_xyzw, float)]))
        it, cv = C.c_int(0), C.c_int(0)
        self.L.lins_oracle_estimate_transform(self.h, ptr(s), len(s), ptr(c), len(c), ptr(pose), C.byref(it), C.byref(cv))
        return pose[:3].copy(), pose[3:].copy(), it.value, bool(cv.value)

    def update_map(self, surf_less_flat, corner_less_sharp, lin_state):
        s, c = as_points(surf_less_flat).copy(), as_points(corner_less_sharp).copy()
        st = np.ascontiguousarray(lin_state, dtype=np.float64).reshape(19)
        rep = C.c_int(0)
        self.L.lins_oracle_update_map(self.h, ptr(s), len(s), ptr(c), len(c), ptr(st), C.byref(rep))
        return s, c, bool(rep.value)

    def nn(self, which, xyz, use_kdtree):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        idx, sq = np.zeros(len(xyz), np.int32), np.zeros(len(xyz), np.float32)
        self.L.lins_oracle_nn(self.h, which, ptr(xyz), len(xyz), int(use_kdtree), ptr(idx), ptr(sq))
        return idx, sq


def ieskf_batch(params, batch, first=0, count=None, form=FORM_B, use_kdtree=True, threads=1, want_cov=True):
    """Scan-parallel CPU run of the oracle.  Returns (state, cov, results, seconds, total_iters)."""
    L = lib()
    count = batch.n - first if count is None else count
    d = batch.desc()
    so = np.zeros((count, 19))
    co = np.zeros((count, 324)) if want_cov else None
    res = np.zeros(count, dtype=SCAN_RESULT_DTYPE)
    sec, its = C.c_double(0), C.c_int64(0)
    L.lins_oracle_ieskf_batch(C.byref(params), C.byref(d), first, count, form, int(use_kdtree), threads, ptr(so), ptr(co),
                              ptr(res), C.byref(sec), C.byref(its))
    return so, co, res, sec.value, its.value


# ---- row F2: mapping-node scan-to-map refinement (oracle/lins_map_oracle.hpp) ------------------------------------------
class MapOracle:
    """CPU restatement of scan2MapOptimization (lidar_mapping_node.cpp:1635-1652)."""

    def __init__(self):
        self._h = lib().lins_map_oracle_create()

    def close(self):
        if self._h:
            lib().lins_map_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, corner_map, surf_map):
        c, s = as_points(corner_map), as_points(surf_map)
        lib().lins_map_oracle_set_map(self._h, ptr(c), len(c), ptr(s), len(s))

    def associate(self, corner_last, surf_last, transform):
        c, s = as_points(corner_last), as_points(surf_last)
        t = np.ascontiguousarray(transform, np.float32)
        out = dict(corner_knn=np.zeros((len(c), 5), np.int32), surf_knn=np.zeros((len(s), 5), np.int32),
                   corner_coeff=np.zeros((len(c), 4), np.float32), surf_coeff=np.zeros((len(s), 4), np.float32),
                   corner_mask=np.zeros(len(c), np.uint8), surf_mask=np.zeros(len(s), np.uint8))
        lib().lins_map_oracle_associate(self._h, ptr(c), len(c), ptr(s), len(s), ptr(t), ptr(out["corner_knn"]), ptr(out["surf_knn"]),
                                        ptr(out["corner_coeff"]), ptr(out["surf_coeff"]), ptr(out["corner_mask"]), ptr(out["surf_mask"]))
        return out

    def scan2map(self, corner_last, surf_last, transform):
        c, s = as_points(corner_last), as_points(surf_last)
        t = np.array(transform, np.float32).copy()
        rep = _defs.LinsMapReport()
        lib().lins_map_oracle_scan2map(self._h, ptr(c), len(c), ptr(s), len(s), ptr(t), C.byref(rep))
        return t, rep

    def lm_solve(self, AtA, AtB, it, transform):
        a, b = np.ascontiguousarray(AtA, np.float32), np.ascontiguousarray(AtB, np.float32)
        t = np.array(transform, np.float32).copy()
        x = np.zeros(6, np.float32)
        deg = C.c_int(0)
        conv = lib().lins_map_oracle_lm_solve(self._h, ptr(a), ptr(b), it, ptr(t), ptr(x), C.byref(deg))
        return t, x, bool(conv), bool(deg.value)


def cv_eigen(A):
    A = np.ascontiguousarray(A, np.float32); n = A.shape[0]
    W, V = np.zeros(n, np.float32), np.zeros((n, n), np.float32)
    lib().lins_map_oracle_eigen(ptr(A), n, ptr(W), ptr(V))
    return W, V


def cv_qr_solve(A, b):
    A = np.ascontiguousarray(A, np.float32); b = np.ascontiguousarray(b, np.float32).reshape(-1)
    x = np.zeros(A.shape[1], np.float32)
    ok = lib().lins_map_oracle_qr_solve(ptr(A), A.shape[0], A.shape[1], ptr(b), ptr(x))
    return bool(ok), x


def cv_lu_invert(A):
    A = np.ascontiguousarray(A, np.float32); n = A.shape[0]
    Ai = np.zeros((n, n), np.float32)
    ok = lib().lins_map_oracle_lu_invert(ptr(A), n, ptr(Ai))
    return bool(ok), Ai


def cv_gemm(A, B):
    A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
    Cm = np.zeros((A.shape[0], B.shape[1]), np.float32)
    lib().lins_map_oracle_gemm(ptr(A), ptr(B), ptr(Cm), A.shape[0], A.shape[1], B.shape[1])
    return Cm
