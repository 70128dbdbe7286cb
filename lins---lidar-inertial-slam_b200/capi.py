"""ctypes binding of the C-ABI shared library (include/lins_gpu.h) — the call a Python user makes.

`LinsGpu` mirrors the seam of the reference's ``fusion::StateEstimator`` that the GPU path replaces
(lins/include/StateEstimator.hpp): ``set_map`` ≙ kdtree*->setInputCloud (:363-364, :1156-1160),
``ieskf`` ≙ performIESKF (:465-600), ``associate`` ≙ findCorrespondingSurf/CornerFeatures (:829-1063),
``ieskf_batch`` ≙ performIESKF over many independent (scan pair, prior) units.

There is NO CPU fallback: if the library is missing or no sm_100 device is present this raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .ctypes_defs import (Batch, COV_SIZE, LinsBatchDesc, LinsMapReport, LinsParams, LinsReport, LinsScanResult, POINT_DTYPE,
                          SCAN_RESULT_DTYPE, STATE_DIM, as_points, ptr)

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("LINS_GPU_LIB") or os.path.join(_PKG, "liblins_gpu.so")  # override only for A/B experiments
CUDA_DIR = os.path.join(_PKG, "csrc", "cuda")

# every symbol include/lins_gpu.h declares
EXPORTS = [
    "lins_gpu_abi_version", "lins_gpu_create", "lins_gpu_destroy", "lins_gpu_last_error", "lins_gpu_set_params",
    "lins_gpu_set_map", "lins_gpu_ieskf", "lins_gpu_associate", "lins_gpu_estimate_transform", "lins_gpu_update_map",
    "lins_gpu_batch_upload", "lins_gpu_batch_run", "lins_gpu_batch_download", "lins_gpu_ieskf_batch",
    "lins_gpu_batch_results_device", "lins_gpu_batch_jacobian_pass", "lins_gpu_launch_count", "lins_gpu_sync",
    "lins_gpu_debug_phase_cycles", "lins_gpu_map_set", "lins_gpu_scan2map", "lins_gpu_map_associate",
    "lins_gpu_host_register", "lins_gpu_host_unregister", "lins_gpu_batch_download_indices", "lins_gpu_update_map_ex",
    "lins_gpu_batch_upload_stats",
]

NVCC_COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
# translation units and their extra flags: lins_gpu.cu (C-ABI + the fused kernel: bit-exact association, no multiply-add
# contraction), lins_jacobian.cu (the tolerance-checked split Jacobian kernel: contraction allowed)
UNITS = [("lins_gpu.cu", ["-fmad=false"]), ("lins_jacobian.cu", [])]
NVCC_FLAGS = NVCC_COMMON + ["-fmad=false", "-shared"]  # (what tools/ scripts print)


def build(force=False, verbose=False, out=None, extra=()):
    """Compile csrc/cuda/*.cu -> liblins_gpu.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = out or LIB_PATH
    deps = [os.path.join(CUDA_DIR, u) for u, _ in UNITS] + [os.path.join(_ROOT, "include", "lins_gpu.h")]
    for d in (CUDA_DIR, os.path.join(os.path.dirname(CUDA_DIR), "host")):  # every header the translation units include
        deps += [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cuh", ".hpp", ".h"))]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in deps):
        return out
    objdir = os.path.join(_PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    tag = os.path.splitext(os.path.basename(out))[0]
    objs = []
    for unit, flags in UNITS:
        obj = os.path.join(objdir, f"{tag}_{os.path.splitext(unit)[0]}.o")
        subprocess.check_call(["nvcc"] + NVCC_COMMON + flags + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CUDA_DIR, unit)])
        objs.append(obj)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs)
    return out


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32p, f32p, u8p, f64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        L.lins_gpu_create.argtypes = [C.POINTER(LinsParams), C.c_int, vp, C.POINTER(vp)]
        L.lins_gpu_destroy.argtypes = [vp]
        L.lins_gpu_destroy.restype = None
        L.lins_gpu_last_error.argtypes = [vp]
        L.lins_gpu_last_error.restype = C.c_char_p
        L.lins_gpu_set_params.argtypes = [vp, C.POINTER(LinsParams)]
        L.lins_gpu_set_map.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.lins_gpu_ieskf.argtypes = [vp, vp, C.c_int, vp, C.c_int, f64p, f64p, f64p, f64p, C.POINTER(LinsReport)]
        L.lins_gpu_associate.argtypes = [vp, vp, C.c_int, vp, C.c_int, f64p, C.c_int, i32p, i32p, f32p, f32p, u8p, u8p, f32p, f32p]
        L.lins_gpu_estimate_transform.argtypes = [vp, vp, C.c_int, vp, C.c_int, f64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.lins_gpu_update_map.argtypes = [vp, vp, C.c_int, vp, C.c_int, f64p, C.POINTER(C.c_int)]
        if hasattr(L, "lins_gpu_update_map_ex"):  # (absent from older A/B variant libraries selected through LINS_GPU_LIB)
            L.lins_gpu_update_map_ex.argtypes = [vp, vp, C.c_int, vp, C.c_int, f64p, vp, vp, C.POINTER(C.c_int)]
        L.lins_gpu_batch_upload.argtypes = [vp, C.POINTER(LinsBatchDesc)]
        L.lins_gpu_batch_run.argtypes = [vp]
        L.lins_gpu_batch_download.argtypes = [vp, f64p, f64p, vp, vp]
        L.lins_gpu_ieskf_batch.argtypes = [vp, C.POINTER(LinsBatchDesc), f64p, f64p, vp]
        L.lins_gpu_batch_results_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int)]
        L.lins_gpu_batch_jacobian_pass.argtypes = [vp, f64p]
        L.lins_gpu_launch_count.argtypes = [vp]
        L.lins_gpu_launch_count.restype = C.c_int64
        L.lins_gpu_sync.argtypes = [vp]
        L.lins_gpu_debug_phase_cycles.argtypes = [vp, C.c_int, vp]
        L.lins_gpu_map_set.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.lins_gpu_scan2map.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.POINTER(LinsMapReport)]
        L.lins_gpu_map_associate.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp] + [vp] * 6
        L.lins_gpu_host_register.argtypes = [vp, C.c_size_t]
        L.lins_gpu_batch_download_indices.argtypes = [vp, vp, vp]
        if hasattr(L, "lins_gpu_batch_upload_stats"):
            L.lins_gpu_batch_upload_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.lins_gpu_host_unregister.argtypes = [vp]
        _LIB = L
    return _LIB


class LinsError(RuntimeError):
    pass


def pin_batch(batch):
    """Page-lock the four cloud arrays of a host batch in place (lins_gpu_host_register): lins_gpu_batch_upload then DMAs
    the raw records straight from them.  Returns the list of pinned arrays; call unpin_batch before they are freed."""
    L = lib()
    pinned = []
    for k in Batch.FIELDS:
        a = batch.clouds[k]
        if a.nbytes == 0:
            continue
        rc = L.lins_gpu_host_register(a.ctypes.data, a.nbytes)
        if rc != 0:
            unpin_arrays(pinned)
            raise LinsError(f"lins_gpu_host_register failed with {rc}")
        pinned.append(a)
    batch._pinned = pinned
    return pinned


def unpin_arrays(arrays):
    L = lib()
    for a in arrays:
        L.lins_gpu_host_unregister(a.ctypes.data)


def unpin_batch(batch):
    unpin_arrays(getattr(batch, "_pinned", []))
    batch._pinned = []


class LinsGpu:
    """One context = one CUDA device + one stream (pass ``stream=torch.cuda.current_stream().cuda_stream`` to
    share torch's stream so torch.cuda.Event timing sees the kernels)."""

    def __init__(self, params=None, device=0, stream=None):
        self.L = lib()
        self.params = params or LinsParams.shipped()
        h = C.c_void_p()
        rc = self.L.lins_gpu_create(C.byref(self.params), device, C.c_void_p(stream or 0), C.byref(h))
        if rc != 0:
            raise LinsError(f"lins_gpu_create failed with {rc} (no sm_100 CUDA device? there is no CPU fallback)")
        self.h = h
        self._batch_n = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.lins_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LinsError(f"error {rc}: {self.L.lins_gpu_last_error(self.h).decode()}")

    def set_params(self, params):
        self.params = params
        self._ck(self.L.lins_gpu_set_params(self.h, C.byref(params)))

    def launch_count(self):
        return int(self.L.lins_gpu_launch_count(self.h))

    def sync(self):
        self._ck(self.L.lins_gpu_sync(self.h))

    def phase_cycles(self, enable=True, read=False):
        out = np.zeros(64, np.int64) if read else None
        self._ck(self.L.lins_gpu_debug_phase_cycles(self.h, int(enable), ptr(out)))
        return out

    # ---- single-scan seam ---------------------------------------------------------------------------------
    def set_map(self, surf_less_flat, corner_less_sharp):
        s, c = as_points(surf_less_flat), as_points(corner_less_sharp)
        self._ck(self.L.lins_gpu_set_map(self.h, ptr(s), len(s), ptr(c), len(c)))

    def ieskf(self, surf_flat, corner_sharp, state, cov):
        s, c = as_points(surf_flat), as_points(corner_sharp)
        st = np.ascontiguousarray(state, dtype=np.float64).reshape(STATE_DIM)
        cv = np.ascontiguousarray(cov, dtype=np.float64).reshape(COV_SIZE)
        so, co, rep = np.zeros(STATE_DIM), np.zeros(COV_SIZE), LinsReport()
        self._ck(self.L.lins_gpu_ieskf(self.h, ptr(s), len(s), ptr(c), len(c), ptr(st), ptr(cv), ptr(so), ptr(co), C.byref(rep)))
        return so, co, rep

    def associate(self, surf_flat, corner_sharp, lin_state, it):
        s, c = as_points(surf_flat), as_points(corner_sharp)
        ns, nc = len(s), len(c)
        st = np.ascontiguousarray(lin_state, dtype=np.float64).reshape(STATE_DIM)
        out = dict(
            surf_ind=np.full((ns, 3), -2, np.int32), corner_ind=np.full((nc, 2), -2, np.int32),
            surf_coeff=np.zeros((ns, 4), np.float32), corner_coeff=np.zeros((nc, 4), np.float32),
            surf_mask=np.zeros(ns, np.uint8), corner_mask=np.zeros(nc, np.uint8),
            surf_sel=np.zeros((ns, 3), np.float32), corner_sel=np.zeros((nc, 3), np.float32),
        )
        self._ck(self.L.lins_gpu_associate(self.h, ptr(s), ns, ptr(c), nc, ptr(st), int(it), ptr(out["surf_ind"]),
                                           ptr(out["corner_ind"]), ptr(out["surf_coeff"]), ptr(out["corner_coeff"]),
                                           ptr(out["surf_mask"]), ptr(out["corner_mask"]), ptr(out["surf_sel"]),
                                           ptr(out["corner_sel"])))
        return out

    def estimate_transform(self, surf_flat, corner_sharp, t, q_xyzw):
        s, c = as_points(surf_flat), as_points(corner_sharp)
        pose = np.ascontiguousarray(np.concatenate([np.asarray(t, float), np.asarray(q_xyzw, float)]))
        it, cv = C.c_int(0), C.c_int(0)
        self._ck(self.L.lins_gpu_estimate_transform(self.h, ptr(s), len(s), ptr(c), len(c), ptr(pose), C.byref(it), C.byref(cv)))
        return pose[:3].copy(), pose[3:].copy(), it.value, bool(cv.value)

    def update_map(self, surf_less_flat, corner_less_sharp, lin_state):
        """In-place transformToEnd of the two clouds (returned) + conditional map refresh."""
        s, c = as_points(surf_less_flat).copy(), as_points(corner_less_sharp).copy()
        st = np.ascontiguousarray(lin_state, dtype=np.float64).reshape(STATE_DIM)
        rep = C.c_int(0)
        self._ck(self.L.lins_gpu_update_map(self.h, ptr(s), len(s), ptr(c), len(c), ptr(st), C.byref(rep)))
        return s, c, bool(rep.value)

    def update_map_device(self, surf_less_flat, corner_less_sharp, lin_state=None):
        """Map refresh that stays on the device: no read-back, no synchronisation; lin_state None = the posterior of the
        last ieskf() call, still resident."""
        s, c = as_points(surf_less_flat), as_points(corner_less_sharp)
        st = None if lin_state is None else np.ascontiguousarray(lin_state, dtype=np.float64).reshape(STATE_DIM)
        rep = C.c_int(0)
        self._ck(self.L.lins_gpu_update_map_ex(self.h, ptr(s), len(s), ptr(c), len(c), ptr(st), None, None, C.byref(rep)))
        return bool(rep.value)

    # ---- row F2: scan-to-map refinement of the mapping node ---------------------------------------------------------
    def map_set(self, corner_from_map, surf_from_map):
        c, s = as_points(corner_from_map), as_points(surf_from_map)
        self._ck(self.L.lins_gpu_map_set(self.h, ptr(c), len(c), ptr(s), len(s)))

    def scan2map(self, corner_last, surf_last, transform):
        """≙ the iteration loop of scan2MapOptimization; returns (transformTobeMapped, LinsMapReport)."""
        c, s = as_points(corner_last), as_points(surf_last)
        t = np.array(transform, np.float32).copy()
        rep = LinsMapReport()
        self._ck(self.L.lins_gpu_scan2map(self.h, ptr(c), len(c), ptr(s), len(s), ptr(t), C.byref(rep)))
        return t, rep

    def map_associate(self, corner_last, surf_last, transform):
        c, s = as_points(corner_last), as_points(surf_last)
        t = np.ascontiguousarray(transform, np.float32)
        out = dict(corner_knn=np.zeros((len(c), 5), np.int32), surf_knn=np.zeros((len(s), 5), np.int32),
                   corner_coeff=np.zeros((len(c), 4), np.float32), surf_coeff=np.zeros((len(s), 4), np.float32),
                   corner_mask=np.zeros(len(c), np.uint8), surf_mask=np.zeros(len(s), np.uint8))
        self._ck(self.L.lins_gpu_map_associate(self.h, ptr(c), len(c), ptr(s), len(s), ptr(t), ptr(out["corner_knn"]), ptr(out["surf_knn"]),
                                               ptr(out["corner_coeff"]), ptr(out["surf_coeff"]), ptr(out["corner_mask"]), ptr(out["surf_mask"])))
        return out

    # ---- batched mode ------------------------------------------------------------------------------------------
    def batch_upload(self, batch):
        d = batch.desc()
        self._keep = batch  # keep the host arrays alive while the async copies are in flight
        self._ck(self.L.lins_gpu_batch_upload(self.h, C.byref(d)))
        self._batch_n = batch.n

    def batch_upload_stats(self):
        """(points uploaded as host-packed 16-B records, points uploaded as raw 32-B records), cumulative."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.lins_gpu_batch_upload_stats(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def batch_run(self):
        self._ck(self.L.lins_gpu_batch_run(self.h))

    def batch_download(self, states=True, covs=True, reports=False):
        n = self._batch_n
        so = np.zeros((n, STATE_DIM)) if states else None
        co = np.zeros((n, COV_SIZE)) if covs else None
        res = np.zeros(n, dtype=SCAN_RESULT_DTYPE)
        reps = (LinsReport * n)() if reports else None
        self._ck(self.L.lins_gpu_batch_download(self.h, ptr(so), ptr(co), ptr(res), C.cast(reps, C.c_void_p) if reports else None))
        return so, co, res, reps

    def batch_download_indices(self, batch: Batch):
        """(surf_ind (Ns_total, 3), corner_ind (Nc_total, 2)) of the resident batch's last search iteration."""
        si = np.full((int(batch.offsets["surf_flat"][-1]), 3), -2, np.int32)
        ci = np.full((int(batch.offsets["corner_sharp"][-1]), 2), -2, np.int32)
        self._ck(self.L.lins_gpu_batch_download_indices(self.h, ptr(si), ptr(ci)))
        return si, ci

    def ieskf_batch(self, batch, covs=True):
        """upload + run + download through host buffers: the end-to-end entry point."""
        d = batch.desc()
        n = batch.n
        so = np.zeros((n, STATE_DIM))
        co = np.zeros((n, COV_SIZE)) if covs else None
        res = np.zeros(n, dtype=SCAN_RESULT_DTYPE)
        self._ck(self.L.lins_gpu_ieskf_batch(self.h, C.byref(d), ptr(so), ptr(co), ptr(res)))
        self._batch_n = n
        return so, co, res

    def batch_results_device(self):
        p, n = C.c_void_p(), C.c_int(0)
        self._ck(self.L.lins_gpu_batch_results_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def batch_jacobian_pass(self, want_accum=False):
        acc = np.zeros((self._batch_n, 32)) if want_accum else None
        self._ck(self.L.lins_gpu_batch_jacobian_pass(self.h, ptr(acc)))
        return acc
