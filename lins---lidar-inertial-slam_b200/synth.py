"""ctypes binding of tools/synth/liblins_synth.so — the synthetic scan-pair generator (inputs only).

The generator runs the product's own host-side CPU stages (image projection, feature extraction, IMU
propagation; all stay on the CPU per BASELINE.json north_star) over a seeded ray-cast world and returns a
:class:`Batch` of independent (scan pair, prior) units in the C-ABI's ``lins_batch_desc`` layout.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .ctypes_defs import Batch, LinsBatchDesc, POINT_DTYPE

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DIR = os.path.join(_ROOT, "tools", "synth")
_LIB = None


class SynthCfg(C.Structure):
    _fields_ = [
        ("lidar", C.c_int32),
        ("world", C.c_int32),
        ("fixed_motion", C.c_int32),
        ("stress_queries", C.c_int32),
        ("v_max", C.c_double),
        ("w_max", C.c_double),
        ("range_noise", C.c_double),
        ("prior_vel_sigma", C.c_double),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_DIR, "liblins_synth.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.lins_synth_batch_create.restype = C.c_void_p
        L.lins_synth_batch_create.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_int, C.c_int]
        L.lins_synth_batch_destroy.argtypes = [C.c_void_p]
        L.lins_synth_batch_desc.argtypes = [C.c_void_p, C.POINTER(LinsBatchDesc)]
        L.lins_synth_batch_truth.restype = C.POINTER(C.c_double)
        L.lins_synth_batch_truth.argtypes = [C.c_void_p]
        L.lins_synth_batch_new_less.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _LIB = L
    return _LIB


# BASELINE.json configs -> generator settings
CONFIGS = {
    # configs[0] "single synthetic VLP-16 scan ... correctness gate" 
    "config1": dict(lidar=0, world=0, fixed_motion=1, stress_queries=0, v_max=10.0, w_max=0.3, range_noise=0.01,
                    prior_vel_sigma=0.05),
    # config 1b: the literal "~2k surf + 500 edge feats" as queries
    "config1b": dict(lidar=0, world=0, fixed_motion=1, stress_queries=1, v_max=10.0, w_max=0.3, range_noise=0.01,
                     prior_vel_sigma=0.05),
    # configs[2] "synthetic 1000-scan sequence, flat-ground map" / configs[4] (8000 scans)
    "config3": dict(lidar=0, world=1, fixed_motion=0, stress_queries=0, v_max=10.0, w_max=0.3, range_noise=0.01,
                    prior_vel_sigma=0.05),
    # configs[3] "synthetic 64-ring (64x1024) dense scan"
    "config4": dict(lidar=1, world=0, fixed_motion=0, stress_queries=0, v_max=10.0, w_max=0.3, range_noise=0.01,
                    prior_vel_sigma=0.05),
}


def _copy(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


def generate(config="config3", n=1, seed0=1, threads=None, **overrides):
    """n independent units, seeds seed0..seed0+n-1."""
    L = lib()
    kw = dict(CONFIGS[config])
    kw.update(overrides)
    cfg = SynthCfg(**kw)
    threads = threads or min(os.cpu_count() or 1, 32)
    h = L.lins_synth_batch_create(C.byref(cfg), seed0, n, threads)
    try:
        d = LinsBatchDesc()
        L.lins_synth_batch_desc(h, C.byref(d))
        clouds, offsets = {}, {}
        for k in Batch.FIELDS:
            off = _copy(getattr(d, k + "_off"), n + 1, np.int32)
            offsets[k] = off
            clouds[k] = _copy(getattr(d, k), int(off[-1]), POINT_DTYPE)
        state = _copy(d.state_in, n * 19, np.float64)
        cov = _copy(d.cov_in, n * 324, np.float64)
        truth = np.ctypeslib.as_array(L.lins_synth_batch_truth(h), shape=(n * 7,)).copy()
        extra = {}
        for which, name in ((0, "new_surf_less_flat"), (1, "new_corner_less_sharp")):
            p, o = C.c_void_p(), C.c_void_p()
            L.lins_synth_batch_new_less(h, which, C.byref(p), C.byref(o))
            off = _copy(o.value, n + 1, np.int32)
            extra[name] = _copy(p.value, int(off[-1]), POINT_DTYPE)
            extra[name + "_off"] = off
        return Batch(clouds, offsets, state, cov, truth, extra)
    finally:
        L.lins_synth_batch_destroy(h)
