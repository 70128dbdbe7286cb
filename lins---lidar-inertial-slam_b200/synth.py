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


# ---- offline sequence driver (tools/synth/lins_sequence.cpp): the C++ StateEstimator shim on a synthetic drive ----
_SEQ = None


def seq_lib():
    global _SEQ
    if _SEQ is None:
        path = os.path.join(_DIR, "liblins_seq.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _DIR, "seq"])
        L = C.CDLL(path)
        L.lins_seq_run.restype = C.c_void_p
        L.lins_seq_run.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_int, C.c_int]
        L.lins_seq_destroy.argtypes = [C.c_void_p]
        L.lins_seq_num_units.argtypes = [C.c_void_p]
        L.lins_seq_num_scans.argtypes = [C.c_void_p]
        L.lins_seq_desc.argtypes = [C.c_void_p, C.POINTER(LinsBatchDesc)]
        L.lins_seq_array.restype = C.POINTER(C.c_double)
        L.lins_seq_array.argtypes = [C.c_void_p, C.c_int]
        L.lins_seq_ints.restype = C.POINTER(C.c_int32)
        L.lins_seq_ints.argtypes = [C.c_void_p, C.c_int]
        L.lins_seq_write_bag.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
        L.lins_seq_run_bag.restype = C.c_void_p
        L.lins_seq_run_bag.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        _SEQ = L
    return _SEQ


def _seq_record(L, h):
    n, ns = L.lins_seq_num_units(h), L.lins_seq_num_scans(h)
    d = LinsBatchDesc()
    L.lins_seq_desc(h, C.byref(d))
    clouds, offsets = {}, {}
    for k in Batch.FIELDS:
        off = _copy(getattr(d, k + "_off"), n + 1, np.int32)
        offsets[k] = off
        clouds[k] = _copy(getattr(d, k), int(off[-1]), POINT_DTYPE)
    arr = lambda which, cnt: np.ctypeslib.as_array(L.lins_seq_array(h, which), shape=(cnt,)).copy() if cnt else np.zeros(0)  # noqa: E731
    ints = lambda which, cnt: np.ctypeslib.as_array(L.lins_seq_ints(h, which), shape=(cnt,)).copy() if cnt else np.zeros(0, np.int32)  # noqa: E731
    units = Batch(clouds, offsets, _copy(d.state_in, n * 19, np.float64), _copy(d.cov_in, n * 324, np.float64), arr(1, n * 7)) if n else None
    return dict(units=units, state_out=arr(0, n * 19).reshape(-1, 19), iters=ints(0, n), flags=ints(1, n), scan_index=ints(2, n),
                status=ints(3, ns), global_est=arr(2, ns * 7).reshape(-1, 7), global_true=arr(3, ns * 7).reshape(-1, 7))


def write_sequence_bag(path, config="config3", seed=1, n_scans=12, lidar_topic="/velodyne_points", imu_topic="/imu/data", **overrides):
    """The synthetic drive of run_sequence written as a ROS1 bag (raw sweeps + IMU): no GPU needed."""
    L = seq_lib()
    kw = dict(CONFIGS[config])
    kw.update(overrides)
    cfg = SynthCfg(**kw)
    rc = L.lins_seq_write_bag(C.byref(cfg), seed, n_scans, path.encode(), lidar_topic.encode(), imu_topic.encode())
    if rc != 0:
        raise RuntimeError(f"lins_seq_write_bag failed with {rc}")


def run_bag(path, lidar_topic="/velodyne_points", imu_topic="/imu/data", max_scans=0, lidar_model=0, device=0):
    """BASELINE.json configs[1] runner: replay a ROS1 bag (sensor_msgs/PointCloud2 + sensor_msgs/Imu) through image
    projection, feature extraction and the GPU IESKF update, the way LinsFusion does (Estimator.cpp:123-284)."""
    L = seq_lib()
    err = C.c_int(0)
    h = L.lins_seq_run_bag(path.encode(), lidar_topic.encode(), imu_topic.encode(), max_scans, lidar_model, device, C.byref(err))
    if not h:
        raise RuntimeError(f"lins_seq_run_bag failed with {err.value}")
    try:
        return _seq_record(L, h)
    finally:
        L.lins_seq_destroy(h)


def run_sequence(config="config3", seed=1, n_scans=12, device=0, **overrides):
    """Drive fusion::StateEstimator (GPU hot path) over a synthetic sequence.  Returns the recorded performIESKF
    units as a Batch plus the shim's outputs."""
    L = seq_lib()
    kw = dict(CONFIGS[config])
    kw.update(overrides)
    cfg = SynthCfg(**kw)
    h = L.lins_seq_run(C.byref(cfg), seed, n_scans, device)
    try:
        return _seq_record(L, h)
    finally:
        L.lins_seq_destroy(h)


# ---- row F2: scan-to-map units (tools/synth/lins_synth.cpp: lins_synth_map_unit_create) -------------------------------
class MapUnit:
    """One scan2MapOptimization input: map clouds, the newest scan's (down-sampled) features, true / guessed transform."""

    def __init__(self, corner_map, surf_map, corner_last, surf_last, truth, guess):
        self.corner_map, self.surf_map, self.corner_last, self.surf_last = corner_map, surf_map, corner_last, surf_last
        self.truth, self.guess = truth, guess


def generate_map_unit(config="config3", seed=1, n_keyframes=20, sigma_t=0.1, sigma_r=0.01, **overrides):
    L = lib()
    L.lins_synth_map_unit_create.restype = C.c_void_p
    L.lins_synth_map_unit_create.argtypes = [C.POINTER(SynthCfg), C.c_uint64, C.c_int, C.c_double, C.c_double]
    L.lins_synth_map_unit_destroy.argtypes = [C.c_void_p]
    L.lins_synth_map_unit_cloud.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.lins_synth_map_unit_transforms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    kw = dict(CONFIGS[config])
    kw.update(overrides)
    cfg = SynthCfg(**kw)
    h = L.lins_synth_map_unit_create(C.byref(cfg), seed, n_keyframes, sigma_t, sigma_r)
    try:
        clouds = []
        for which in range(4):
            p = C.c_void_p()
            n = L.lins_synth_map_unit_cloud(h, which, C.byref(p))
            clouds.append(_copy(p.value, n, POINT_DTYPE))
        truth, guess = np.zeros(6, np.float32), np.zeros(6, np.float32)
        L.lins_synth_map_unit_transforms(h, truth.ctypes.data_as(C.c_void_p), guess.ctypes.data_as(C.c_void_p))
        return MapUnit(*clouds, truth, guess)
    finally:
        L.lins_synth_map_unit_destroy(h)
