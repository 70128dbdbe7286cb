"""ctypes mirrors of the C structs in include/lins_gpu.h (the drop-in boundary).

Plain data only: no compute lives here.  numpy structured dtype POINT_DTYPE is layout-identical to
``lins_point`` / ``pcl::PointXYZI`` (32 B; reference lins/include/parameters.h:52).
"""
import ctypes as C

import numpy as np

LINS_MAX_ITER = 64
STATE_DIM = 19
COV_SIZE = 324

POINT_DTYPE = np.dtype(
    {
        "names": ["x", "y", "z", "pad0", "intensity", "pad1", "pad2", "pad3"],
        "formats": [np.float32] * 8,
        "itemsize": 32,
    }
)


class LinsParams(C.Structure):
    _fields_ = [
        ("num_iter", C.c_int32),
        ("icp_freq", C.c_int32),
        ("nearest_feature_search_sq_dist", C.c_double),
        ("lidar_std", C.c_double),
        ("lidar_scale", C.c_double),
        ("scan_period", C.c_double),
        ("verbose", C.c_int32),
        ("force_all_iters", C.c_int32),
    ]

    @classmethod
    def shipped(cls, **kw):
        """lins/config/exp_config/exp_port.yaml:9-20 values."""
        p = cls(30, 1, 25.0, 0.01, 1.0, 0.1, 0, 0)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class LinsReport(C.Structure):
    _fields_ = [
        ("iters", C.c_int32),
        ("converged", C.c_int32),
        ("diverged", C.c_int32),
        ("has_nan", C.c_int32),
        ("m_surf", C.c_int32 * LINS_MAX_ITER),
        ("m_corner", C.c_int32 * LINS_MAX_ITER),
        ("residual_norm", C.c_double * LINS_MAX_ITER),
        ("update_norm", C.c_double * LINS_MAX_ITER),
    ]


class LinsMapReport(C.Structure):
    """lins_map_report (include/lins_gpu.h, row F2)."""
    _fields_ = [("iters", C.c_int32), ("converged", C.c_int32), ("degenerate", C.c_int32), ("skipped", C.c_int32),
                ("n_sel", C.c_int32 * 10), ("delta_r", C.c_float * 10), ("delta_t", C.c_float * 10)]


class LinsScanResult(C.Structure):
    _fields_ = [
        ("scan_id", C.c_int32),
        ("iters", C.c_uint16),
        ("flags", C.c_uint16),
        ("pose", C.c_double * 7),
    ]


SCAN_RESULT_DTYPE = np.dtype(
    [("scan_id", np.int32), ("iters", np.uint16), ("flags", np.uint16), ("pose", np.float64, 7)]
)
assert SCAN_RESULT_DTYPE.itemsize == 64 == C.sizeof(LinsScanResult)


class LinsBatchDesc(C.Structure):
    _fields_ = [
        ("n_scans", C.c_int32),
        ("surf_flat", C.c_void_p),
        ("surf_flat_off", C.c_void_p),
        ("corner_sharp", C.c_void_p),
        ("corner_sharp_off", C.c_void_p),
        ("surf_less_flat", C.c_void_p),
        ("surf_less_flat_off", C.c_void_p),
        ("corner_less_sharp", C.c_void_p),
        ("corner_less_sharp_off", C.c_void_p),
        ("state_in", C.c_void_p),
        ("cov_in", C.c_void_p),
        ("point_format", C.c_int32),  # 0 = 32-B PointXYZI records, 1 = packed 16-B (x, y, z, intensity)
    ]


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def as_points(a):
    """Return a C-contiguous POINT_DTYPE array view/copy of `a` (n x 8 float32 also accepted)."""
    a = np.asarray(a)
    if a.dtype != POINT_DTYPE:
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 8).view(POINT_DTYPE).reshape(-1)
    return np.ascontiguousarray(a)


def make_points(xyz, intensity):
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    p = np.zeros(len(xyz), dtype=POINT_DTYPE)
    p["x"], p["y"], p["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    p["pad0"] = 1.0
    p["intensity"] = np.asarray(intensity, dtype=np.float32)
    return p


class Batch:
    """Host-side batch of independent (scan pair, prior) units, CSR layout of lins_batch_desc."""

    FIELDS = ("surf_flat", "corner_sharp", "surf_less_flat", "corner_less_sharp")

    def __init__(self, clouds, offsets, state, cov, truth=None, extra=None):
        self.clouds = {k: as_points(clouds[k]) for k in self.FIELDS}
        self.offsets = {k: np.ascontiguousarray(offsets[k], dtype=np.int32) for k in self.FIELDS}
        self.state = np.ascontiguousarray(state, dtype=np.float64).reshape(-1, STATE_DIM)
        self.cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(-1, COV_SIZE)
        self.truth = None if truth is None else np.ascontiguousarray(truth, dtype=np.float64).reshape(-1, 7)
        self.extra = extra or {}
        self.n = len(self.state)
        for k in self.FIELDS:
            assert len(self.offsets[k]) == self.n + 1 and self.offsets[k][-1] == len(self.clouds[k])

    def desc(self):
        d = LinsBatchDesc()
        d.n_scans = self.n
        for k in self.FIELDS:
            setattr(d, k, self.clouds[k].ctypes.data)
            setattr(d, k + "_off", self.offsets[k].ctypes.data)
        d.state_in = self.state.ctypes.data
        d.cov_in = self.cov.ctypes.data
        return d

    def packed16(self):
        """The same batch with its clouds as packed (x, y, z, intensity) float32 records (lins_batch_desc.point_format = 1)."""
        return PackedBatch(self)

    def unit(self, i):
        """The four clouds + prior of unit i."""
        out = {}
        for k in self.FIELDS:
            o = self.offsets[k]
            out[k] = self.clouds[k][o[i] : o[i + 1]]
        out["state"] = self.state[i]
        out["cov"] = self.cov[i]
        return out

    def subset(self, idx):
        idx = list(idx)
        clouds, offsets = {}, {}
        for k in self.FIELDS:
            o = self.offsets[k]
            parts = [self.clouds[k][o[i] : o[i + 1]] for i in idx]
            clouds[k] = np.concatenate(parts) if parts else np.zeros(0, POINT_DTYPE)
            offsets[k] = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int32)
        return Batch(clouds, offsets, self.state[idx], self.cov[idx], None if self.truth is None else self.truth[idx])

    def tile(self, reps):
        """The same units repeated `reps` times (used to push the working set past L2)."""
        clouds, offsets = {}, {}
        for k in self.FIELDS:
            clouds[k] = np.tile(self.clouds[k], reps)
            sizes = np.diff(self.offsets[k])
            offsets[k] = np.concatenate([[0], np.cumsum(np.tile(sizes, reps))]).astype(np.int32)
        return Batch(clouds, offsets, np.tile(self.state, (reps, 1)), np.tile(self.cov, (reps, 1)),
                     None if self.truth is None else np.tile(self.truth, (reps, 1)))

    def save(self, path):
        arrs = {}
        for k in self.FIELDS:
            arrs[k] = self.clouds[k].view(np.float32).reshape(-1, 8)[:, [0, 1, 2, 4]]
            arrs[k + "_off"] = self.offsets[k]
        arrs["state"], arrs["cov"] = self.state, self.cov
        if self.truth is not None:
            arrs["truth"] = self.truth
        np.savez_compressed(path, **arrs)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        clouds = {k: make_points(z[k][:, :3], z[k][:, 3]) for k in cls.FIELDS}
        offsets = {k: z[k + "_off"] for k in cls.FIELDS}
        return cls(clouds, offsets, z["state"], z["cov"], z["truth"] if "truth" in z.files else None)


class PackedBatch:
    """A Batch whose clouds are stored as 16-byte (x, y, z, intensity) records: LINS_POINTS_PACKED16."""

    FIELDS = Batch.FIELDS

    def __init__(self, batch):
        self.n, self.offsets, self.state, self.cov = batch.n, batch.offsets, batch.state, batch.cov
        self.clouds = {}
        for k in self.FIELDS:
            c = batch.clouds[k]
            self.clouds[k] = np.ascontiguousarray(np.stack([c["x"], c["y"], c["z"], c["intensity"]], 1).astype(np.float32)) if len(c) else np.zeros((0, 4), np.float32)

    def desc(self):
        d = LinsBatchDesc()
        d.n_scans = self.n
        for k in self.FIELDS:
            setattr(d, k, self.clouds[k].ctypes.data)
            setattr(d, k + "_off", self.offsets[k].ctypes.data)
        d.state_in = self.state.ctypes.data
        d.cov_in = self.cov.ctypes.data
        d.point_format = 1
        return d
