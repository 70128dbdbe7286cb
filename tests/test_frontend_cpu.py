"""CPU suite: the product's C++ front end (csrc/host/image_projection.hpp, feature_extraction.hpp — what feeds every
synthetic unit, the StateEstimator shim and the bag replay) against an INDEPENDENT Python restatement written from the
reference source (tests/pyfront.py; lins/src/image_projection_node.cpp:191-415, lins/include/StateEstimator.hpp:619-827).
Bit-exact: segmented / outlier clouds, cloud_info, the de-skew time stamps and all four feature clouds.
"""
import ctypes as C
import os

import numpy as np
import pytest

import pyfront

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def front(synth, defs):
    L = C.CDLL(os.path.join(ROOT, "tools", "synth", "liblins_synth.so"))
    vp = C.c_void_p
    L.lins_synth_raw_sweep.argtypes = [C.POINTER(synth.SynthCfg), C.c_uint64, vp, C.c_int]
    L.lins_frontend_run.argtypes = [vp, C.c_int, C.c_int, C.c_int] + [vp] * 13 + [vp]

    def run(config, seed):
        cfg = synth.SynthCfg(**synth.CONFIGS[config])
        cap = 16 * 1800
        raw = np.zeros(cap, defs.POINT_DTYPE)
        n = L.lins_synth_raw_sweep(C.byref(cfg), seed, defs.ptr(raw), cap)
        raw = raw[:n].copy()
        P = lambda: np.zeros(cap, defs.POINT_DTYPE)  # noqa: E731
        seg, outl, und, sharp, lsharp, flat, lflat = P(), P(), P(), P(), P(), P(), P()
        sr, er, ori = np.zeros(16, np.int32), np.zeros(16, np.int32), np.zeros(3, np.float32)
        ground, col, rng, cnt = np.zeros(cap, np.uint8), np.zeros(cap, np.uint32), np.zeros(cap, np.float32), np.zeros(6, np.int32)
        rc = L.lins_frontend_run(defs.ptr(raw), n, 0, cap, defs.ptr(seg), defs.ptr(outl), defs.ptr(sr), defs.ptr(er), defs.ptr(ori), defs.ptr(ground),
                                 defs.ptr(col), defs.ptr(rng), defs.ptr(und), defs.ptr(sharp), defs.ptr(lsharp), defs.ptr(flat), defs.ptr(lflat), defs.ptr(cnt))
        assert rc == 0
        x4 = lambda a, k: np.stack([a["x"], a["y"], a["z"], a["intensity"]], 1)[:k]  # noqa: E731
        return dict(raw=np.stack([raw["x"], raw["y"], raw["z"]], 1), seg=x4(seg, cnt[0]), outlier=x4(outl, cnt[1]), start_ring=sr, end_ring=er, ori=ori,
                    ground=ground[: cnt[0]], col=col[: cnt[0]], range=rng[: cnt[0]], undist=x4(und, cnt[0]), sharp=x4(sharp, cnt[2]),
                    less_sharp=x4(lsharp, cnt[3]), flat=x4(flat, cnt[4]), less_flat=x4(lflat, cnt[5]))

    return run


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shapes {a.shape} vs {b.shape}"
    assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b, ), f"{what} differs"


@pytest.mark.parametrize("config,seed", [("config3", 5), ("config1", 2)])
def test_front_end_matches_python_restatement(front, config, seed):
    cpp = front(config, seed)
    assert len(cpp["raw"]) > 5000
    ip = pyfront.image_projection(cpp["raw"])
    _same(ip["seg"], cpp["seg"], "segmented cloud")
    _same(ip["outlier"], cpp["outlier"], "outlier cloud")
    _same(ip["start_ring"], cpp["start_ring"], "startRingIndex"); _same(ip["end_ring"], cpp["end_ring"], "endRingIndex")
    _same(np.array(ip["ori"], np.float32), cpp["ori"], "start / end orientation")
    _same(ip["ground"], cpp["ground"], "ground flags"); _same(ip["col"], cpp["col"], "column indices"); _same(ip["range"], cpp["range"], "ranges")
    assert 2000 < len(ip["seg"]) < 16 * 1800 and ip["ground"].sum() > 100
    fe = pyfront.extract_features(ip["seg"], ip)
    # (fe["sort_ties"] counts equal curvatures inside a std::sort range — there the reference's order is unspecified; the
    # simulated sweeps have a handful, none of them near a pick, so both restatements must still agree)
    _same(fe["undist"], cpp["undist"], "de-skew time stamps")
    for k in ("sharp", "less_sharp", "flat", "less_flat"):
        _same(fe[k], cpp[k], k)
    assert len(fe["sharp"]) > 10 and len(fe["flat"]) > 50 and len(fe["less_flat"]) > 500


def test_voxel_grid_restatement_properties():
    """pcl::VoxelGrid restatement on its own: one centroid per occupied 0.2 m voxel, all fields averaged, ascending voxel index."""
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.uniform(-3, 3, (500, 3)), rng.uniform(0, 16, (500, 1))], 1).astype(np.float32)
    out = pyfront.voxel_grid(pts)
    vox = np.floor(pts[:, :3] * np.float32(5.0)).astype(int)
    assert len(out) == len({tuple(v) for v in vox})
    assert np.allclose(out[:, :3].mean(0), pts[:, :3].mean(0), atol=0.15)
    assert np.isclose(out[:, 3].min(), pts[:, 3].min(), atol=2) and len(pyfront.voxel_grid(np.zeros((0, 4), np.float32))) == 0
