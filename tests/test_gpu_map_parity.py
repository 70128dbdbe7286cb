"""GPU suite, row F2: the mapping node's scan-to-map refinement through the C-ABI
(lins_gpu_map_set / lins_gpu_map_associate / lins_gpu_scan2map) against the CPU oracle and the fixture generated
with the real OpenCV (tests/golden/make_map_golden.py).

Bar: 5-NN indices, accept masks and coefficients bit-exact (the device computes them with IEEE f32 + - * / sqrt in the
reference's order; sin / cos of the transform are evaluated on the host); transform within 1e-5 rad / m of the
oracle (only the summation order of A^T A differs), well inside north_star's 1e-4."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "map_unit.npz")
T_TOL = 1e-5


def _same_pass(a, b, ctx=""):
    for k in ("corner_knn", "surf_knn", "corner_mask", "surf_mask"):
        assert np.array_equal(a[k], b[k]), f"{ctx} {k}"
    for k in ("corner_coeff", "surf_coeff"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), f"{ctx} {k} not bit-exact"


def test_map_pass_matches_cv2_golden_and_oracle(gpu, ob):
    g = np.load(GOLD)
    gpu.map_set(g["corner_map"], g["surf_map"])
    out = gpu.map_associate(g["corner_last"], g["surf_last"], g["guess"])
    _same_pass(out, {k[4:]: g[k] for k in g.files if k.startswith("it0_")}, "vs cv2 golden")
    m = ob.MapOracle(); m.set_map(g["corner_map"], g["surf_map"])
    # a second linearisation point: the refined transform
    _same_pass(gpu.map_associate(g["corner_last"], g["surf_last"], g["T_out"]), m.associate(g["corner_last"], g["surf_last"], g["T_out"]), "vs oracle")


def test_scan2map_matches_golden_and_oracle(gpu, ob):
    g = np.load(GOLD)
    gpu.map_set(g["corner_map"], g["surf_map"])
    T, rep = gpu.scan2map(g["corner_last"], g["surf_last"], g["guess"])
    assert rep.iters == int(g["iters"]) and rep.converged == int(g["converged"]) and rep.degenerate == int(g["degenerate"])
    assert list(rep.n_sel)[:rep.iters] == list(g["n_sel"])
    assert np.abs(T - g["T_out"]).max() < T_TOL
    m = ob.MapOracle(); m.set_map(g["corner_map"], g["surf_map"])
    To, ro = m.scan2map(g["corner_last"], g["surf_last"], g["guess"])
    assert np.abs(T - To).max() < T_TOL and rep.iters == ro.iters
    assert np.allclose(list(rep.delta_r)[:rep.iters], list(ro.delta_r)[:ro.iters], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("seed,kf", [(3, 20), (8, 30)])
def test_scan2map_larger_maps(gpu, ob, synth, seed, kf):
    """20 / 40 key-frames (27 k / 50 k+ surf map points): several map slices per query block."""
    u = synth.generate_map_unit("config3", seed=seed, n_keyframes=kf, sigma_t=0.1, sigma_r=0.01)
    gpu.map_set(u.corner_map, u.surf_map)
    m = ob.MapOracle(); m.set_map(u.corner_map, u.surf_map)
    _same_pass(gpu.map_associate(u.corner_last, u.surf_last, u.guess), m.associate(u.corner_last, u.surf_last, u.guess), f"seed {seed}")
    T, rep = gpu.scan2map(u.corner_last, u.surf_last, u.guess)
    To, ro = m.scan2map(u.corner_last, u.surf_last, u.guess)
    assert rep.iters == ro.iters and rep.converged == ro.converged and list(rep.n_sel) == list(ro.n_sel)
    assert np.abs(T - To).max() < T_TOL
    # it refines (how much depends on the scene: along a straight road the driving direction is weakly constrained)
    assert np.linalg.norm(T[3:] - u.truth[3:]) < np.linalg.norm(u.guess[3:] - u.truth[3:])
    assert np.abs(T[:3] - u.truth[:3]).max() < np.abs(u.guess[:3] - u.truth[:3]).max()


def test_map_edge_cases(gpu, ob):
    g = np.load(GOLD)
    m = ob.MapOracle()
    # map too small: nothing happens (:1636)
    gpu.map_set(g["corner_map"][:10], g["surf_map"])
    T, rep = gpu.scan2map(g["corner_last"], g["surf_last"], g["guess"])
    assert rep.skipped == 1 and rep.iters == 0 and np.array_equal(T, g["guess"])
    # fewer than 50 selected points: 10 passes, transform untouched (:1535)
    gpu.map_set(g["corner_map"], g["surf_map"])
    T, rep = gpu.scan2map(g["corner_last"][:20], g["surf_last"][:20], g["guess"])
    assert rep.iters == 10 and rep.converged == 0 and np.array_equal(T, g["guess"])
    # empty feature clouds, fewer than five map points, far-away queries: same dense outputs as the oracle
    for cm, sm, cq, sq in ((g["corner_map"], g["surf_map"][:3], g["corner_last"][:0], g["surf_last"][:7]),
                           (g["corner_map"][:4], g["surf_map"], g["corner_last"][:9], g["surf_last"][:0])):
        gpu.map_set(cm, sm); m.set_map(cm, sm)
        _same_pass(gpu.map_associate(cq, sq, g["guess"]), m.associate(cq, sq, g["guess"]), "tiny map")
    far = g["surf_last"][:64].copy(); far["x"] += 300
    gpu.map_set(g["corner_map"], g["surf_map"]); m.set_map(g["corner_map"], g["surf_map"])
    a = gpu.map_associate(g["corner_last"][:5], far, g["guess"])
    _same_pass(a, m.associate(g["corner_last"][:5], far, g["guess"]), "far queries")
    assert a["surf_mask"].sum() == 0


@pytest.mark.parametrize("seed,kf", [(5, 50), (11, 25)])
def test_grid_knn_equals_brute_force(gpu, ob, synth, seed, kf):
    """The hashed-grid 5-NN that lins_gpu_scan2map uses (LINS_MAP_KNN=grid forces it for the dense parity hook too) must
    give, for every point the 1 m gate can accept, exactly the brute-force neighbours / coefficients / masks; the points
    whose fifth neighbour is farther than 1 m are rejected by both (lidar_mapping_node.cpp:1374, :1481)."""
    u = synth.generate_map_unit("config3", seed=seed, n_keyframes=kf, sigma_t=0.1, sigma_r=0.01)
    gpu.map_set(u.corner_map, u.surf_map)
    old = os.environ.get("LINS_MAP_KNN")
    try:
        os.environ["LINS_MAP_KNN"] = "brute"
        b = gpu.map_associate(u.corner_last, u.surf_last, u.guess)
        os.environ["LINS_MAP_KNN"] = "grid"
        g = gpu.map_associate(u.corner_last, u.surf_last, u.guess)
    finally:
        if old is None:
            os.environ.pop("LINS_MAP_KNN", None)
        else:
            os.environ["LINS_MAP_KNN"] = old
    for name, cloud, mp in (("corner", u.corner_last, u.corner_map), ("surf", u.surf_last, u.surf_map)):
        assert np.array_equal(b[name + "_mask"], g[name + "_mask"])
        assert np.array_equal(b[name + "_coeff"].view(np.uint32), g[name + "_coeff"].view(np.uint32))
        # where the brute-force fifth neighbour is within 1 m the five indices agree exactly
        m = ob.MapOracle(); m.set_map(u.corner_map, u.surf_map)
        sel = b[name + "_mask"].astype(bool)
        assert sel.sum() > 20
        assert np.array_equal(b[name + "_knn"][sel], g[name + "_knn"][sel])
        # and every index the grid returns is a real neighbour index
        assert (g[name + "_knn"] >= -1).all() and (g[name + "_knn"] < len(mp)).all()
