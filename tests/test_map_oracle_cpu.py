"""Row F2 (mapping-node scan-to-map refinement): the oracle's numerical kernels are PINNED against the real OpenCV.

The reference calls cv::eigen / cv::solve(DECOMP_QR) / cv::Mat::inv on CV_32F data
(lins/src/lidar_mapping_node.cpp:1402, :1478, :1598-1620).  OpenCV is not part of /root/reference, but cv2 4.13 is
importable in the build container, so oracle/lins_map_oracle.hpp's restatements are compared bit for bit with it
(skipped where cv2 is missing, e.g. on a box without it).  The whole refinement is pinned through the golden
fixtures of tests/golden/make_map_golden.py (a Python restatement that calls cv2.eigen / cv2.solve itself)."""
import numpy as np
import pytest

from oracle import oracle_binding as ob

cv2 = pytest.importorskip("cv2")
f32 = np.float32


def _sym(rng, n):
    B = rng.standard_normal((5 if n == 3 else 60, n)).astype(f32) * f32(rng.uniform(0.01, 3))
    A = (B.T @ B / f32(5)).astype(f32)
    return ((A + A.T) / 2).astype(f32)


def test_jacobi_eigen_is_cv_eigen_bit_for_bit():
    rng = np.random.default_rng(11)
    for t in range(1500):
        A = _sym(rng, 3 if t % 2 == 0 else 6)
        ok, w, v = cv2.eigen(A)
        W, V = ob.cv_eigen(A)
        assert ok and np.array_equal(w.ravel(), W) and np.array_equal(v, V), t
    # near-degenerate inputs of the corner fit: five almost collinear points
    for t in range(300):
        d = rng.standard_normal(3); d /= np.linalg.norm(d)
        P = (np.outer(rng.uniform(-0.5, 0.5, 5), d) + rng.standard_normal((5, 3)) * 1e-3 + rng.uniform(-30, 30, 3)).astype(f32)
        c = (P.sum(0) / f32(5)).astype(f32)
        Q = (P - c).astype(f32)
        A = (Q.T @ Q / f32(5)).astype(f32); A = ((A + A.T) / 2).astype(f32)
        ok, w, v = cv2.eigen(A)
        W, V = ob.cv_eigen(A)
        assert np.array_equal(w.ravel(), W) and np.array_equal(v, V), t


def test_qr_solve_is_cv_solve_bit_for_bit():
    rng = np.random.default_rng(12)
    for t in range(1500):
        if t % 2 == 0:  # the 5 x 3 plane fit (:1471-1478): map points, rhs -1
            A = (rng.standard_normal((5, 3)) * rng.uniform(0.05, 2) + rng.uniform(-40, 40, 3)).astype(f32)
            b = -np.ones((5, 1), f32)
        else:  # the 6 x 6 normal equations (:1598)
            B = rng.standard_normal((80, 6)).astype(f32) * f32(rng.uniform(0.1, 5))
            A = (B.T @ B).astype(f32); b = rng.standard_normal((6, 1)).astype(f32)
        ok, x = cv2.solve(A, b, flags=cv2.DECOMP_QR)
        ok2, y = ob.cv_qr_solve(A, b)
        assert ok == ok2 and np.array_equal(x.ravel(), y), t


def test_lu_invert_and_small_gemm_are_cv_bit_for_bit():
    rng = np.random.default_rng(13)
    for t in range(500):
        A = _sym(rng, 6)
        _, _, V = cv2.eigen(A)  # what LMOptimization inverts is an eigenvector matrix (:1618)
        Vi = cv2.invert(V, flags=cv2.DECOMP_LU)[1]
        ok, Wi = ob.cv_lu_invert(V)
        assert ok and np.array_equal(Vi, Wi), t
        V2 = V.copy(); V2[rng.integers(3, 6):] = 0
        assert np.array_equal(cv2.gemm(Vi, V2, 1.0, None, 0.0), ob.cv_gemm(Wi, V2))
        M = rng.standard_normal((6, 6)).astype(f32) * f32(rng.uniform(0.1, 10))
        ok, Mi = ob.cv_lu_invert(M)
        assert ok and np.array_equal(cv2.invert(M, flags=cv2.DECOMP_LU)[1], Mi), t


# ---- the whole refinement against the fixture made with the real OpenCV (tests/golden/make_map_golden.py) -----------
import os  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "map_unit.npz")


def _gold():
    return np.load(GOLD)


def test_oracle_pass_matches_cv2_golden_bit_for_bit():
    """Iteration 0 of cornerOptimization / surfOptimization: 5-NN indices, coefficients and masks are pure f32
    arithmetic + cv::eigen / cv::solve, so the oracle must reproduce the cv2-made fixture exactly."""
    g = _gold()
    m = ob.MapOracle()
    m.set_map(g["corner_map"], g["surf_map"])
    out = m.associate(g["corner_last"], g["surf_last"], g["guess"])
    for k in ("corner_knn", "surf_knn", "corner_mask", "surf_mask"):
        assert np.array_equal(out[k], g["it0_" + k]), k
    for k in ("corner_coeff", "surf_coeff"):
        assert np.array_equal(out[k].view(np.uint32), g["it0_" + k].view(np.uint32)), k
    assert out["corner_mask"].sum() > 50 and out["surf_mask"].sum() > 1000


def test_oracle_scan2map_matches_cv2_golden():
    """Whole loop.  matAtA = matAt * matA goes through cv::gemm, whose summation order depends on the OpenCV build
    (BLAS here); the oracle accumulates in f64, so the transform is compared at 1e-5 (rad / m), the rest exactly."""
    g = _gold()
    m = ob.MapOracle()
    m.set_map(g["corner_map"], g["surf_map"])
    T, rep = m.scan2map(g["corner_last"], g["surf_last"], g["guess"])
    assert rep.iters == int(g["iters"]) and rep.converged == int(g["converged"]) and rep.degenerate == int(g["degenerate"])
    assert list(rep.n_sel)[:rep.iters] == list(g["n_sel"])
    assert np.abs(T - g["T_out"]).max() < 1e-5
    assert np.allclose(np.array(list(rep.delta_r)[:rep.iters]), g["delta_r"], rtol=2e-3, atol=1e-5)
    # and the refinement does refine: translation error shrinks by > 5x on this unit
    assert np.abs(T[3:] - g["truth"][3:]).max() < 0.2 * np.abs(g["guess"][3:] - g["truth"][3:]).max()


def test_lm_step_on_cv2_normal_equations_is_exact():
    """Given the fixture's own matAtA / matAtB (made by cv2.gemm), the 6x6 step (QR solve, eigen-degeneracy test,
    update, deltaR / deltaT) is bit-exact."""
    g = _gold()
    m = ob.MapOracle()
    T = g["guess"].astype(np.float32).copy()
    m2 = ob.MapOracle()
    for it in range(int(g["iters"])):
        A, b = g["AtA"][it], g["AtB"][it]
        T2, x, conv, deg = m2.lm_solve(A, b, it, T)
        ok, xr = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_QR)
        assert np.array_equal(x, xr.ravel()) and not deg
        assert abs(float(np.sqrt(sum((float(np.float32(v * np.float32(57.29578))) ** 2 for v in x[:3])))) - float(g["delta_r"][it])) < 1e-6
        T = T2


def test_map_edge_cases():
    g = _gold()
    m = ob.MapOracle()
    # map too small: scan2MapOptimization does nothing (:1636)
    m.set_map(g["corner_map"][:10], g["surf_map"])
    T, rep = m.scan2map(g["corner_last"], g["surf_last"], g["guess"])
    assert rep.skipped == 1 and rep.iters == 0 and np.array_equal(T, g["guess"])
    # fewer than 50 selected points: LMOptimization returns false every time, 10 passes, transform untouched (:1535)
    m.set_map(g["corner_map"], g["surf_map"])
    T, rep = m.scan2map(g["corner_last"][:20], g["surf_last"][:20], g["guess"])
    assert rep.iters == 10 and rep.converged == 0 and np.array_equal(T, g["guess"]) and max(list(rep.n_sel)) < 50
    # fewer than 5 map points around: indices -1, nothing selected
    far = g["surf_last"][:4].copy(); far["x"] += 500
    out = m.associate(g["corner_last"][:0], far, g["guess"])
    assert out["surf_mask"].sum() == 0 and (out["surf_knn"] >= 0).all()  # the map has >= 5 points, they are just > 1 m away
    m.set_map(g["corner_map"], g["surf_map"][:3])
    out = m.associate(g["corner_last"][:0], g["surf_last"][:4], g["guess"])
    assert (out["surf_knn"][:, 3:] == -1).all() and out["surf_mask"].sum() == 0
