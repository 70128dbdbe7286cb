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
