// lins_cv_small.hpp — the small dense f32 kernels the mapping node takes from OpenCV core (row F2), for host and
// device.  PRODUCT code.  OpenCV is a third-party dependency of the reference (not in /root/reference), so these are
// restated from its published algorithms — the same ones the CPU oracle restates independently and pins bit for bit
// against cv2 (tests/test_map_oracle_cpu.py):
//   cv::eigen on a symmetric CV_32F matrix      cyclic Jacobi with per-row / per-column maxima bookkeeping, eigenvalues
//                                               descending, eigenvectors as rows           (lidar_mapping_node.cpp:1402, :1605)
//   cv::solve(A, b, x, DECOMP_QR), CV_32F       Householder QR + back substitution (least squares when rows > cols)
//                                                                                          (:1478, :1598)
//   cv::Mat::inv() (DECOMP_LU), CV_32F          partial-pivot LU on [A | I]                (:1618)
//   small cv::gemm, CV_32F                      f64 accumulation in k order, one rounding  (:1618, :1624)
// Every operation is a single IEEE f32 (or, where noted, f64) operation in a fixed order; the translation unit is
// compiled without FMA contraction.
#ifndef LINS_HOST_CV_SMALL_HPP_
#define LINS_HOST_CV_SMALL_HPP_

#include <math.h>

#if defined(__CUDACC__)
#define LINS_HD __host__ __device__ __forceinline__
#else
#define LINS_HD inline
#endif

namespace lins_cv {

constexpr float kFltEps = 1.1920928955078125e-07f;

LINS_HD float hypot2(float a, float b) {
  a = fabsf(a); b = fabsf(b);
  if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
  if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
  return 0.f;
}

// A: N x N row-major symmetric (destroyed).  W: eigenvalues, descending.  V: eigenvectors as rows.
template <int N>
LINS_HD void jacobi_eigen(float* A, float* W, float* V) {
  int indR[N], indC[N];
  for (int i = 0; i < N; ++i) { for (int j = 0; j < N; ++j) V[i * N + j] = 0.f; V[i * N + i] = 1.f; }
  for (int k = 0; k < N; ++k) {
    W[k] = A[(N + 1) * k];
    indR[k] = 0; indC[k] = 0;
    if (k < N - 1) {
      int m = k + 1; float mv = fabsf(A[N * k + m]);
      for (int i = k + 2; i < N; ++i) { const float val = fabsf(A[N * k + i]); if (mv < val) { mv = val; m = i; } }
      indR[k] = m;
    }
    if (k > 0) {
      int m = 0; float mv = fabsf(A[k]);
      for (int i = 1; i < k; ++i) { const float val = fabsf(A[N * i + k]); if (mv < val) { mv = val; m = i; } }
      indC[k] = m;
    }
  }
  if (N > 1) for (int iters = 0; iters < N * N * 30; ++iters) {
    int k = 0; float mv = fabsf(A[indR[0]]);
    for (int i = 1; i < N - 1; ++i) { const float val = fabsf(A[N * i + indR[i]]); if (mv < val) { mv = val; k = i; } }
    int l = indR[k];
    for (int i = 1; i < N; ++i) { const float val = fabsf(A[N * indC[i] + i]); if (mv < val) { mv = val; k = indC[i]; l = i; } }
    const float p = A[N * k + l];
    if (fabsf(p) <= kFltEps) break;
    const float y = (float)((double)(W[l] - W[k]) * 0.5);
    float t = fabsf(y) + hypot2(p, y);
    float s = hypot2(p, t);
    const float c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[N * k + l] = 0;
    W[k] -= t; W[l] += t;
    for (int i = 0; i < k; ++i) { const float a0 = A[N * i + k], b0 = A[N * i + l]; A[N * i + k] = a0 * c - b0 * s; A[N * i + l] = a0 * s + b0 * c; }
    for (int i = k + 1; i < l; ++i) { const float a0 = A[N * k + i], b0 = A[N * i + l]; A[N * k + i] = a0 * c - b0 * s; A[N * i + l] = a0 * s + b0 * c; }
    for (int i = l + 1; i < N; ++i) { const float a0 = A[N * k + i], b0 = A[N * l + i]; A[N * k + i] = a0 * c - b0 * s; A[N * l + i] = a0 * s + b0 * c; }
    for (int i = 0; i < N; ++i) { const float a0 = V[N * k + i], b0 = V[N * l + i]; V[N * k + i] = a0 * c - b0 * s; V[N * l + i] = a0 * s + b0 * c; }
    for (int j = 0; j < 2; ++j) {
      const int idx = j == 0 ? k : l;
      if (idx < N - 1) {
        int m = idx + 1; float mv2 = fabsf(A[N * idx + m]);
        for (int i = idx + 2; i < N; ++i) { const float val = fabsf(A[N * idx + i]); if (mv2 < val) { mv2 = val; m = i; } }
        indR[idx] = m;
      }
      if (idx > 0) {
        int m = 0; float mv2 = fabsf(A[idx]);
        for (int i = 1; i < idx; ++i) { const float val = fabsf(A[N * i + idx]); if (mv2 < val) { mv2 = val; m = i; } }
        indC[idx] = m;
      }
    }
  }
  for (int k = 0; k < N - 1; ++k) {
    int m = k;
    for (int i = k + 1; i < N; ++i) if (W[m] < W[i]) m = i;
    if (k != m) {
      const float tw = W[m]; W[m] = W[k]; W[k] = tw;
      for (int i = 0; i < N; ++i) { const float tv = V[N * m + i]; V[N * m + i] = V[N * k + i]; V[N * k + i] = tv; }
    }
  }
}

// A: M x N row-major (destroyed), b: M (destroyed) -> x = b[0..N).  false: rank deficient (|R_ii| < eps).
template <int M, int N>
LINS_HD bool qr_solve(float* A, float* b) {
  float vl[M], hF[N];
  for (int l = 0; l < N; ++l) {
    const int vs = M - l;
    float nrm = 0.f;
    for (int i = 0; i < vs; ++i) { vl[i] = A[(l + i) * N + l]; nrm += vl[i] * vl[i]; }
    const float tmp = vl[0];
    vl[0] = vl[0] + (vl[0] > 0 ? 1 : -1) * sqrtf(nrm);
    nrm = sqrtf(nrm + vl[0] * vl[0] - tmp * tmp);
    for (int i = 0; i < vs; ++i) vl[i] /= nrm;
    for (int j = l; j < N; ++j) {
      float v = 0.f;
      for (int i = l; i < M; ++i) v += vl[i - l] * A[i * N + j];
      for (int i = l; i < M; ++i) A[i * N + j] -= 2 * vl[i - l] * v;
    }
    hF[l] = vl[0] * vl[0];
    for (int i = 1; i < vs; ++i) A[(l + i) * N + l] = vl[i] / vl[0];
  }
  for (int l = 0; l < N; ++l) {
    vl[0] = 1.f;
    for (int j = 1; j < M - l; ++j) vl[j] = A[(j + l) * N + l];
    float v = 0.f;
    for (int i = l; i < M; ++i) v += vl[i - l] * b[i];
    for (int i = l; i < M; ++i) b[i] -= 2 * vl[i - l] * v * hF[l];
  }
  for (int i = N - 1; i >= 0; --i) {
    for (int j = N - 1; j > i; --j) b[i] -= b[j] * A[i * N + j];
    if (fabsf(A[i * N + i]) < kFltEps) return false;
    b[i] /= A[i * N + i];
  }
  return true;
}

template <int N>
LINS_HD bool lu_invert(float* A, float* Ainv) {
  const float eps = kFltEps * 10;
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) Ainv[i * N + j] = i == j ? 1.f : 0.f;
  for (int i = 0; i < N; ++i) {
    int k = i;
    for (int j = i + 1; j < N; ++j) if (fabsf(A[j * N + i]) > fabsf(A[k * N + i])) k = j;
    if (fabsf(A[k * N + i]) < eps) return false;
    if (k != i) {
      for (int j = i; j < N; ++j) { const float t = A[i * N + j]; A[i * N + j] = A[k * N + j]; A[k * N + j] = t; }
      for (int j = 0; j < N; ++j) { const float t = Ainv[i * N + j]; Ainv[i * N + j] = Ainv[k * N + j]; Ainv[k * N + j] = t; }
    }
    const float d = -1 / A[i * N + i];
    for (int j = i + 1; j < N; ++j) {
      const float alpha = A[j * N + i] * d;
      for (int c = i + 1; c < N; ++c) A[j * N + c] += alpha * A[i * N + c];
      for (int c = 0; c < N; ++c) Ainv[j * N + c] += alpha * Ainv[i * N + c];
    }
  }
  for (int i = N - 1; i >= 0; --i)
    for (int j = 0; j < N; ++j) {
      float s = Ainv[i * N + j];
      for (int c = i + 1; c < N; ++c) s -= A[i * N + c] * Ainv[c * N + j];
      Ainv[i * N + j] = s / A[i * N + i];
    }
  return true;
}

template <int M, int K, int N>
LINS_HD void gemm(const float* A, const float* B, float* C) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0;
      for (int c = 0; c < K; ++c) s += (double)A[i * K + c] * (double)B[c * N + j];
      C[i * N + j] = (float)s;
    }
}

}  // namespace lins_cv
#endif
