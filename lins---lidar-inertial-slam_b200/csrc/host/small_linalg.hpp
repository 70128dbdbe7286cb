// Host-side dense kernels for the 6-DoF Gauss-Newton fallback (reference: lins/include/StateEstimator.hpp
// calculateTransformation :1260-1302): the reference calls Eigen's colPivHouseholderQr().solve,
// SelfAdjointEigenSolver and inverse() on 6x6 matrices.  Eigen is not available here; these are independent
// N x N templates (N small).  PRODUCT code — shares nothing with oracle/.
#ifndef LINS_HOST_SMALL_LINALG_HPP_
#define LINS_HOST_SMALL_LINALG_HPP_

#include <array>
#include <cmath>
#include <limits>
#include <numeric>
#include <utility>

namespace lins {
namespace linalg {

template <int N>
using Mat = std::array<std::array<double, N>, N>;
template <int N>
using Vec = std::array<double, N>;

// x = argmin ||A x - b|| through Householder QR with column pivoting.  Rank is decided like
// Eigen::ColPivHouseholderQR (|R_kk| > eps * N * max|R_kk|); free components are set to zero.
template <int N>
Vec<N> colPivQrSolve(Mat<N> A, Vec<N> b) {
  std::array<int, N> perm;
  std::iota(perm.begin(), perm.end(), 0);
  double maxpiv = 0.0;
  for (int k = 0; k < N; ++k) {
    // pivot: remaining column with the largest tail norm
    int best = k;
    double bestn = -1.0;
    for (int j = k; j < N; ++j) {
      double s = 0;
      for (int i = k; i < N; ++i) s += A[i][j] * A[i][j];
      if (s > bestn) { bestn = s; best = j; }
    }
    if (best != k) {
      for (int i = 0; i < N; ++i) std::swap(A[i][k], A[i][best]);
      std::swap(perm[k], perm[best]);
    }
    double nrm = std::sqrt(bestn > 0 ? bestn : 0.0);
    if (nrm == 0.0) continue;
    // reflector v = x - alpha e1
    double alpha = A[k][k] > 0 ? -nrm : nrm;
    Vec<N> v{};
    for (int i = k; i < N; ++i) v[i] = A[i][k];
    v[k] -= alpha;
    double vv = 0;
    for (int i = k; i < N; ++i) vv += v[i] * v[i];
    if (vv > 0) {
      for (int j = k; j < N; ++j) {
        double d = 0;
        for (int i = k; i < N; ++i) d += v[i] * A[i][j];
        d *= 2.0 / vv;
        for (int i = k; i < N; ++i) A[i][j] -= d * v[i];
      }
      double d = 0;
      for (int i = k; i < N; ++i) d += v[i] * b[i];
      d *= 2.0 / vv;
      for (int i = k; i < N; ++i) b[i] -= d * v[i];
    }
    maxpiv = std::max(maxpiv, std::fabs(A[k][k]));
  }
  const double thr = std::numeric_limits<double>::epsilon() * N * maxpiv;
  int rank = 0;
  while (rank < N && std::fabs(A[rank][rank]) > thr) ++rank;
  Vec<N> y{};
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < rank; ++j) s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  Vec<N> x{};
  for (int k = 0; k < N; ++k) x[perm[k]] = y[k];
  return x;
}

// Symmetric eigen-decomposition (cyclic Jacobi).  Eigenvalues ascending, eigenvector k = column k of V,
// normalised so that its largest-magnitude component is positive.
template <int N>
void symmetricEigen(Mat<N> A, Vec<N>& evals, Mat<N>& V) {
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int i = 0; i < N; ++i) for (int j = i + 1; j < N; ++j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        if (A[p][q] == 0.0) continue;
        double tau = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < N; ++k) { double x = A[k][p], y = A[k][q]; A[k][p] = c * x - s * y; A[k][q] = s * x + c * y; }
        for (int k = 0; k < N; ++k) { double x = A[p][k], y = A[q][k]; A[p][k] = c * x - s * y; A[q][k] = s * x + c * y; }
        for (int k = 0; k < N; ++k) { double x = V[k][p], y = V[k][q]; V[k][p] = c * x - s * y; V[k][q] = s * x + c * y; }
      }
  }
  std::array<int, N> ord;
  std::iota(ord.begin(), ord.end(), 0);
  for (int i = 1; i < N; ++i) {  // insertion sort by eigenvalue
    int o = ord[i], j = i - 1;
    while (j >= 0 && A[ord[j]][ord[j]] > A[o][o]) { ord[j + 1] = ord[j]; --j; }
    ord[j + 1] = o;
  }
  Mat<N> Vs;
  for (int k = 0; k < N; ++k) {
    evals[k] = A[ord[k]][ord[k]];
    int big = 0;
    for (int i = 1; i < N; ++i) if (std::fabs(V[i][ord[k]]) > std::fabs(V[big][ord[k]])) big = i;
    double sg = V[big][ord[k]] < 0 ? -1.0 : 1.0;
    for (int i = 0; i < N; ++i) Vs[i][k] = sg * V[i][ord[k]];
  }
  V = Vs;
}

// Gauss-Jordan inverse with partial pivoting; returns false if singular.
template <int N>
bool inverse(Mat<N> A, Mat<N>& inv) {
  for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) inv[i][j] = i == j ? 1.0 : 0.0;
  for (int k = 0; k < N; ++k) {
    int piv = k;
    for (int i = k + 1; i < N; ++i) if (std::fabs(A[i][k]) > std::fabs(A[piv][k])) piv = i;
    if (A[piv][k] == 0.0 || std::isnan(A[piv][k])) return false;
    if (piv != k) { std::swap(A[piv], A[k]); std::swap(inv[piv], inv[k]); }
    double d = 1.0 / A[k][k];
    for (int j = 0; j < N; ++j) { A[k][j] *= d; inv[k][j] *= d; }
    for (int i = 0; i < N; ++i) {
      if (i == k) continue;
      double f = A[i][k];
      if (f == 0.0) continue;
      for (int j = 0; j < N; ++j) { A[i][j] -= f * A[k][j]; inv[i][j] -= f * inv[k][j]; }
    }
  }
  return true;
}

}  // namespace linalg
}  // namespace lins
#endif
