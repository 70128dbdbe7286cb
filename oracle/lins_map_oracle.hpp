// oracle/lins_map_oracle.hpp — TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// Row F2 of SURVEY.md §8(f): the mapping node's scan-to-map refinement
//   scan2MapOptimization  lins/src/lidar_mapping_node.cpp:1635-1652
//   cornerOptimization    :1351-1461      surfOptimization  :1463-1524      LMOptimization  :1526-1633
//   pointAssociateToMap   :594-608        updatePointAssociateToMapSinCos   :579-592
// restated on plain arrays, all f32 exactly where the reference is f32 (float/double promotion of every literal is
// reproduced: `0.1 * v`, `1 - 0.9 * fabs(..)`, `> 0.2`, `> 0.1` are double arithmetic on float operands).
//
// Third-party pieces the reference calls and that are NOT in /root/reference (OpenCV 3.x/4.x core, FLANN through
// pcl::KdTreeFLANN) are restated from their published algorithms:
//   cv::eigen(symmetric CV_32F)      -> jacobi_eigen   (cyclic Jacobi with row/column maxima bookkeeping)
//   cv::solve(.., DECOMP_QR) CV_32F  -> qr_solve       (Householder QR, hFactors, back substitution)
//   cv::Mat::inv() (DECOMP_LU) 6x6   -> lu_invert      (partial-pivot LU on [A | I])
//   small cv::gemm CV_32F            -> f64 accumulation in k order, rounded to f32 once
//   nearestKSearch(k = 5)            -> exact 5-NN, f32 L2_Simple ((dx*dx)+dy*dy)+dz*dz, ascending distance,
//                                       ties by lower index (FLANN's tie order is traversal dependent)
// PARITY: unlike the IESKF rows this row is PINNED for its numerical kernels: tests/test_map_oracle_cpu.py checks
// jacobi_eigen / qr_solve / lu_invert bit-for-bit against the real OpenCV (cv2 4.13, importable in the build
// container) on thousands of random inputs, and the whole refinement against a Python restatement that calls
// cv2.eigen / cv2.solve directly (tests/golden/make_map_golden.py).  Not pinned: the big A^T A product (the
// reference's cv::gemm may run through BLAS with a different summation order) — f64 accumulation here.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/lins_gpu.h"

namespace lins_map_oracle {

// ---- OpenCV core restatements (f32) ------------------------------------------------------------------------------
inline float cv_hypot(float a, float b) {
  a = std::fabs(a); b = std::fabs(b);
  if (a > b) { b /= a; return a * std::sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * std::sqrt(1 + a * a); }
  return 0.f;
}

// A (n x n, row-major, symmetric, destroyed) -> eigenvalues W (descending) and eigenvectors as ROWS of V
inline void jacobi_eigen(float* A, int n, float* W, float* V) {
  const float eps = std::numeric_limits<float>::epsilon();
  std::vector<int> indR(n), indC(n);
  for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) V[i * n + j] = 0.f; V[i * n + i] = 1.f; }
  auto updR = [&](int k) {
    if (k < n - 1) {
      int m = k + 1; float mv = std::fabs(A[n * k + m]);
      for (int i = k + 2; i < n; ++i) { const float val = std::fabs(A[n * k + i]); if (mv < val) { mv = val; m = i; } }
      indR[k] = m;
    }
  };
  auto updC = [&](int k) {
    if (k > 0) {
      int m = 0; float mv = std::fabs(A[k]);
      for (int i = 1; i < k; ++i) { const float val = std::fabs(A[n * i + k]); if (mv < val) { mv = val; m = i; } }
      indC[k] = m;
    }
  };
  for (int k = 0; k < n; ++k) { W[k] = A[(n + 1) * k]; updR(k); updC(k); }
  if (n > 1) for (int iters = 0, maxIters = n * n * 30; iters < maxIters; ++iters) {
    int k = 0; float mv = std::fabs(A[indR[0]]);
    for (int i = 1; i < n - 1; ++i) { const float val = std::fabs(A[n * i + indR[i]]); if (mv < val) { mv = val; k = i; } }
    int l = indR[k];
    for (int i = 1; i < n; ++i) { const float val = std::fabs(A[n * indC[i] + i]); if (mv < val) { mv = val; k = indC[i]; l = i; } }
    const float p = A[n * k + l];
    if (std::fabs(p) <= eps) break;
    const float y = (float)((W[l] - W[k]) * 0.5);
    float t = std::fabs(y) + cv_hypot(p, y);
    float s = cv_hypot(p, t);
    const float c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[n * k + l] = 0;
    W[k] -= t; W[l] += t;
    auto rot = [&](float& v0, float& v1) { const float a0 = v0, b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; };
    for (int i = 0; i < k; ++i) rot(A[n * i + k], A[n * i + l]);
    for (int i = k + 1; i < l; ++i) rot(A[n * k + i], A[n * i + l]);
    for (int i = l + 1; i < n; ++i) rot(A[n * k + i], A[n * l + i]);
    for (int i = 0; i < n; ++i) rot(V[n * k + i], V[n * l + i]);
    for (int j = 0; j < 2; ++j) { const int idx = j == 0 ? k : l; updR(idx); updC(idx); }
  }
  for (int k = 0; k < n - 1; ++k) {
    int m = k;
    for (int i = k + 1; i < n; ++i) if (W[m] < W[i]) m = i;
    if (k != m) { std::swap(W[m], W[k]); for (int i = 0; i < n; ++i) std::swap(V[n * m + i], V[n * k + i]); }
  }
}

// least squares / linear solve by Householder QR: A (m x n, row-major, destroyed), b (m, destroyed) -> x = b[0..n)
inline bool qr_solve(float* A, int m, int n, float* b) {
  const float eps = std::numeric_limits<float>::epsilon();
  std::vector<float> vl(m), hF(n);
  for (int l = 0; l < n; ++l) {
    const int vs = m - l;
    float nrm = 0.f;
    for (int i = 0; i < vs; ++i) { vl[i] = A[(l + i) * n + l]; nrm += vl[i] * vl[i]; }
    const float tmp = vl[0];
    vl[0] = vl[0] + (vl[0] > 0 ? 1 : -1) * std::sqrt(nrm);
    nrm = std::sqrt(nrm + vl[0] * vl[0] - tmp * tmp);
    for (int i = 0; i < vs; ++i) vl[i] /= nrm;
    for (int j = l; j < n; ++j) {
      float v = 0.f;
      for (int i = l; i < m; ++i) v += vl[i - l] * A[i * n + j];
      for (int i = l; i < m; ++i) A[i * n + j] -= 2 * vl[i - l] * v;
    }
    hF[l] = vl[0] * vl[0];
    for (int i = 1; i < vs; ++i) A[(l + i) * n + l] = vl[i] / vl[0];
  }
  for (int l = 0; l < n; ++l) {
    vl[0] = 1.f;
    for (int j = 1; j < m - l; ++j) vl[j] = A[(j + l) * n + l];
    float v = 0.f;
    for (int i = l; i < m; ++i) v += vl[i - l] * b[i];
    for (int i = l; i < m; ++i) b[i] -= 2 * vl[i - l] * v * hF[l];
  }
  for (int i = n - 1; i >= 0; --i) {
    for (int j = n - 1; j > i; --j) b[i] -= b[j] * A[i * n + j];
    if (std::fabs(A[i * n + i]) < eps) return false;
    b[i] /= A[i * n + i];
  }
  return true;
}

// inverse by LU with partial pivoting on [A | I]; A (n x n, destroyed), Ainv (n x n); false = singular
inline bool lu_invert(float* A, int n, float* Ainv) {
  const float eps = std::numeric_limits<float>::epsilon() * 10;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Ainv[i * n + j] = i == j ? 1.f : 0.f;
  for (int i = 0; i < n; ++i) {
    int k = i;
    for (int j = i + 1; j < n; ++j) if (std::fabs(A[j * n + i]) > std::fabs(A[k * n + i])) k = j;
    if (std::fabs(A[k * n + i]) < eps) return false;
    if (k != i) {
      for (int j = i; j < n; ++j) std::swap(A[i * n + j], A[k * n + j]);
      for (int j = 0; j < n; ++j) std::swap(Ainv[i * n + j], Ainv[k * n + j]);
    }
    const float d = -1 / A[i * n + i];
    for (int j = i + 1; j < n; ++j) {
      const float alpha = A[j * n + i] * d;
      for (int c = i + 1; c < n; ++c) A[j * n + c] += alpha * A[i * n + c];
      for (int c = 0; c < n; ++c) Ainv[j * n + c] += alpha * Ainv[i * n + c];
    }
  }
  for (int i = n - 1; i >= 0; --i)
    for (int j = 0; j < n; ++j) {
      float s = Ainv[i * n + j];
      for (int c = i + 1; c < n; ++c) s -= A[i * n + c] * Ainv[c * n + j];
      Ainv[i * n + j] = s / A[i * n + i];
    }
  return true;
}

// C (m x n) = A (m x k) * B (k x n), f32 data, f64 accumulation in k order (OpenCV's built-in small GEMM)
inline void gemm_f32(const float* A, const float* B, float* C, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int c = 0; c < k; ++c) s += (double)A[i * k + c] * (double)B[c * n + j];
      C[i * n + j] = (float)s;
    }
}

// ---- the refinement ----------------------------------------------------------------------------------------------
struct P3 { float x, y, z, intensity; };

struct Transform {  // transformTobeMapped: rx, ry, rz, tx, ty, tz
  float v[6];
};
struct SinCos { float cRoll, sRoll, cPitch, sPitch, cYaw, sYaw, tX, tY, tZ; };
inline SinCos sincos_of(const Transform& T) {  // :579-592 (float overloads of cos / sin)
  SinCos s;
  s.cRoll = std::cos(T.v[0]); s.sRoll = std::sin(T.v[0]);
  s.cPitch = std::cos(T.v[1]); s.sPitch = std::sin(T.v[1]);
  s.cYaw = std::cos(T.v[2]); s.sYaw = std::sin(T.v[2]);
  s.tX = T.v[3]; s.tY = T.v[4]; s.tZ = T.v[5];
  return s;
}
inline P3 point_associate_to_map(const P3& pi, const SinCos& s) {  // :594-608
  const float x1 = s.cYaw * pi.x - s.sYaw * pi.y;
  const float y1 = s.sYaw * pi.x + s.cYaw * pi.y;
  const float z1 = pi.z;
  const float x2 = x1;
  const float y2 = s.cRoll * y1 - s.sRoll * z1;
  const float z2 = s.sRoll * y1 + s.cRoll * z1;
  P3 po;
  po.x = s.cPitch * x2 + s.sPitch * z2 + s.tX;
  po.y = y2 + s.tY;
  po.z = -s.sPitch * x2 + s.cPitch * z2 + s.tZ;
  po.intensity = pi.intensity;
  return po;
}

inline float sqdist(const P3& a, const P3& b) {
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return ((dx * dx) + dy * dy) + dz * dz;
}
// exact 5-NN (brute force): ascending (distance, index); fewer than 5 targets -> idx -1, dist +inf for the rest
inline void knn5(const std::vector<P3>& map, const P3& q, int idx[5], float dist[5]) {
  for (int k = 0; k < 5; ++k) { idx[k] = -1; dist[k] = std::numeric_limits<float>::infinity(); }
  for (int j = 0; j < (int)map.size(); ++j) {
    const float d = sqdist(q, map[j]);
    if (!(d < dist[4])) continue;  // strict: a later index never displaces an equal distance
    int k = 4;
    while (k > 0 && d < dist[k - 1]) { dist[k] = dist[k - 1]; idx[k] = idx[k - 1]; --k; }
    dist[k] = d; idx[k] = j;
  }
}

struct FitOut { float coeff[4]; bool sel; };

// the body of cornerOptimization for one point (:1360-1458)
inline FitOut corner_fit(const std::vector<P3>& map, const P3& sel, const int ind[5], const float dist[5]) {
  FitOut o{{0, 0, 0, 0}, false};
  if (!(dist[4] < 1.0)) return o;
  float cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < 5; ++j) { cx += map[ind[j]].x; cy += map[ind[j]].y; cz += map[ind[j]].z; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
  for (int j = 0; j < 5; ++j) {
    const float ax = map[ind[j]].x - cx, ay = map[ind[j]].y - cy, az = map[ind[j]].z - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az; a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  float A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33}, D[3], V[9];
  jacobi_eigen(A, 3, D, V);
  if (D[0] > 3 * D[1]) {
    const float x0 = sel.x, y0 = sel.y, z0 = sel.z;
    const float x1 = cx + 0.1 * V[0], y1 = cy + 0.1 * V[1], z1 = cz + 0.1 * V[2];
    const float x2 = cx - 0.1 * V[0], y2 = cy - 0.1 * V[1], z2 = cz - 0.1 * V[2];
    const float a012 = std::sqrt(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                                 ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                                 ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
    const float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
    const float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                      (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
    const float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) -
                       (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
    const float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                       (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
    const float ld2 = a012 / l12;
    const float s = 1 - 0.9 * std::fabs(ld2);
    o.coeff[0] = s * la; o.coeff[1] = s * lb; o.coeff[2] = s * lc; o.coeff[3] = s * ld2;
    o.sel = s > 0.1;
  }
  return o;
}

// the body of surfOptimization for one point (:1471-1521)
inline FitOut surf_fit(const std::vector<P3>& map, const P3& sel, const int ind[5], const float dist[5]) {
  FitOut o{{0, 0, 0, 0}, false};
  if (!(dist[4] < 1.0)) return o;
  float A0[15], B0[5] = {-1, -1, -1, -1, -1};
  for (int j = 0; j < 5; ++j) { A0[3 * j] = map[ind[j]].x; A0[3 * j + 1] = map[ind[j]].y; A0[3 * j + 2] = map[ind[j]].z; }
  float X0[3] = {0, 0, 0};
  if (qr_solve(A0, 5, 3, B0)) { X0[0] = B0[0]; X0[1] = B0[1]; X0[2] = B0[2]; }  // a failed cv::solve leaves matX0 at its last value; zeros here
  float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
  const float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  bool planeValid = true;
  for (int j = 0; j < 5; ++j)
    if (std::fabs(pa * map[ind[j]].x + pb * map[ind[j]].y + pc * map[ind[j]].z + pd) > 0.2) { planeValid = false; break; }
  if (planeValid) {
    const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
    const float s = 1 - 0.9 * std::fabs(pd2) / std::sqrt(std::sqrt(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z));
    o.coeff[0] = s * pa; o.coeff[1] = s * pb; o.coeff[2] = s * pc; o.coeff[3] = s * pd2;
    o.sel = s > 0.1;
  }
  return o;
}

// one row of matA / matB (:1549-1593)
inline void lm_row(const P3& po, const float c[4], float srx, float crx, float sry, float cry, float srz, float crz, float row[6],
                   float& b) {
  const float arx = (crx * sry * srz * po.x + crx * crz * sry * po.y - srx * sry * po.z) * c[0] +
                    (-srx * srz * po.x - crz * srx * po.y - crx * po.z) * c[1] +
                    (crx * cry * srz * po.x + crx * cry * crz * po.y - cry * srx * po.z) * c[2];
  const float ary = ((cry * srx * srz - crz * sry) * po.x + (sry * srz + cry * crz * srx) * po.y + crx * cry * po.z) * c[0] +
                    ((-cry * crz - srx * sry * srz) * po.x + (cry * srz - crz * srx * sry) * po.y - crx * sry * po.z) * c[2];
  const float arz = ((crz * srx * sry - cry * srz) * po.x + (-cry * crz - srx * sry * srz) * po.y) * c[0] +
                    (crx * crz * po.x - crx * srz * po.y) * c[1] +
                    ((sry * srz + cry * crz * srx) * po.x + (crz * sry - cry * srx * srz) * po.y) * c[2];
  row[0] = arx; row[1] = ary; row[2] = arz; row[3] = c[0]; row[4] = c[1]; row[5] = c[2];
  b = -c[3];
}

struct Report {
  int iters = 0, converged = 0, degenerate = 0, skipped = 0;
  std::vector<int> n_sel;
  std::vector<float> delta_r, delta_t;
};

struct Mapper {
  std::vector<P3> cornerMap, surfMap;  // laserCloudCornerFromMapDS / laserCloudSurfFromMapDS
  bool isDegenerate = false;
  float matP[36];

  void setMap(const lins_point* c, int nc, const lins_point* s, int ns) {
    cornerMap.resize(nc); surfMap.resize(ns);
    for (int i = 0; i < nc; ++i) cornerMap[i] = P3{c[i].x, c[i].y, c[i].z, c[i].intensity};
    for (int i = 0; i < ns; ++i) surfMap[i] = P3{s[i].x, s[i].y, s[i].z, s[i].intensity};
  }

  // one cornerOptimization + surfOptimization pass with dense outputs (any pointer may be null)
  void associate(const std::vector<P3>& cornerLast, const std::vector<P3>& surfLast, const Transform& T, int32_t* cknn, int32_t* sknn,
                 float* ccoeff, float* scoeff, uint8_t* cmask, uint8_t* smask, std::vector<P3>* ori, std::vector<FitOut>* sel) const {
    const SinCos sc = sincos_of(T);
    for (int pass = 0; pass < 2; ++pass) {
      const std::vector<P3>& q = pass == 0 ? cornerLast : surfLast;
      const std::vector<P3>& map = pass == 0 ? cornerMap : surfMap;
      int32_t* knn = pass == 0 ? cknn : sknn;
      float* coeff = pass == 0 ? ccoeff : scoeff;
      uint8_t* mask = pass == 0 ? cmask : smask;
      for (int i = 0; i < (int)q.size(); ++i) {
        const P3 ps = point_associate_to_map(q[i], sc);
        int ind[5]; float dist[5];
        knn5(map, ps, ind, dist);
        const FitOut f = pass == 0 ? corner_fit(map, ps, ind, dist) : surf_fit(map, ps, ind, dist);
        if (knn) for (int k = 0; k < 5; ++k) knn[5 * i + k] = ind[k];
        if (coeff) for (int k = 0; k < 4; ++k) coeff[4 * i + k] = f.coeff[k];
        if (mask) mask[i] = f.sel ? 1 : 0;
        if (f.sel && ori) { ori->push_back(q[i]); sel->push_back(f); }
      }
    }
  }

  // LMOptimization (:1526-1633); returns true on convergence
  bool lm(const std::vector<P3>& ori, const std::vector<FitOut>& sel, int iterCount, Transform& T, Report& rep) {
    const float srx = std::sin(T.v[0]), crx = std::cos(T.v[0]), sry = std::sin(T.v[1]), cry = std::cos(T.v[1]);
    const float srz = std::sin(T.v[2]), crz = std::cos(T.v[2]);
    const int n = (int)ori.size();
    rep.n_sel.push_back(n);
    if (n < 50) { rep.delta_r.push_back(0); rep.delta_t.push_back(0); return false; }
    double AtA[36] = {0}, AtB[6] = {0};
    for (int i = 0; i < n; ++i) {
      float row[6], b;
      lm_row(ori[i], sel[i].coeff, srx, crx, sry, cry, srz, crz, row, b);
      for (int a = 0; a < 6; ++a) { for (int c = 0; c < 6; ++c) AtA[a * 6 + c] += (double)row[a] * (double)row[c]; AtB[a] += (double)row[a] * (double)b; }
    }
    float fAtA[36], fAtB[6], X[6];
    for (int i = 0; i < 36; ++i) fAtA[i] = (float)AtA[i];
    for (int i = 0; i < 6; ++i) fAtB[i] = (float)AtB[i];
    return lm_solve(fAtA, fAtB, iterCount, T, rep, X);
  }
  // the part of LMOptimization after matAtA / matAtB exist (:1598-1632)
  bool lm_solve(const float* fAtA, const float* fAtB, int iterCount, Transform& T, Report& rep, float X[6]) {
    float Aw[36], Bw[6];
    std::memcpy(Aw, fAtA, sizeof(Aw)); std::memcpy(Bw, fAtB, sizeof(Bw));
    if (qr_solve(Aw, 6, 6, Bw)) std::memcpy(X, Bw, sizeof(float) * 6);
    else for (int i = 0; i < 6; ++i) X[i] = 0.f;
    if (iterCount == 0) {
      float E[6], V[36], V2[36], Ae[36];
      std::memcpy(Ae, fAtA, sizeof(Ae));
      jacobi_eigen(Ae, 6, E, V);
      std::memcpy(V2, V, sizeof(V2));
      isDegenerate = false;
      const float eignThre[6] = {100, 100, 100, 100, 100, 100};
      for (int i = 5; i >= 0; --i) {
        if (E[i] < eignThre[i]) { for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0; isDegenerate = true; }
        else break;
      }
      float Vc[36], Vinv[36];
      std::memcpy(Vc, V, sizeof(Vc));
      if (!lu_invert(Vc, 6, Vinv)) std::memset(Vinv, 0, sizeof(Vinv));
      gemm_f32(Vinv, V2, matP, 6, 6, 6);
    }
    if (isDegenerate) {
      float X2[6];
      std::memcpy(X2, X, sizeof(X2));
      gemm_f32(matP, X2, X, 6, 6, 1);
    }
    for (int i = 0; i < 6; ++i) T.v[i] += X[i];
    const double r2d = 180.0 / M_PI;  // pcl::rad2deg(float) = alpha * 57.29578f
    auto rad2deg = [](float a) { return a * 57.29578f; };
    (void)r2d;
    const float deltaR = std::sqrt(std::pow(rad2deg(X[0]), 2) + std::pow(rad2deg(X[1]), 2) + std::pow(rad2deg(X[2]), 2));
    const float deltaT = std::sqrt(std::pow(X[3] * 100, 2) + std::pow(X[4] * 100, 2) + std::pow(X[5] * 100, 2));
    rep.delta_r.push_back(deltaR); rep.delta_t.push_back(deltaT);
    rep.degenerate = isDegenerate ? 1 : 0;
    return deltaR < 0.05 && deltaT < 0.05;
  }

  // scan2MapOptimization (:1635-1652) without transformUpdate
  void scan2map(const std::vector<P3>& cornerLast, const std::vector<P3>& surfLast, Transform& T, Report& rep) {
    rep = Report();
    if (!(cornerMap.size() > 10 && surfMap.size() > 100)) { rep.skipped = 1; return; }
    for (int iterCount = 0; iterCount < 10; ++iterCount) {
      std::vector<P3> ori; std::vector<FitOut> sel;
      associate(cornerLast, surfLast, T, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &ori, &sel);
      rep.iters = iterCount + 1;
      if (lm(ori, sel, iterCount, T, rep)) { rep.converged = 1; break; }
    }
  }
};

}  // namespace lins_map_oracle
