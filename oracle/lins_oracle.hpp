// oracle/lins_oracle.hpp — TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// CPU restatement of the reference's iterated-ESKF update path, written
// from the reference's behaviour, with zero third-party dependencies.  Each function cites the reference
// file:line it follows (paths relative to /root/reference/lins/include/).
//
// PARITY UNPINNED: the reference ships no tests / golden vectors and cannot be compiled in
// this image (needs ROS, PCL, FLANN, Eigen, OpenCV, gtsam), so this restatement is the definition of the
// parity contract rather than being checked against reference outputs.  Two third-party behaviours are
// restated from their published algorithms:
//   * pcl::KdTreeFLANN<PointXYZI>::nearestKSearch(k=1) -> flann::KDTreeSingleIndex<L2_Simple<float>>:
//     exact 1-NN, squared distance accumulated in f32 as ((dx*dx)+dy*dy)+dz*dz, non-finite targets skipped.
//     FLANN's tie-break between exactly-equidistant targets depends on tree traversal order; here it is
//     fixed to the LOWEST target index (tests assert the fixtures contain no exact 1-NN ties).
//   * Eigen expression evaluation order (see lins_math.hpp), LLT / ColPivHouseholderQR /
//     SelfAdjointEigenSolver (tolerance-level only).
// One deliberate, documented deviation: the forward ring walk's loop bound `j < surfPointsFlatNum`
// (StateEstimator.hpp:859, :983 — the QUERY count, a reference quirk that is kept) is additionally clamped to
// the target cloud size, where the reference would read out of bounds (undefined behaviour).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use oracle/.
#ifndef LINS_ORACLE_HPP_
#define LINS_ORACLE_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/lins_gpu.h"
#include "lins_math.hpp"

namespace lins_oracle {

typedef lins_point PointType;  // parameters.h:52

// ---------------------------------------------------------------------------------------------------------
// 18-dim error-state helpers
// ---------------------------------------------------------------------------------------------------------
struct Vec18 {
  double v[18];
  Vec18() { std::memset(v, 0, sizeof(v)); }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double norm() const {
    double s = 0;
    for (int i = 0; i < 18; ++i) s += v[i] * v[i];
    return std::sqrt(s);
  }
};
struct Mat18 {
  double m[18][18];  // m[row][col]
  Mat18() { std::memset(m, 0, sizeof(m)); }
  static Mat18 Identity() {
    Mat18 r;
    for (int i = 0; i < 18; ++i) r.m[i][i] = 1.0;
    return r;
  }
  double& operator()(int r, int c) { return m[r][c]; }
  double operator()(int r, int c) const { return m[r][c]; }
};
inline Mat18 mul(const Mat18& a, const Mat18& b) {
  Mat18 r;
  for (int i = 0; i < 18; ++i)
    for (int k = 0; k < 18; ++k) {
      double aik = a.m[i][k];
      for (int j = 0; j < 18; ++j) r.m[i][j] += aik * b.m[k][j];
    }
  return r;
}
inline Mat18 transpose(const Mat18& a) {
  Mat18 r;
  for (int i = 0; i < 18; ++i)
    for (int j = 0; j < 18; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
inline Mat18 add(const Mat18& a, const Mat18& b) {
  Mat18 r;
  for (int i = 0; i < 18; ++i)
    for (int j = 0; j < 18; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
// math_utils.h:39-41 enforceSymmetry
inline void enforceSymmetry(Mat18& a) {
  Mat18 t = transpose(a);
  for (int i = 0; i < 18; ++i)
    for (int j = 0; j < 18; ++j) a.m[i][j] = 0.5 * (a.m[i][j] + t.m[i][j]);
}
inline void cov_from_colmajor(const double* p, Mat18& P) {
  for (int c = 0; c < 18; ++c)
    for (int r = 0; r < 18; ++r) P.m[r][c] = p[c * 18 + r];
}
inline void cov_to_colmajor(const Mat18& P, double* p) {
  for (int c = 0; c < 18; ++c)
    for (int r = 0; r < 18; ++r) p[c * 18 + r] = P.m[r][c];
}

// In-place LU with partial pivoting of an n x n row-major matrix, solving A X = B for nrhs right-hand sides
// (B row-major n x nrhs, overwritten by X).  Returns false on an exactly singular pivot.
inline bool lu_solve(double* A, int n, double* B, int nrhs) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      double v = std::fabs(A[i * n + k]);
      if (v > best) { best = v; piv = i; }
    }
    if (best == 0.0 || std::isnan(best)) return false;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[piv * n + j]);
      for (int j = 0; j < nrhs; ++j) std::swap(B[k * nrhs + j], B[piv * nrhs + j]);
    }
    double inv = 1.0 / A[k * n + k];
    for (int i = k + 1; i < n; ++i) {
      double f = A[i * n + k] * inv;
      if (f == 0.0) continue;
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
      for (int j = 0; j < nrhs; ++j) B[i * nrhs + j] -= f * B[k * nrhs + j];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double inv = 1.0 / A[k * n + k];
    for (int j = 0; j < nrhs; ++j) {
      double s = B[k * nrhs + j];
      for (int c = k + 1; c < n; ++c) s -= A[k * n + c] * B[c * nrhs + j];
      B[k * nrhs + j] = s * inv;
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// filter::GlobalState  (KalmanFilter.hpp:35-116)
// ---------------------------------------------------------------------------------------------------------
static const double G0 = 9.81;  // parameters.h:62
struct GlobalState {
  static constexpr int pos_ = 0, vel_ = 3, att_ = 6, acc_ = 9, gyr_ = 12, gra_ = 15;  // KalmanFilter.hpp:40-45
  V3 rn_, vn_;
  Q4 qbn_;
  V3 ba_, bw_, gn_;
  GlobalState() { setIdentity(); }
  void setIdentity() {  // KalmanFilter.hpp:61-68
    rn_ = V3(); vn_ = V3(); qbn_ = Q4(); ba_ = V3(); bw_ = V3();
    gn_ = V3(0.0, 0.0, -G0);
  }
  // KalmanFilter.hpp:71-81
  void boxPlus(const Vec18& xk, GlobalState& out) const {
    GlobalState r;
    r.rn_ = rn_ + V3(xk[pos_], xk[pos_ + 1], xk[pos_ + 2]);
    r.vn_ = vn_ + V3(xk[vel_], xk[vel_ + 1], xk[vel_ + 2]);
    r.ba_ = ba_ + V3(xk[acc_], xk[acc_ + 1], xk[acc_ + 2]);
    r.bw_ = bw_ + V3(xk[gyr_], xk[gyr_ + 1], xk[gyr_ + 2]);
    Q4 dq = axis2Quat(V3(xk[att_], xk[att_ + 1], xk[att_ + 2]));
    r.qbn_ = normalized(qbn_ * dq);
    r.gn_ = gn_ + V3(xk[gra_], xk[gra_ + 1], xk[gra_ + 2]);
    out = r;
  }
  // KalmanFilter.hpp:84-94
  void boxMinus(const GlobalState& in, Vec18& xk) const {
    V3 d;
    d = rn_ - in.rn_; xk[pos_] = d.x; xk[pos_ + 1] = d.y; xk[pos_ + 2] = d.z;
    d = vn_ - in.vn_; xk[vel_] = d.x; xk[vel_ + 1] = d.y; xk[vel_ + 2] = d.z;
    d = ba_ - in.ba_; xk[acc_] = d.x; xk[acc_ + 1] = d.y; xk[acc_ + 2] = d.z;
    d = bw_ - in.bw_; xk[gyr_] = d.x; xk[gyr_ + 1] = d.y; xk[gyr_ + 2] = d.z;
    V3 da = Quat2axis(inverse(in.qbn_) * qbn_);
    xk[att_] = da.x; xk[att_ + 1] = da.y; xk[att_ + 2] = da.z;
    d = gn_ - in.gn_; xk[gra_] = d.x; xk[gra_ + 1] = d.y; xk[gra_ + 2] = d.z;
  }
  // C-ABI layout: rn vn q(x,y,z,w) ba bw gn
  void to_array(double* s) const {
    s[0] = rn_.x; s[1] = rn_.y; s[2] = rn_.z; s[3] = vn_.x; s[4] = vn_.y; s[5] = vn_.z;
    s[6] = qbn_.x; s[7] = qbn_.y; s[8] = qbn_.z; s[9] = qbn_.w;
    s[10] = ba_.x; s[11] = ba_.y; s[12] = ba_.z; s[13] = bw_.x; s[14] = bw_.y; s[15] = bw_.z;
    s[16] = gn_.x; s[17] = gn_.y; s[18] = gn_.z;
  }
  static GlobalState from_array(const double* s) {
    GlobalState g;
    g.rn_ = V3(s[0], s[1], s[2]); g.vn_ = V3(s[3], s[4], s[5]);
    g.qbn_ = Q4(s[9], s[6], s[7], s[8]);
    g.ba_ = V3(s[10], s[11], s[12]); g.bw_ = V3(s[13], s[14], s[15]); g.gn_ = V3(s[16], s[17], s[18]);
    return g;
  }
};

// ---------------------------------------------------------------------------------------------------------
// Exact 1-NN restating pcl::KdTreeFLANN::nearestKSearch(k=1) semantics (see header).
// ---------------------------------------------------------------------------------------------------------
inline float sqdist_f32(const PointType& t, float qx, float qy, float qz) {
  float dx = qx - t.x, dy = qy - t.y, dz = qz - t.z;
  float r = dx * dx;  // flann L2_Simple: result += diff*diff, starting from 0
  r = r + dy * dy;
  r = r + dz * dz;
  return r;
}
inline bool finite_xyz(const PointType& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

// Brute-force ground truth. Returns index or -1 (empty cloud / no finite point / NaN query); *sq gets the distance.
inline int nn_brute(const std::vector<PointType>& cloud, float qx, float qy, float qz, float* sq) {
  int best = -1;
  float bestd = std::numeric_limits<float>::infinity();
  for (int j = 0; j < (int)cloud.size(); ++j) {
    if (!finite_xyz(cloud[j])) continue;
    float d = sqdist_f32(cloud[j], qx, qy, qz);
    if (d < bestd) { bestd = d; best = j; }  // strict <  ==> lowest index wins exact ties
  }
  *sq = bestd;
  return best;
}

// kd-tree (median split on the widest dimension, leaf size 15 like pcl's KDTreeSingleIndexParams(15)),
// returning exactly nn_brute()'s answer (same f32 distance expression, lowest-index tie-break, pruning
// only when the one-dimensional f32 lower bound is strictly greater than the current best).
class KdTree {
 public:
  void build(const std::vector<PointType>& cloud) {
    cloud_ = &cloud;
    idx_.clear(); nodes_.clear();
    for (int i = 0; i < (int)cloud.size(); ++i)
      if (finite_xyz(cloud[i])) idx_.push_back(i);
    if (!idx_.empty()) build_rec(0, (int)idx_.size());
    // reorder (flann `reorder=true`): copy xyz into traversal order for locality
    pts_.resize(idx_.size() * 3);
    for (size_t k = 0; k < idx_.size(); ++k) {
      const PointType& p = cloud[idx_[k]];
      pts_[3 * k] = p.x; pts_[3 * k + 1] = p.y; pts_[3 * k + 2] = p.z;
    }
  }
  int nearest(float qx, float qy, float qz, float* sq) const {
    int best = -1;
    float bestd = std::numeric_limits<float>::infinity();
    if (!nodes_.empty() && !(std::isnan(qx) || std::isnan(qy) || std::isnan(qz))) {
      float q[3] = {qx, qy, qz};
      search(0, q, best, bestd);
    }
    *sq = bestd;
    return best;
  }
  bool empty() const { return nodes_.empty(); }

 private:
  struct Node { int lo, hi, left, right, dim; float split_lo, split_hi; };
  const std::vector<PointType>* cloud_ = nullptr;
  std::vector<int> idx_;
  std::vector<float> pts_;
  std::vector<Node> nodes_;
  float coord(int i, int d) const {
    const PointType& p = (*cloud_)[i];
    return d == 0 ? p.x : (d == 1 ? p.y : p.z);
  }
  int build_rec(int lo, int hi) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, -1, 0.f, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3], mx[3];
    for (int d = 0; d < 3; ++d) { mn[d] = std::numeric_limits<float>::infinity(); mx[d] = -mn[d]; }
    for (int k = lo; k < hi; ++k)
      for (int d = 0; d < 3; ++d) { float c = coord(idx_[k], d); mn[d] = std::min(mn[d], c); mx[d] = std::max(mx[d], c); }
    int dim = 0;
    for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    if (!(mx[dim] > mn[dim])) return id;  // all coincident: keep as a (large) leaf
    int mid = (lo + hi) / 2;
    std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi,
                     [&](int a, int b) { float ca = coord(a, dim), cb = coord(b, dim); return ca < cb || (ca == cb && a < b); });
    float slo = -std::numeric_limits<float>::infinity(), shi = std::numeric_limits<float>::infinity();
    for (int k = lo; k < mid; ++k) slo = std::max(slo, coord(idx_[k], dim));
    for (int k = mid; k < hi; ++k) shi = std::min(shi, coord(idx_[k], dim));
    int l = build_rec(lo, mid);
    int r = build_rec(mid, hi);
    nodes_[id].left = l; nodes_[id].right = r; nodes_[id].dim = dim;
    nodes_[id].split_lo = slo;  // max coordinate of the left child
    nodes_[id].split_hi = shi;  // min coordinate of the right child
    return id;
  }
  void search(int id, const float* q, int& best, float& bestd) const {
    const Node& n = nodes_[id];
    if (n.left < 0) {
      for (int k = n.lo; k < n.hi; ++k) {
        float dx = q[0] - pts_[3 * k], dy = q[1] - pts_[3 * k + 1], dz = q[2] - pts_[3 * k + 2];
        float d = dx * dx;
        d = d + dy * dy;
        d = d + dz * dz;
        int i = idx_[k];
        if (d < bestd || (d == bestd && i < best)) { bestd = d; best = i; }
      }
      return;
    }
    float qd = q[n.dim];
    // one-dimensional f32 lower bounds of the f32 distance to anything in each child
    float bl = 0.f, br = 0.f;
    if (qd > n.split_lo) { float t = qd - n.split_lo; bl = t * t; }
    if (qd < n.split_hi) { float t = n.split_hi - qd; br = t * t; }
    int first = n.left, second = n.right;
    float bsecond = br;
    float bfirst = bl;
    if (br < bl) { first = n.right; second = n.left; bsecond = bl; bfirst = br; }
    if (!(bfirst > bestd)) search(first, q, best, bestd);
    if (!(bsecond > bestd)) search(second, q, best, bestd);
  }
};

// ---------------------------------------------------------------------------------------------------------
// The estimator: the part of fusion::StateEstimator on the hot path.
// ---------------------------------------------------------------------------------------------------------
struct Params {
  int num_iter = 30;                // NUM_ITER
  int icp_freq = 1;                 // ICP_FREQ
  double nearest_sq_dist = 25.0;    // NEAREST_FEATURE_SEARCH_SQ_DIST
  double lidar_std = 0.01;          // LIDAR_STD
  double lidar_scale = 1.0;         // LIDAR_SCALE
  double scan_period = 0.1;         // SCAN_PERIOD
  int force_all_iters = 0;          // test hook (lins_params::force_all_iters)
  static Params from_c(const lins_params& p) {
    Params r;
    r.num_iter = p.num_iter; r.icp_freq = p.icp_freq; r.nearest_sq_dist = p.nearest_feature_search_sq_dist;
    r.lidar_std = p.lidar_std; r.lidar_scale = p.lidar_scale; r.scan_period = p.scan_period;
    r.force_all_iters = p.force_all_iters;
    return r;
  }
};

enum GainForm { FORM_A_REFERENCE = 0, FORM_B_INFORMATION = 1 };

struct Report {
  int iters = 0;
  bool converged = false, diverged = false, has_nan = false;
  std::vector<int> m_surf, m_corner;
  std::vector<double> residual_norm, update_norm;
  // optional trace of every iteration's association (dense per query), for parity tests
  bool keep_trace = false;
  std::vector<std::vector<int>> surf_ind, corner_ind;
  std::vector<std::vector<uint8_t>> surf_mask, corner_mask;
  std::vector<GlobalState> lin_states;  // linearisation point at the START of each iteration
};

class Estimator {
 public:
  Params prm;
  bool use_kdtree = true;  // false: brute-force NN (ground truth)

  // scan_last_ feature clouds (walks + tripods index THESE, StateEstimator.hpp:841-842, :967-968)
  std::vector<PointType> lastSurf, lastCorner;
  // the clouds the kd-trees were built on (may lag behind scan_last_: the >=5 && >=20 guard, :1156-1157)
  std::vector<PointType> treeSurf, treeCorner;
  KdTree kdSurf, kdCorner;

  // scan_new_ query features
  std::vector<PointType> surfFlat, cornerSharp;

  // pointSearch*Ind persist across iterations and scans (StateEstimator.hpp:1459-1465)
  std::vector<int> surfInd1, surfInd2, surfInd3, cornerInd1, cornerInd2;

  GlobalState linState_;
  Vec18 updateVec_, difVecLinInv_;

  // per-call outputs of the find* functions (keypoints_/jacobianCoff*)
  std::vector<PointType> keypointSurfs, coeffSurfs, keypointCorns, coeffCorns;
  // dense per-query records of the last find* call (parity hooks)
  std::vector<float> surfSel, cornerSel, surfCoeffDense, cornerCoeffDense;
  std::vector<uint8_t> surfMask, cornerMask;

  // ≙ kdtreeCorner_/kdtreeSurf_->setInputCloud + scan_last_ assignment (processFirstScan, :363-371)
  void setMap(const PointType* surf, int ns, const PointType* corner, int nc) {
    lastSurf.assign(surf, surf + ns);
    lastCorner.assign(corner, corner + nc);
    treeSurf = lastSurf; treeCorner = lastCorner;
    kdSurf.build(treeSurf); kdCorner.build(treeCorner);
  }
  void setQueries(const PointType* surf, int ns, const PointType* corner, int nc) {
    surfFlat.assign(surf, surf + ns);
    cornerSharp.assign(corner, corner + nc);
    if ((int)surfInd1.size() < ns) { surfInd1.resize(ns, -1); surfInd2.resize(ns, -1); surfInd3.resize(ns, -1); }
    if ((int)cornerInd1.size() < nc) { cornerInd1.resize(nc, -1); cornerInd2.resize(nc, -1); }
  }

  // StateEstimator.hpp:1066-1080 transformToStart
  void transformToStart(const PointType* pi, PointType* po) const {
    double s = (1.f / prm.scan_period) * (pi->intensity - int(pi->intensity));
    V3 P2xyz(pi->x, pi->y, pi->z);
    V3 phi = Quat2axis(linState_.qbn_);
    Q4 R21xyz = axis2Quat(s * phi);
    // R21xyz.normalized();  -- result discarded in the reference (:1072): a no-op, kept a no-op
    V3 T112xyz = s * linState_.rn_;
    V3 P1xyz = (R21xyz * P2xyz) + T112xyz;
    po->x = (float)P1xyz.x; po->y = (float)P1xyz.y; po->z = (float)P1xyz.z;
    po->intensity = pi->intensity;
  }
  // StateEstimator.hpp:1083-1101 transformToEnd
  void transformToEnd(const PointType* pi, PointType* po) const {
    double s = (1.f / prm.scan_period) * (pi->intensity - int(pi->intensity));
    V3 P2xyz(pi->x, pi->y, pi->z);
    V3 phi = Quat2axis(linState_.qbn_);
    Q4 R21xyz = axis2Quat(s * phi);
    V3 T112xyz = s * linState_.rn_;
    V3 P1xyz = (R21xyz * P2xyz) + T112xyz;
    R21xyz = linState_.qbn_;
    T112xyz = linState_.rn_;
    P2xyz = inverse(R21xyz) * (P1xyz - T112xyz);
    float inten = pi->intensity;
    po->x = (float)P2xyz.x; po->y = (float)P2xyz.y; po->z = (float)P2xyz.z;
    po->intensity = inten;
  }

  int nnSurf(const PointType& p, float* sq) const {
    return use_kdtree ? kdSurf.nearest(p.x, p.y, p.z, sq) : nn_brute(treeSurf, p.x, p.y, p.z, sq);
  }
  int nnCorner(const PointType& p, float* sq) const {
    return use_kdtree ? kdCorner.nearest(p.x, p.y, p.z, sq) : nn_brute(treeCorner, p.x, p.y, p.z, sq);
  }

  // StateEstimator.hpp:829-953 findCorrespondingSurfFeatures
  void findCorrespondingSurfFeatures(int iterCount) {
    const int surfPointsFlatNum = (int)surfFlat.size();
    const std::vector<PointType>& laserCloudSurfLast = lastSurf;
    const int T = (int)laserCloudSurfLast.size();
    keypointSurfs.clear(); coeffSurfs.clear();
    surfSel.assign(3 * (size_t)surfPointsFlatNum, 0.f);
    surfCoeffDense.assign(4 * (size_t)surfPointsFlatNum, 0.f);
    surfMask.assign(surfPointsFlatNum, 0);
    const float NEAR = (float)prm.nearest_sq_dist;
    for (int i = 0; i < surfPointsFlatNum; i++) {
      PointType pointSel, coeff;
      std::memset(&coeff, 0, sizeof(coeff));
      std::memset(&pointSel, 0, sizeof(pointSel));
      transformToStart(&surfFlat[i], &pointSel);
      surfSel[3 * i] = pointSel.x; surfSel[3 * i + 1] = pointSel.y; surfSel[3 * i + 2] = pointSel.z;

      if (iterCount % prm.icp_freq == 0) {
        float sq = 0.f;
        int nn = nnSurf(pointSel, &sq);
        int closestPointInd = -1, minPointInd2 = -1, minPointInd3 = -1;
        // (double)sq < NEAREST_FEATURE_SEARCH_SQ_DIST ; nn >= T guards the stale-tree quirk (UB in the reference)
        if (nn >= 0 && nn < T && (double)sq < prm.nearest_sq_dist) {
          closestPointInd = nn;
          int closestPointScan = int(laserCloudSurfLast[closestPointInd].intensity);
          float pointSqDis, minPointSqDis2 = NEAR, minPointSqDis3 = NEAR;
          const int fwdBound = std::min(surfPointsFlatNum, T);  // :859 quirk (query count) + OOB clamp
          for (int j = closestPointInd + 1; j < fwdBound; j++) {
            if (int(laserCloudSurfLast[j].intensity) > closestPointScan + 2.5) break;
            pointSqDis = (laserCloudSurfLast[j].x - pointSel.x) * (laserCloudSurfLast[j].x - pointSel.x) +
                         (laserCloudSurfLast[j].y - pointSel.y) * (laserCloudSurfLast[j].y - pointSel.y) +
                         (laserCloudSurfLast[j].z - pointSel.z) * (laserCloudSurfLast[j].z - pointSel.z);
            if (int(laserCloudSurfLast[j].intensity) <= closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            } else {
              if (pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
            }
          }
          for (int j = closestPointInd - 1; j >= 0; j--) {
            if (int(laserCloudSurfLast[j].intensity) < closestPointScan - 2.5) break;
            pointSqDis = (laserCloudSurfLast[j].x - pointSel.x) * (laserCloudSurfLast[j].x - pointSel.x) +
                         (laserCloudSurfLast[j].y - pointSel.y) * (laserCloudSurfLast[j].y - pointSel.y) +
                         (laserCloudSurfLast[j].z - pointSel.z) * (laserCloudSurfLast[j].z - pointSel.z);
            if (int(laserCloudSurfLast[j].intensity) >= closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            } else {
              if (pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
            }
          }
        }
        surfInd1[i] = closestPointInd; surfInd2[i] = minPointInd2; surfInd3[i] = minPointInd3;
      }

      if (surfInd2[i] >= 0 && surfInd3[i] >= 0 && surfInd1[i] >= 0 && surfInd1[i] < T && surfInd2[i] < T &&
          surfInd3[i] < T) {
        const PointType& tripod1 = laserCloudSurfLast[surfInd1[i]];
        const PointType& tripod2 = laserCloudSurfLast[surfInd2[i]];
        const PointType& tripod3 = laserCloudSurfLast[surfInd3[i]];
        V3 P0xyz(pointSel.x, pointSel.y, pointSel.z);
        V3 P1xyz(tripod1.x, tripod1.y, tripod1.z);
        V3 P2xyz(tripod2.x, tripod2.y, tripod2.z);
        V3 P3xyz(tripod3.x, tripod3.y, tripod3.z);
        V3 M = cross(P1xyz - P2xyz, P1xyz - P3xyz);  // skew(P1-P2)*(P1-P3)
        double r = dot(P0xyz - P1xyz, M);
        double m = norm(M);
        float res = r / m;
        V3 jacxyz = M / m;
        float s = 1;
        if (iterCount >= prm.icp_freq) {
          s = 1 - 1.8 * std::fabs(res) /
                      std::sqrt(std::sqrt(pointSel.x * pointSel.x + pointSel.y * pointSel.y + pointSel.z * pointSel.z));
        }
        if (s > 0.1 && res != 0) {
          coeff.x = s * jacxyz.x; coeff.y = s * jacxyz.y; coeff.z = s * jacxyz.z;
          coeff.intensity = s * res;
          keypointSurfs.push_back(surfFlat[i]);
          coeffSurfs.push_back(coeff);
          surfMask[i] = 1;
          surfCoeffDense[4 * i] = coeff.x; surfCoeffDense[4 * i + 1] = coeff.y;
          surfCoeffDense[4 * i + 2] = coeff.z; surfCoeffDense[4 * i + 3] = coeff.intensity;
        }
      }
    }
  }

  // StateEstimator.hpp:955-1063 findCorrespondingCornerFeatures
  void findCorrespondingCornerFeatures(int iterCount) {
    const int cornerPointsSharpNum = (int)cornerSharp.size();
    const std::vector<PointType>& laserCloudCornerLast = lastCorner;
    const int T = (int)laserCloudCornerLast.size();
    keypointCorns.clear(); coeffCorns.clear();
    cornerSel.assign(3 * (size_t)cornerPointsSharpNum, 0.f);
    cornerCoeffDense.assign(4 * (size_t)cornerPointsSharpNum, 0.f);
    cornerMask.assign(cornerPointsSharpNum, 0);
    const float NEAR = (float)prm.nearest_sq_dist;
    for (int i = 0; i < cornerPointsSharpNum; i++) {
      PointType pointSel, coeff;
      std::memset(&coeff, 0, sizeof(coeff));
      std::memset(&pointSel, 0, sizeof(pointSel));
      transformToStart(&cornerSharp[i], &pointSel);
      cornerSel[3 * i] = pointSel.x; cornerSel[3 * i + 1] = pointSel.y; cornerSel[3 * i + 2] = pointSel.z;

      if (iterCount % prm.icp_freq == 0) {
        float sq = 0.f;
        int nn = nnCorner(pointSel, &sq);
        int closestPointInd = -1, minPointInd2 = -1;
        if (nn >= 0 && nn < T && (double)sq < prm.nearest_sq_dist) {
          closestPointInd = nn;
          int closestPointScan = int(laserCloudCornerLast[closestPointInd].intensity);
          float pointSqDis, minPointSqDis2 = NEAR;
          const int fwdBound = std::min(cornerPointsSharpNum, T);  // :983 quirk + OOB clamp
          for (int j = closestPointInd + 1; j < fwdBound; j++) {
            if (int(laserCloudCornerLast[j].intensity) > closestPointScan + 2.5) break;
            pointSqDis = (laserCloudCornerLast[j].x - pointSel.x) * (laserCloudCornerLast[j].x - pointSel.x) +
                         (laserCloudCornerLast[j].y - pointSel.y) * (laserCloudCornerLast[j].y - pointSel.y) +
                         (laserCloudCornerLast[j].z - pointSel.z) * (laserCloudCornerLast[j].z - pointSel.z);
            if (int(laserCloudCornerLast[j].intensity) > closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
          }
          for (int j = closestPointInd - 1; j >= 0; j--) {
            if (int(laserCloudCornerLast[j].intensity) < closestPointScan - 2.5) break;
            pointSqDis = (laserCloudCornerLast[j].x - pointSel.x) * (laserCloudCornerLast[j].x - pointSel.x) +
                         (laserCloudCornerLast[j].y - pointSel.y) * (laserCloudCornerLast[j].y - pointSel.y) +
                         (laserCloudCornerLast[j].z - pointSel.z) * (laserCloudCornerLast[j].z - pointSel.z);
            if (int(laserCloudCornerLast[j].intensity) < closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
          }
        }
        cornerInd1[i] = closestPointInd; cornerInd2[i] = minPointInd2;
      }

      if (cornerInd2[i] >= 0 && cornerInd1[i] >= 0 && cornerInd1[i] < T && cornerInd2[i] < T) {
        const PointType& tripod1 = laserCloudCornerLast[cornerInd1[i]];
        const PointType& tripod2 = laserCloudCornerLast[cornerInd2[i]];
        V3 P0xyz(pointSel.x, pointSel.y, pointSel.z);
        V3 P1xyz(tripod1.x, tripod1.y, tripod1.z);
        V3 P2xyz(tripod2.x, tripod2.y, tripod2.z);
        V3 P = cross(P0xyz - P1xyz, P0xyz - P2xyz);  // skew(P0-P1)*(P0-P2)
        float r = norm(P);
        float d12 = norm(P1xyz - P2xyz);
        float res = r / d12;
        // P^T * skew(P2-P1) / (d12*r)
        V3 a = P2xyz - P1xyz;
        V3 num(P.y * a.z + P.z * (-a.y), P.x * (-a.z) + P.z * a.x, P.x * a.y + P.y * (-a.x));
        double den = (double)(d12 * r);
        V3 jacxyz = num / den;
        float s = 1;
        if (iterCount >= prm.icp_freq) s = 1 - 1.8 * std::fabs(res);
        if (s > 0.1 && res != 0) {
          coeff.x = s * jacxyz.x; coeff.y = s * jacxyz.y; coeff.z = s * jacxyz.z;
          coeff.intensity = s * res;
          keypointCorns.push_back(cornerSharp[i]);
          coeffCorns.push_back(coeff);
          cornerMask[i] = 1;
          cornerCoeffDense[4 * i] = coeff.x; cornerCoeffDense[4 * i + 1] = coeff.y;
          cornerCoeffDense[4 * i + 2] = coeff.z; cornerCoeffDense[4 * i + 3] = coeff.intensity;
        }
      }
    }
  }

  // Measurement row of one accepted feature, StateEstimator.hpp:516-532: residual + the 6 structural
  // non-zeros of the 1x18 Jacobian row (pos cols 0-2, att cols 6-8).
  struct HCommon { M3 negR; M3 Rinv; };
  HCommon hCommon() const {
    HCommon c;
    V3 axis = Quat2axis(linState_.qbn_);
    c.negR = -toRotationMatrix(linState_.qbn_);
    c.Rinv = Rinvleft(-axis);
    return c;
  }
  void measurementRow(const HCommon& hc, const PointType& kp, const PointType& cf, double* h6, double* r) const {
    V3 P2xyz(kp.x, kp.y, kp.z);
    V3 coff_xyz(cf.x, cf.y, cf.z);
    *r = prm.lidar_scale * cf.intensity;
    M3 N = hc.negR * skew(P2xyz);
    V3 att = rowTimes(rowTimes(coff_xyz, N), hc.Rinv);
    h6[0] = coff_xyz.x; h6[1] = coff_xyz.y; h6[2] = coff_xyz.z;
    h6[3] = att.x; h6[4] = att.y; h6[5] = att.z;
  }

  // StateEstimator.hpp:465-600 performIESKF.  state/P in, state/P out (what filter_->update receives).
  // When diverged, stateOut/POut are the prior (the caller runs estimateTransform, :585-592).
  void performIESKF(const GlobalState& filterStateIn, const Mat18& Pin, GainForm form, GlobalState& stateOut,
                    Mat18& POut, Report& rep) {
    Mat18 Pk = Pin;
    GlobalState filterState = filterStateIn;
    linState_ = filterState;
    double residualNorm = 1e6;
    bool hasConverged = false, hasDiverged = false;
    { bool kt = rep.keep_trace; rep = Report(); rep.keep_trace = kt; }

    // last iteration's gain pieces, needed by the covariance update
    std::vector<double> H;   // M x 18 row-major (form A)
    std::vector<double> K;   // 18 x M row-major (form A)
    double A6[6][6];         // form B:  sum h h^T on the {0,1,2,6,7,8} pattern
    int M = 0;
    const double sig2 = prm.lidar_std * prm.lidar_std;
    static const int col6[6] = {0, 1, 2, 6, 7, 8};

    for (int iter = 0; iter < prm.num_iter && !hasConverged && !hasDiverged; iter++) {
      if (rep.keep_trace) rep.lin_states.push_back(linState_);
      findCorrespondingSurfFeatures(iter);
      findCorrespondingCornerFeatures(iter);
      if (rep.keep_trace) {
        std::vector<int> si(3 * surfFlat.size()), ci(2 * cornerSharp.size());
        for (size_t i = 0; i < surfFlat.size(); ++i) { si[3 * i] = surfInd1[i]; si[3 * i + 1] = surfInd2[i]; si[3 * i + 2] = surfInd3[i]; }
        for (size_t i = 0; i < cornerSharp.size(); ++i) { ci[2 * i] = cornerInd1[i]; ci[2 * i + 1] = cornerInd2[i]; }
        rep.surf_ind.push_back(si); rep.corner_ind.push_back(ci);
        rep.surf_mask.push_back(surfMask); rep.corner_mask.push_back(cornerMask);
      }
      // keypoints_ = surfs ++ corners (:499-504)
      const int Ms = (int)keypointSurfs.size(), Mc = (int)keypointCorns.size();
      M = Ms + Mc;
      rep.m_surf.push_back(Ms); rep.m_corner.push_back(Mc);

      std::vector<double> residual(M);
      H.assign((size_t)M * 18, 0.0);
      HCommon hc = hCommon();
      for (int i = 0; i < M; ++i) {
        const PointType& kp = i < Ms ? keypointSurfs[i] : keypointCorns[i - Ms];
        const PointType& cf = i < Ms ? coeffSurfs[i] : coeffCorns[i - Ms];
        double h6[6], r;
        measurementRow(hc, kp, cf, h6, &r);
        residual[i] = r;
        double* row = &H[(size_t)i * 18];
        row[GlobalState::pos_] = h6[0]; row[GlobalState::pos_ + 1] = h6[1]; row[GlobalState::pos_ + 2] = h6[2];
        row[GlobalState::att_] = h6[3]; row[GlobalState::att_ + 1] = h6[4]; row[GlobalState::att_ + 2] = h6[5];
      }
      double rnorm = 0;
      for (int i = 0; i < M; ++i) rnorm += residual[i] * residual[i];
      rnorm = std::sqrt(rnorm);

      filterState.boxMinus(linState_, difVecLinInv_);  // :548

      Vec18 upd;
      if (form == FORM_A_REFERENCE) {
        gainFormA(H, M, Pk, sig2, K);
        // updateVec_ = -Kk_ * (residual_ + Hk_ * difVecLinInv_) + difVecLinInv_   (:549)
        std::vector<double> y(M);
        for (int i = 0; i < M; ++i) {
          double s = 0;
          for (int j = 0; j < 18; ++j) s += H[(size_t)i * 18 + j] * difVecLinInv_[j];
          y[i] = residual[i] + s;
        }
        for (int a = 0; a < 18; ++a) {
          double s = 0;
          for (int i = 0; i < M; ++i) s += K[(size_t)a * M + i] * y[i];
          upd[a] = -s + difVecLinInv_[a];
        }
      } else {
        // Form B: K(r + H d) = (P A + sig2 I)^-1 P (b + A d),  A = H^T H, b = H^T r
        double b6[6];
        std::memset(A6, 0, sizeof(A6)); std::memset(b6, 0, sizeof(b6));
        for (int i = 0; i < M; ++i) {
          const double* row = &H[(size_t)i * 18];
          double h[6];
          for (int a = 0; a < 6; ++a) h[a] = row[col6[a]];
          for (int a = 0; a < 6; ++a) {
            b6[a] += h[a] * residual[i];
            for (int c = 0; c < 6; ++c) A6[a][c] += h[a] * h[c];
          }
        }
        double y[18];  // b + A d
        std::memset(y, 0, sizeof(y));
        for (int a = 0; a < 6; ++a) {
          double s = b6[a];
          for (int c = 0; c < 6; ++c) s += A6[a][c] * difVecLinInv_[col6[c]];
          y[col6[a]] = s;
        }
        double S[18 * 18], rhs[18];
        formS(Pk, A6, sig2, S);
        for (int a = 0; a < 18; ++a) {
          double s = 0;
          for (int c = 0; c < 18; ++c) s += Pk.m[a][c] * y[c];
          rhs[a] = s;
        }
        bool ok = lu_solve(S, 18, rhs, 1);
        for (int a = 0; a < 18; ++a) upd[a] = ok ? (-rhs[a] + difVecLinInv_[a]) : std::numeric_limits<double>::quiet_NaN();
      }
      updateVec_ = upd;
      rep.iters = iter + 1;
      rep.residual_norm.push_back(rnorm);

      // Divergence determination (:552-570)
      bool hasNaN = false;
      for (int i = 0; i < 18; i++)
        if (std::isnan(updateVec_[i])) { updateVec_[i] = 0; hasNaN = true; }
      if (hasNaN) {
        rep.has_nan = true; hasDiverged = true;
        rep.update_norm.push_back(updateVec_.norm());
        break;
      }
      if (rnorm > residualNorm * 10) {
        hasDiverged = true;
        rep.update_norm.push_back(updateVec_.norm());
        break;
      }
      linState_.boxPlus(updateVec_, linState_);  // :573
      double updateVecNorm = updateVec_.norm();
      rep.update_norm.push_back(updateVecNorm);
      if (updateVecNorm <= 1e-2 && !prm.force_all_iters) hasConverged = true;
      residualNorm = rnorm;
    }
    rep.converged = hasConverged; rep.diverged = hasDiverged;

    if (hasDiverged) {
      stateOut = filterState;  // caller runs estimateTransform and patches rn_/qbn_ (:585-592)
      POut = Pin;
    } else {
      if (rep.iters > 0) {
        if (form == FORM_A_REFERENCE) {
          // IKH_ = I - Kk_*Hk_ ; Pk_ = IKH_*Pk_*IKH_^T + Kk_*Rk_*Kk_^T   (:595-596)
          Mat18 IKH = Mat18::Identity(), KRK;
          for (int a = 0; a < 18; ++a)
            for (int i = 0; i < M; ++i) {
              double k = K[(size_t)a * M + i];
              for (int c = 0; c < 18; ++c) IKH.m[a][c] -= k * H[(size_t)i * 18 + c];
            }
          for (int a = 0; a < 18; ++a)
            for (int c = 0; c < 18; ++c) {
              double s = 0;
              for (int i = 0; i < M; ++i) s += (K[(size_t)a * M + i] * sig2) * K[(size_t)c * M + i];
              KRK.m[a][c] = s;
            }
          Pk = add(mul(mul(IKH, Pk), transpose(IKH)), KRK);
        } else {
          // K H = G A,  K R K^T = sig2 G A G^T,  G = (P A + sig2 I)^-1 P
          double S[18 * 18], G[18 * 18];
          formS(Pk, A6, sig2, S);
          for (int a = 0; a < 18; ++a)
            for (int c = 0; c < 18; ++c) G[a * 18 + c] = Pk.m[a][c];
          bool ok = lu_solve(S, 18, G, 18);
          Mat18 Gm, A18, IKH = Mat18::Identity();
          for (int a = 0; a < 18; ++a)
            for (int c = 0; c < 18; ++c) Gm.m[a][c] = ok ? G[a * 18 + c] : std::numeric_limits<double>::quiet_NaN();
          for (int a = 0; a < 6; ++a)
            for (int c = 0; c < 6; ++c) A18.m[col6[a]][col6[c]] = A6[a][c];
          Mat18 GA = mul(Gm, A18);
          for (int a = 0; a < 18; ++a)
            for (int c = 0; c < 18; ++c) IKH.m[a][c] -= GA.m[a][c];
          Mat18 KRK = mul(GA, transpose(Gm));
          for (int a = 0; a < 18; ++a)
            for (int c = 0; c < 18; ++c) KRK.m[a][c] *= sig2;
          Pk = add(mul(mul(IKH, Pk), transpose(IKH)), KRK);
        }
        enforceSymmetry(Pk);  // :597
      }
      stateOut = linState_;  // filter_->update(linState_, Pk_)  (:598)
      POut = Pk;
    }
  }

  // S = P A + sig2 I  (18x18 row-major), A non-zero only on the {0,1,2,6,7,8} pattern
  static void formS(const Mat18& P, const double A6[6][6], double sig2, double* S) {
    static const int col6[6] = {0, 1, 2, 6, 7, 8};
    std::memset(S, 0, sizeof(double) * 324);
    for (int a = 0; a < 18; ++a) {
      for (int c = 0; c < 6; ++c) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += P.m[a][col6[k]] * A6[k][c];
        S[a * 18 + col6[c]] = s;
      }
      S[a * 18 + a] += sig2;
    }
  }

  // Form A gain, StateEstimator.hpp:535-546:  Py = H P H^T + R (M x M), Pyinv = Py.llt().solve(I),
  // K = P H^T Pyinv.  Dense, like the reference (R stored dense there; adding the zeros is exact).
  static void gainFormA(const std::vector<double>& H, int M, const Mat18& P, double sig2, std::vector<double>& K) {
    K.assign((size_t)18 * M, 0.0);
    if (M == 0) return;
    // PHt = P H^T  (18 x M row-major);  HP = H P (M x 18)
    std::vector<double> HP((size_t)M * 18), PHt((size_t)18 * M);
    for (int i = 0; i < M; ++i)
      for (int c = 0; c < 18; ++c) {
        double s = 0;
        for (int k = 0; k < 18; ++k) s += H[(size_t)i * 18 + k] * P.m[k][c];
        HP[(size_t)i * 18 + c] = s;
      }
    for (int a = 0; a < 18; ++a)
      for (int i = 0; i < M; ++i) {
        double s = 0;
        for (int k = 0; k < 18; ++k) s += P.m[a][k] * H[(size_t)i * 18 + k];
        PHt[(size_t)a * M + i] = s;
      }
    // Py (row-major, full)
    std::vector<double> Py((size_t)M * M);
    for (int i = 0; i < M; ++i) {
      double* row = &Py[(size_t)i * M];
      for (int j = 0; j < M; ++j) {
        double s = 0;
        const double* hp = &HP[(size_t)i * 18];
        const double* hj = &H[(size_t)j * 18];
        for (int k = 0; k < 18; ++k) s += hp[k] * hj[k];
        row[j] = s;
      }
      row[i] += sig2;
    }
    // LLT: lower Cholesky, row-major, L stored in Py's lower triangle.  L[i][j] for j<=i.
    // Right-looking by columns of L would stride; use the row (Cholesky-Banachiewicz) form with
    // 4-way split accumulators so the dot products are not latency bound.
    bool ok = true;
    for (int i = 0; i < M && ok; ++i) {
      double* Li = &Py[(size_t)i * M];
      for (int j = 0; j <= i; ++j) {
        const double* Lj = &Py[(size_t)j * M];
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int k = 0;
        for (; k + 3 < j; k += 4) {
          s0 += Li[k] * Lj[k]; s1 += Li[k + 1] * Lj[k + 1]; s2 += Li[k + 2] * Lj[k + 2]; s3 += Li[k + 3] * Lj[k + 3];
        }
        for (; k < j; ++k) s0 += Li[k] * Lj[k];
        double s = Li[j] - ((s0 + s1) + (s2 + s3));
        if (i == j) {
          if (!(s > 0)) { ok = false; break; }  // Eigen's LLT would report NumericalIssue and carry NaNs on
          Li[j] = std::sqrt(s);
        } else {
          Li[j] = s / Lj[j];
        }
      }
    }
    if (!ok) {  // propagate NaN like a failed factorisation would
      for (auto& k : K) k = std::numeric_limits<double>::quiet_NaN();
      return;
    }
    // Pyinv: solve L Y = I, then L^T X = Y, on a dense identity (the reference's solveInPlace does the full
    // 2 M^3 flops: it does not exploit the identity's structure).  X row-major; rows are the RHS index so the
    // inner loops run over contiguous memory (axpy form, vectorisable without reassociation).
    std::vector<double> X((size_t)M * M, 0.0);
    for (int i = 0; i < M; ++i) X[(size_t)i * M + i] = 1.0;
    // forward: for each row k of the unknown (all RHS at once): Y[k,:] = (B[k,:] - sum_{j<k} L[k][j] Y[j,:]) / L[k][k]
    for (int k = 0; k < M; ++k) {
      double* Yk = &X[(size_t)k * M];
      const double* Lk = &Py[(size_t)k * M];
      for (int j = 0; j < k; ++j) {
        double l = Lk[j];
        const double* Yj = &X[(size_t)j * M];
        for (int c = 0; c < M; ++c) Yk[c] -= l * Yj[c];
      }
      double inv = 1.0 / Lk[k];
      for (int c = 0; c < M; ++c) Yk[c] *= inv;
    }
    // backward: X[k,:] = (Y[k,:] - sum_{j>k} L[j][k] X[j,:]) / L[k][k]
    for (int k = M - 1; k >= 0; --k) {
      double* Xk = &X[(size_t)k * M];
      for (int j = k + 1; j < M; ++j) {
        double l = Py[(size_t)j * M + k];
        const double* Xj = &X[(size_t)j * M];
        for (int c = 0; c < M; ++c) Xk[c] -= l * Xj[c];
      }
      double inv = 1.0 / Py[(size_t)k * M + k];
      for (int c = 0; c < M; ++c) Xk[c] *= inv;
    }
    // K = PHt * Pyinv   (18 x M)
    for (int a = 0; a < 18; ++a) {
      double* Ka = &K[(size_t)a * M];
      for (int i = 0; i < M; ++i) {
        double p = PHt[(size_t)a * M + i];
        const double* Xi = &X[(size_t)i * M];
        for (int c = 0; c < M; ++c) Ka[c] += p * Xi[c];
      }
    }
  }

  // -------------------------------------------------------------------------------------------------------
  // Fallback / scan-2 initialiser: StateEstimator.hpp:1163-1196 estimateTransform
  // -------------------------------------------------------------------------------------------------------
  bool estimateTransform(V3& t, Q4& q, int* iters_out = nullptr) {
    linState_.rn_ = t;
    linState_.qbn_ = q;
    bool conv = false;
    int it = 0;
    for (int iter = 0; iter < prm.num_iter; iter++) {
      it = iter + 1;
      findCorrespondingSurfFeatures(iter);
      if (keypointSurfs.size() < 10) continue;
      findCorrespondingCornerFeatures(iter);
      if (keypointCorns.size() < 5) continue;
      if (calculateTransformation(iter)) { conv = true; break; }
    }
    t = linState_.rn_;
    q = linState_.qbn_;
    if (iters_out) *iters_out = it;
    return conv;
  }

  // StateEstimator.hpp:1198-1320 calculateTransformation
  bool calculateTransformation(int iterCount) {
    const int Ms = (int)keypointSurfs.size(), Mc = (int)keypointCorns.size();
    const int pointNum = Ms + Mc;
    double JTJ[36], JTb[6], x[6];
    std::memset(JTJ, 0, sizeof(JTJ)); std::memset(JTb, 0, sizeof(JTb)); std::memset(x, 0, sizeof(x));
    for (int i = 0; i < pointNum; ++i) {
      const PointType& keypoint = i < Ms ? keypointSurfs[i] : keypointCorns[i - Ms];
      const PointType& coeff = i < Ms ? coeffSurfs[i] : coeffCorns[i - Ms];
      V3 P2xyz(keypoint.x, keypoint.y, keypoint.z);
      V3 coff_xyz(coeff.x, coeff.y, coeff.z);
      double s = (1.f / prm.scan_period) * (keypoint.intensity - int(keypoint.intensity));
      V3 phi = Quat2axis(linState_.qbn_);
      Q4 R21xyz = axis2Quat(s * phi);
      M3 N = (-toRotationMatrix(R21xyz)) * skew(P2xyz);
      V3 jacobian1xyz = rowTimes(coff_xyz, N);
      V3 jacobian2xyz = coff_xyz;
      double residual = coeff.intensity;
      double J[6] = {jacobian1xyz.x, jacobian1xyz.y, jacobian1xyz.z, jacobian2xyz.x, jacobian2xyz.y, jacobian2xyz.z};
      double b = -0.05 * residual;
      for (int a = 0; a < 6; ++a) {
        JTb[a] += J[a] * b;
        for (int c = 0; c < 6; ++c) JTJ[a * 6 + c] += J[a] * J[c];
      }
    }
    colPivHouseholderQrSolve6(JTJ, JTb, x);

    bool isDegenerate = false;
    double matP[36];
    if (iterCount == 0) {
      double matE[6], matV[36], matV2[36];
      symEig6(JTJ, matE, matV);  // ascending eigenvalues, eigenvectors in COLUMNS (Eigen convention)
      std::memcpy(matV2, matV, sizeof(matV));
      for (int i = 0; i < 6; i++) {
        if (matE[i] < 10.) {
          for (int j = 0; j < 6; j++) matV2[i * 6 + j] = 0;  // ROW i zeroed, as the reference does (:1287-1289)
          isDegenerate = true;
        } else {
          break;
        }
      }
      double Vinv[36];
      inv6(matV, Vinv);
      for (int a = 0; a < 6; ++a)
        for (int c = 0; c < 6; ++c) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += Vinv[a * 6 + k] * matV2[k * 6 + c];
          matP[a * 6 + c] = s;
        }
    }
    if (isDegenerate) {
      double x2[6];
      std::memcpy(x2, x, sizeof(x2));
      for (int a = 0; a < 6; ++a) {
        double s = 0;
        for (int c = 0; c < 6; ++c) s += matP[a * 6 + c] * x2[c];
        x[a] = s;
      }
    }
    Q4 dq = rpy2Quat(V3(x[0], x[1], x[2]));
    linState_.qbn_ = normalized(linState_.qbn_ * dq);
    linState_.rn_ = linState_.rn_ + V3(x[3], x[4], x[5]);
    V3 rpy_deg(x[0] * 180.0 / M_PI, x[1] * 180.0 / M_PI, x[2] * 180.0 / M_PI);
    double deltaR = norm(rpy_deg);
    V3 trans(100 * x[3], 100 * x[4], 100 * x[5]);
    double deltaT = norm(trans);
    return deltaR < 0.1 && deltaT < 0.1;
  }

  // Householder QR with column pivoting, 6x6, following Eigen::ColPivHouseholderQR::solve semantics
  // (rank from |R_ii| > eps*6*max|R_ii|; solution components beyond the rank set to 0).
  static void colPivHouseholderQrSolve6(const double* Ain, const double* bin, double* x) {
    const int n = 6;
    double A[36], b[6];
    std::memcpy(A, Ain, sizeof(A)); std::memcpy(b, bin, sizeof(b));
    int perm[6];
    for (int i = 0; i < n; ++i) perm[i] = i;
    double colnorm2[6];
    for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += A[i * n + j] * A[i * n + j]; colnorm2[j] = s; }
    double maxpivot = 0;
    int nonzero = n;
    for (int k = 0; k < n; ++k) {
      int piv = k; double best = -1;
      for (int j = k; j < n; ++j) {
        double s = 0; for (int i = k; i < n; ++i) s += A[i * n + j] * A[i * n + j];
        colnorm2[j] = s;
        if (s > best) { best = s; piv = j; }
      }
      if (piv != k) {
        for (int i = 0; i < n; ++i) std::swap(A[i * n + k], A[i * n + piv]);
        std::swap(perm[k], perm[piv]);
      }
      // Householder on column k, rows k..n-1
      double normx = 0; for (int i = k; i < n; ++i) normx += A[i * n + k] * A[i * n + k];
      normx = std::sqrt(normx);
      if (normx == 0.0) { if (nonzero == n) nonzero = k; continue; }
      double alpha = A[k * n + k] >= 0 ? -normx : normx;
      double v[6] = {0, 0, 0, 0, 0, 0};
      for (int i = k; i < n; ++i) v[i] = A[i * n + k];
      v[k] -= alpha;
      double vnorm2 = 0; for (int i = k; i < n; ++i) vnorm2 += v[i] * v[i];
      if (vnorm2 > 0) {
        for (int j = k; j < n; ++j) {
          double s = 0; for (int i = k; i < n; ++i) s += v[i] * A[i * n + j];
          s = 2.0 * s / vnorm2;
          for (int i = k; i < n; ++i) A[i * n + j] -= s * v[i];
        }
        double s = 0; for (int i = k; i < n; ++i) s += v[i] * b[i];
        s = 2.0 * s / vnorm2;
        for (int i = k; i < n; ++i) b[i] -= s * v[i];
      }
      if (std::fabs(A[k * n + k]) > maxpivot) maxpivot = std::fabs(A[k * n + k]);
    }
    double thr = std::numeric_limits<double>::epsilon() * n * maxpivot;
    int rank = 0;
    for (int k = 0; k < n; ++k) if (std::fabs(A[k * n + k]) > thr) ++rank;
    if (nonzero < rank) rank = nonzero;
    double y[6] = {0, 0, 0, 0, 0, 0};
    for (int k = rank - 1; k >= 0; --k) {
      double s = b[k];
      for (int j = k + 1; j < rank; ++j) s -= A[k * n + j] * y[j];
      y[k] = s / A[k * n + k];
    }
    for (int k = 0; k < n; ++k) x[perm[k]] = y[k];
  }

  // cyclic Jacobi eigen-decomposition of a symmetric 6x6; eigenvalues ascending; eigenvector k in COLUMN k,
  // sign-normalised so the largest-magnitude component is positive (documented convention: Eigen's signs
  // are not reproducible without its tridiagonal QL, and the reference's row-zeroing makes matP depend on them).
  static void symEig6(const double* Ain, double* E, double* V) {
    const int n = 6;
    double A[36];
    std::memcpy(A, Ain, sizeof(A));
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 100; ++sweep) {
      double off = 0;
      for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
      if (off < 1e-300) break;
      for (int p = 0; p < n; ++p)
        for (int q = p + 1; q < n; ++q) {
          double apq = A[p * n + q];
          if (apq == 0.0) continue;
          double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
          double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
          double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
          for (int k = 0; k < n; ++k) {
            double akp = A[k * n + p], akq = A[k * n + q];
            A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
          }
          for (int k = 0; k < n; ++k) {
            double apk = A[p * n + k], aqk = A[q * n + k];
            A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
          }
          for (int k = 0; k < n; ++k) {
            double vkp = V[k * n + p], vkq = V[k * n + q];
            V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
          }
        }
    }
    int order[6];
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order, order + n, [&](int a, int b) { return A[a * n + a] < A[b * n + b]; });
    double Vs[36];
    for (int k = 0; k < n; ++k) {
      E[k] = A[order[k] * n + order[k]];
      int big = 0;
      for (int i = 1; i < n; ++i) if (std::fabs(V[i * n + order[k]]) > std::fabs(V[big * n + order[k]])) big = i;
      double sg = V[big * n + order[k]] < 0 ? -1.0 : 1.0;
      for (int i = 0; i < n; ++i) Vs[i * n + k] = sg * V[i * n + order[k]];
    }
    std::memcpy(V, Vs, sizeof(Vs));
  }
  static void inv6(const double* Ain, double* inv) {
    double A[36], B[36];
    std::memcpy(A, Ain, sizeof(A));
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) B[i * 6 + j] = (i == j);
    if (!lu_solve(A, 6, B, 6)) for (int i = 0; i < 36; ++i) B[i] = std::numeric_limits<double>::quiet_NaN();
    std::memcpy(inv, B, sizeof(B));
  }

  // -------------------------------------------------------------------------------------------------------
  // F1: StateEstimator.hpp:1116-1161 updatePointCloud (the XYZ part + the index-rebuild guard).
  // Transforms the new scan's less-* clouds to the scan-end frame in place with linState_ = lin, makes them
  // scan_last_, and rebuilds the kd-trees only if corner>=5 && surf>=20.
  // -------------------------------------------------------------------------------------------------------
  bool updatePointCloud(std::vector<PointType>& surfLess, std::vector<PointType>& cornerLess, const GlobalState& lin) {
    linState_ = lin;
    for (size_t i = 0; i < cornerLess.size(); i++) transformToEnd(&cornerLess[i], &cornerLess[i]);
    for (size_t i = 0; i < surfLess.size(); i++) transformToEnd(&surfLess[i], &surfLess[i]);
    bool rebuilt = false;
    if (cornerLess.size() >= 5 && surfLess.size() >= 20) {
      treeCorner = cornerLess; treeSurf = surfLess;
      kdCorner.build(treeCorner); kdSurf.build(treeSurf);
      rebuilt = true;
    }
    lastSurf = surfLess; lastCorner = cornerLess;  // scan_last_.swap(scan_new_) (:459)
    return rebuilt;
  }
};

}  // namespace lins_oracle
#endif
