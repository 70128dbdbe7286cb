// oracle/lins_oracle_capi.cpp — TEST INFRASTRUCTURE (CPU oracle), not product code.
// C entry points over lins_oracle.hpp so tests/ (ctypes) can run the CPU restatement next to the CUDA path
// with the same argument conventions as include/lins_gpu.h.  PARITY UNPINNED (see lins_oracle.hpp).
#include <atomic>
#include <chrono>
#include <thread>

#include "lins_oracle.hpp"

using namespace lins_oracle;

namespace {
struct Handle {
  Estimator est;
};
void fill_report(const Report& r, lins_report* out) {
  if (!out) return;
  std::memset(out, 0, sizeof(*out));
  out->iters = r.iters; out->converged = r.converged; out->diverged = r.diverged; out->has_nan = r.has_nan;
  for (int i = 0; i < r.iters && i < LINS_MAX_ITER; ++i) {
    out->m_surf[i] = r.m_surf[i]; out->m_corner[i] = r.m_corner[i];
    out->residual_norm[i] = r.residual_norm[i];
    out->update_norm[i] = i < (int)r.update_norm.size() ? r.update_norm[i] : 0.0;
  }
}
}  // namespace

extern "C" {

void* lins_oracle_create(const lins_params* p, int use_kdtree) {
  Handle* h = new Handle();
  h->est.prm = Params::from_c(*p);
  h->est.use_kdtree = use_kdtree != 0;
  return h;
}
void lins_oracle_destroy(void* hv) { delete static_cast<Handle*>(hv); }

int lins_oracle_set_map(void* hv, const lins_point* surf, int ns, const lins_point* corner, int nc) {
  static_cast<Handle*>(hv)->est.setMap(surf, ns, corner, nc);
  return 0;
}

// form: 0 = reference-faithful M x M gain (form A), 1 = 18 x 18 information form (form B)
int lins_oracle_ieskf(void* hv, const lins_point* surf_flat, int ns, const lins_point* corner_sharp, int nc,
                      const double* state_in, const double* cov_in, int form, double* state_out, double* cov_out,
                      lins_report* rep) {
  Estimator& e = static_cast<Handle*>(hv)->est;
  e.setQueries(surf_flat, ns, corner_sharp, nc);
  GlobalState s = GlobalState::from_array(state_in), so;
  Mat18 P, Po;
  cov_from_colmajor(cov_in, P);
  Report r;
  e.performIESKF(s, P, form == 0 ? FORM_A_REFERENCE : FORM_B_INFORMATION, so, Po, r);
  so.to_array(state_out);
  cov_to_colmajor(Po, cov_out);
  fill_report(r, rep);
  return 0;
}

// As above plus the per-iteration association trace.  Buffers are [num_iter][...] dense, caller-allocated:
// surf_ind_tr [num_iter][3*ns], corner_ind_tr [num_iter][2*nc], surf_mask_tr [num_iter][ns],
// corner_mask_tr [num_iter][nc], lin_state_tr [num_iter][19]
int lins_oracle_ieskf_trace(void* hv, const lins_point* surf_flat, int ns, const lins_point* corner_sharp, int nc,
                            const double* state_in, const double* cov_in, int form, double* state_out,
                            double* cov_out, lins_report* rep, int32_t* surf_ind_tr, int32_t* corner_ind_tr,
                            uint8_t* surf_mask_tr, uint8_t* corner_mask_tr, double* lin_state_tr) {
  Estimator& e = static_cast<Handle*>(hv)->est;
  e.setQueries(surf_flat, ns, corner_sharp, nc);
  GlobalState s = GlobalState::from_array(state_in), so;
  Mat18 P, Po;
  cov_from_colmajor(cov_in, P);
  Report r;
  r.keep_trace = true;
  e.performIESKF(s, P, form == 0 ? FORM_A_REFERENCE : FORM_B_INFORMATION, so, Po, r);
  so.to_array(state_out);
  cov_to_colmajor(Po, cov_out);
  fill_report(r, rep);
  for (int it = 0; it < r.iters; ++it) {
    if (surf_ind_tr) std::memcpy(surf_ind_tr + (size_t)it * 3 * ns, r.surf_ind[it].data(), sizeof(int32_t) * 3 * ns);
    if (corner_ind_tr) std::memcpy(corner_ind_tr + (size_t)it * 2 * nc, r.corner_ind[it].data(), sizeof(int32_t) * 2 * nc);
    if (surf_mask_tr) std::memcpy(surf_mask_tr + (size_t)it * ns, r.surf_mask[it].data(), ns);
    if (corner_mask_tr) std::memcpy(corner_mask_tr + (size_t)it * nc, r.corner_mask[it].data(), nc);
    if (lin_state_tr) r.lin_states[it].to_array(lin_state_tr + (size_t)it * 19);
  }
  return 0;
}

int lins_oracle_associate(void* hv, const lins_point* surf_flat, int ns, const lins_point* corner_sharp, int nc,
                          const double* lin_state, int iter, int32_t* surf_ind, int32_t* corner_ind, float* surf_coeff,
                          float* corner_coeff, uint8_t* surf_mask, uint8_t* corner_mask, float* surf_sel,
                          float* corner_sel) {
  Estimator& e = static_cast<Handle*>(hv)->est;
  e.setQueries(surf_flat, ns, corner_sharp, nc);
  e.linState_ = GlobalState::from_array(lin_state);
  e.findCorrespondingSurfFeatures(iter);
  e.findCorrespondingCornerFeatures(iter);
  for (int i = 0; i < ns; ++i) {
    if (surf_ind) { surf_ind[3 * i] = e.surfInd1[i]; surf_ind[3 * i + 1] = e.surfInd2[i]; surf_ind[3 * i + 2] = e.surfInd3[i]; }
  }
  for (int i = 0; i < nc; ++i) {
    if (corner_ind) { corner_ind[2 * i] = e.cornerInd1[i]; corner_ind[2 * i + 1] = e.cornerInd2[i]; }
  }
  if (surf_coeff && ns) std::memcpy(surf_coeff, e.surfCoeffDense.data(), sizeof(float) * 4 * ns);
  if (corner_coeff && nc) std::memcpy(corner_coeff, e.cornerCoeffDense.data(), sizeof(float) * 4 * nc);
  if (surf_mask && ns) std::memcpy(surf_mask, e.surfMask.data(), ns);
  if (corner_mask && nc) std::memcpy(corner_mask, e.cornerMask.data(), nc);
  if (surf_sel && ns) std::memcpy(surf_sel, e.surfSel.data(), sizeof(float) * 3 * ns);
  if (corner_sel && nc) std::memcpy(corner_sel, e.cornerSel.data(), sizeof(float) * 3 * nc);
  return 0;
}

int lins_oracle_estimate_transform(void* hv, const lins_point* surf_flat, int ns, const lins_point* corner_sharp,
                                   int nc, double* pose_io, int* iters_out, int* converged_out) {
  Estimator& e = static_cast<Handle*>(hv)->est;
  e.setQueries(surf_flat, ns, corner_sharp, nc);
  V3 t(pose_io[0], pose_io[1], pose_io[2]);
  Q4 q(pose_io[6], pose_io[3], pose_io[4], pose_io[5]);
  int it = 0;
  bool conv = e.estimateTransform(t, q, &it);
  pose_io[0] = t.x; pose_io[1] = t.y; pose_io[2] = t.z;
  pose_io[3] = q.x; pose_io[4] = q.y; pose_io[5] = q.z; pose_io[6] = q.w;
  if (iters_out) *iters_out = it;
  if (converged_out) *converged_out = conv;
  return 0;
}

int lins_oracle_update_map(void* hv, lins_point* surf_less, int ns, lins_point* corner_less, int nc,
                           const double* lin_state, int* map_replaced) {
  Estimator& e = static_cast<Handle*>(hv)->est;
  std::vector<PointType> s(surf_less, surf_less + ns), c(corner_less, corner_less + nc);
  bool r = e.updatePointCloud(s, c, GlobalState::from_array(lin_state));
  if (ns) std::memcpy(surf_less, s.data(), sizeof(lins_point) * ns);
  if (nc) std::memcpy(corner_less, c.data(), sizeof(lins_point) * nc);
  if (map_replaced) *map_replaced = r;
  return 0;
}

// 1-NN probe on the current surf (which=0) / corner (which=1) tree cloud: kd-tree vs brute force.
int lins_oracle_nn(void* hv, int which, const float* xyz, int n, int use_kdtree, int32_t* idx, float* sq) {
  Estimator& e = static_cast<Handle*>(hv)->est;
  for (int i = 0; i < n; ++i) {
    float d;
    int j;
    if (use_kdtree) j = (which == 0 ? e.kdSurf : e.kdCorner).nearest(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &d);
    else j = nn_brute(which == 0 ? e.treeSurf : e.treeCorner, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &d);
    idx[i] = j; sq[i] = d;
  }
  return 0;
}

// The batched CPU baseline: every scan is an independent unit; n_threads std::threads pull
// scans from a shared counter.  Per scan: setMap (kd-tree build, the reference's setInputCloud) + performIESKF.
// Returns wall seconds in *seconds_out and the total iterations executed in *iters_out.
int lins_oracle_ieskf_batch(const lins_params* p, const lins_batch_desc* b, int first, int count, int form,
                            int use_kdtree, int n_threads, double* state_out, double* cov_out,
                            lins_scan_result* results, double* seconds_out, int64_t* iters_out) {
  if (b->point_format != LINS_POINTS_XYZI32) return -1;  // the oracle reads pcl::PointXYZI records only
  if (n_threads < 1) n_threads = 1;
  std::atomic<int> next(0);
  std::atomic<long long> total_iters(0);
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {
    Estimator e;
    e.prm = Params::from_c(*p);
    e.use_kdtree = use_kdtree != 0;
    for (;;) {
      int k = next.fetch_add(1);
      if (k >= count) break;
      int i = first + k;
      e.setMap(b->surf_less_flat + b->surf_less_flat_off[i], b->surf_less_flat_off[i + 1] - b->surf_less_flat_off[i],
               b->corner_less_sharp + b->corner_less_sharp_off[i],
               b->corner_less_sharp_off[i + 1] - b->corner_less_sharp_off[i]);
      e.setQueries(b->surf_flat + b->surf_flat_off[i], b->surf_flat_off[i + 1] - b->surf_flat_off[i],
                   b->corner_sharp + b->corner_sharp_off[i], b->corner_sharp_off[i + 1] - b->corner_sharp_off[i]);
      GlobalState s = GlobalState::from_array(b->state_in + (size_t)i * 19), so;
      Mat18 P, Po;
      cov_from_colmajor(b->cov_in + (size_t)i * 324, P);
      Report r;
      e.performIESKF(s, P, form == 0 ? FORM_A_REFERENCE : FORM_B_INFORMATION, so, Po, r);
      total_iters += r.iters;
      if (state_out) so.to_array(state_out + (size_t)k * 19);
      if (cov_out) cov_to_colmajor(Po, cov_out + (size_t)k * 324);
      if (results) {
        lins_scan_result& o = results[k];
        std::memset(&o, 0, sizeof(o));
        o.scan_id = i; o.iters = (uint16_t)r.iters;
        o.flags = (uint16_t)((r.converged ? 1u : 0u) | (r.diverged ? 2u : 0u) | (r.has_nan ? 4u : 0u));
        o.pose[0] = so.rn_.x; o.pose[1] = so.rn_.y; o.pose[2] = so.rn_.z;
        o.pose[3] = so.qbn_.x; o.pose[4] = so.qbn_.y; o.pose[5] = so.qbn_.z; o.pose[6] = so.qbn_.w;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
  if (iters_out) *iters_out = total_iters.load();
  return 0;
}

// small-matrix helpers exposed for self tests
void lins_oracle_sym_eig6(const double* A, double* E, double* V) { Estimator::symEig6(A, E, V); }
void lins_oracle_qr_solve6(const double* A, const double* b, double* x) { Estimator::colPivHouseholderQrSolve6(A, b, x); }
void lins_oracle_boxplus(const double* state, const double* dx, double* out) {
  GlobalState s = GlobalState::from_array(state), o;
  Vec18 v;
  for (int i = 0; i < 18; ++i) v[i] = dx[i];
  s.boxPlus(v, o);
  o.to_array(out);
}
void lins_oracle_boxminus(const double* a, const double* b, double* dx) {
  GlobalState sa = GlobalState::from_array(a), sb = GlobalState::from_array(b);
  Vec18 v;
  sa.boxMinus(sb, v);
  for (int i = 0; i < 18; ++i) dx[i] = v[i];
}
// form A gain alone (StateEstimator.hpp:542-546): H is M x 18 row-major, P 18 x 18 row-major, K out 18 x M row-major —
// exposed so that tests can pin it against an independent Cholesky (scipy / LAPACK)
void lins_oracle_gain_form_a(const double* H, int M, const double* P, double sig2, double* K) {
  std::vector<double> h(H, H + (size_t)M * 18), k;
  Mat18 Pm;
  for (int a = 0; a < 18; ++a) for (int b = 0; b < 18; ++b) Pm.m[a][b] = P[a * 18 + b];
  Estimator::gainFormA(h, M, Pm, sig2, k);
  std::memcpy(K, k.data(), sizeof(double) * 18 * (size_t)M);
}
// de-skew helpers: mode 0 = transformToStart, 1 = transformToEnd
void lins_oracle_transform(const lins_params* p, const double* lin_state, int mode, const lins_point* in, int n,
                           lins_point* out) {
  Estimator e;
  e.prm = Params::from_c(*p);
  e.linState_ = GlobalState::from_array(lin_state);
  for (int i = 0; i < n; ++i) {
    lins_point o = in[i];
    if (mode == 0) e.transformToStart(&in[i], &o); else e.transformToEnd(&in[i], &o);
    out[i] = o;
  }
}
// measurement rows (residual + 6 structural Jacobian entries) for given keypoints/coeffs at lin_state
void lins_oracle_measurement_rows(const lins_params* p, const double* lin_state, const lins_point* kp,
                                  const float* coeff4, int n, double* h6, double* r) {
  Estimator e;
  e.prm = Params::from_c(*p);
  e.linState_ = GlobalState::from_array(lin_state);
  Estimator::HCommon hc = e.hCommon();
  for (int i = 0; i < n; ++i) {
    lins_point c;
    std::memset(&c, 0, sizeof(c));
    c.x = coeff4[4 * i]; c.y = coeff4[4 * i + 1]; c.z = coeff4[4 * i + 2]; c.intensity = coeff4[4 * i + 3];
    e.measurementRow(hc, kp[i], c, h6 + 6 * i, r + i);
  }
}

}  // extern "C"

// ---- row F2: mapping-node scan-to-map refinement (lins_map_oracle.hpp) -----------------------------------------------
#include "lins_map_oracle.hpp"

namespace {
std::vector<lins_map_oracle::P3> to_p3(const lins_point* p, int n) {
  std::vector<lins_map_oracle::P3> v(n);
  for (int i = 0; i < n; ++i) v[i] = lins_map_oracle::P3{p[i].x, p[i].y, p[i].z, p[i].intensity};
  return v;
}
}  // namespace

extern "C" {

void* lins_map_oracle_create() { return new lins_map_oracle::Mapper(); }
void lins_map_oracle_destroy(void* h) { delete static_cast<lins_map_oracle::Mapper*>(h); }
int lins_map_oracle_set_map(void* h, const lins_point* corner, int nc, const lins_point* surf, int ns) {
  static_cast<lins_map_oracle::Mapper*>(h)->setMap(corner, nc, surf, ns);
  return 0;
}
int lins_map_oracle_associate(void* h, const lins_point* corner, int nc, const lins_point* surf, int ns, const float* T,
                              int32_t* cknn, int32_t* sknn, float* ccoeff, float* scoeff, uint8_t* cmask, uint8_t* smask) {
  lins_map_oracle::Transform t;
  std::memcpy(t.v, T, sizeof(t.v));
  static_cast<lins_map_oracle::Mapper*>(h)->associate(to_p3(corner, nc), to_p3(surf, ns), t, cknn, sknn, ccoeff, scoeff, cmask, smask,
                                                      nullptr, nullptr);
  return 0;
}
int lins_map_oracle_scan2map(void* h, const lins_point* corner, int nc, const lins_point* surf, int ns, float* T_io,
                             lins_map_report* rep) {
  lins_map_oracle::Transform t;
  std::memcpy(t.v, T_io, sizeof(t.v));
  lins_map_oracle::Report r;
  static_cast<lins_map_oracle::Mapper*>(h)->scan2map(to_p3(corner, nc), to_p3(surf, ns), t, r);
  std::memcpy(T_io, t.v, sizeof(t.v));
  if (rep) {
    std::memset(rep, 0, sizeof(*rep));
    rep->iters = r.iters; rep->converged = r.converged; rep->degenerate = r.degenerate; rep->skipped = r.skipped;
    for (int i = 0; i < (int)r.n_sel.size() && i < LINS_MAP_MAX_ITER; ++i) {
      rep->n_sel[i] = r.n_sel[i]; rep->delta_r[i] = r.delta_r[i]; rep->delta_t[i] = r.delta_t[i];
    }
  }
  return 0;
}
// the 6x6 step alone (what the product's host side does after the device reduction)
int lins_map_oracle_lm_solve(void* h, const float* AtA, const float* AtB, int iter, float* T_io, float* X_out, int* degenerate) {
  lins_map_oracle::Mapper* m = static_cast<lins_map_oracle::Mapper*>(h);
  lins_map_oracle::Transform t;
  std::memcpy(t.v, T_io, sizeof(t.v));
  lins_map_oracle::Report r;
  float X[6];
  const bool conv = m->lm_solve(AtA, AtB, iter, t, r, X);
  std::memcpy(T_io, t.v, sizeof(t.v));
  if (X_out) std::memcpy(X_out, X, sizeof(X));
  if (degenerate) *degenerate = r.degenerate;
  return conv ? 1 : 0;
}
// numerical kernels, for pinning against the real OpenCV (tests/test_map_oracle_cpu.py)
void lins_map_oracle_eigen(const float* A, int n, float* W, float* V) {
  std::vector<float> a(A, A + (size_t)n * n);
  lins_map_oracle::jacobi_eigen(a.data(), n, W, V);
}
int lins_map_oracle_qr_solve(const float* A, int m, int n, const float* b, float* x) {
  std::vector<float> a(A, A + (size_t)m * n), bb(b, b + m);
  const bool ok = lins_map_oracle::qr_solve(a.data(), m, n, bb.data());
  for (int i = 0; i < n; ++i) x[i] = ok ? bb[i] : 0.f;
  return ok ? 1 : 0;
}
int lins_map_oracle_lu_invert(const float* A, int n, float* Ainv) {
  std::vector<float> a(A, A + (size_t)n * n);
  return lins_map_oracle::lu_invert(a.data(), n, Ainv) ? 1 : 0;
}
void lins_map_oracle_gemm(const float* A, const float* B, float* C, int m, int k, int n) { lins_map_oracle::gemm_f32(A, B, C, m, k, n); }

}  // extern "C"
