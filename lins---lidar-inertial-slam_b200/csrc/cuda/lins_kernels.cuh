// Fused iterated-ESKF update kernel for sm_100a (B200).
//
// One CTA owns one scan (a "unit": query features of the new scan, target features of the last scan, prior)
// and runs the WHOLE performIESKF loop (reference lins/include/StateEstimator.hpp:465-600) on device:
//   A2  transformToStart de-skew of every query               (:1066-1080)   f64 math, f32 store
//   A3/4 exact 1-NN + the +-2.5-ring index walks               (:844-915, :970-1029) f32, bit-exact indices
//   A5/6 point-to-plane / point-to-line residual + coeff       (:917-951, :1031-1060)
//   A7-9 Jacobian row (6 structural non-zeros) folded straight into the 18x18 information form
//        (21 + 6 + 1 scalars per scan; SURVEY.md §8 A9 form B) — H (Mx18) and S (MxM) are never formed
//   A10 boxMinus / gain solve / NaN + divergence tests / boxPlus / convergence (:548-580)
//   A11 Joseph covariance update at exit                        (:595-598)
// CTAs pull scans from a global counter (persistent grid sized to the SM count), so iteration-count imbalance
// between scans does not idle SMs.  Queries + prior are staged into shared memory with 1-D TMA
// (cp.async.bulk + mbarrier); the target clouds stay in global memory as packed float4 (x,y,z,intensity)
// and are read through L1/L2 with broadcast or coalesced 16-B loads.  All reductions use a fixed shuffle tree
// + fixed-order cross-warp sum, so results are run-to-run deterministic (no floating-point atomics).
#pragma once
#include "../../../include/lins_gpu.h"
#include "lins_device_math.cuh"

namespace lins_dev {

#ifndef LINS_THREADS
#define LINS_THREADS 512
#endif
#ifndef LINS_MIN_CTAS
#define LINS_MIN_CTAS 1
#endif
constexpr int kThreads = LINS_THREADS;   // threads per CTA (one unit per CTA); LINS_MIN_CTAS = resident CTAs per SM the register budget is set for
constexpr int kMinCtas = LINS_MIN_CTAS;
constexpr int kWarps = kThreads / 32;
constexpr int kRingTab = 260;   // ring-start table: first target index with ring >= r, r = 0..259
constexpr int kMaxRing = 256;   // rings outside [0, kMaxRing) or unsorted clouds take the sequential walk
constexpr int kNAcc = 28;       // 21 (sym 6x6) + 6 (g*r) + 1 (r*r)
constexpr int kNNChunk = 1024;  // targets per brute-force work item
constexpr int kAzTabS = 4096;   // (ring, azimuth bin) table entries, surf targets (see lins_assoc_az.cuh)
constexpr int kAzTabC = 1024;   // corner targets
constexpr float kPiF = 3.14159265358979f;
constexpr unsigned long long kKeyMax = 0xFFFFFFFFFFFFFFFFull;

enum KernelMode { MODE_IESKF = 0, MODE_ASSOC = 1, MODE_ICP_REDUCE = 2, MODE_JACOBIAN = 3 };

struct BatchView {
  int n_scans;
  const float4* qs; const int* qs_off;   // surf queries   (surfPointsFlat_)
  const float4* qc; const int* qc_off;   // corner queries (cornerPointsSharp_)
  const float4* ts; const int* ts_off;   // surf targets   (last surfPointsLessFlat_)
  const float4* tc; const int* tc_off;   // corner targets (last cornerPointsLessSharp_)
  // clouds the 1-NN index was built on; null = same as ts/tc.  They differ only after a map refresh that
  // failed the >=5 && >=20 guard (StateEstimator.hpp:1156-1157): scan_last_ advanced, the kd-trees did not.
  const float4* nn_s; const int* nn_s_off;
  const float4* nn_c; const int* nn_c_off;
  const double* state_in;                // n x 20 (19 used)
  const double* cov_in;                  // n x 324, column-major
  double* state_out;                     // n x 20
  double* cov_out;                       // n x 324, column-major
  lins_scan_result* results;             // n
  lins_report* reports;                  // n or null
  int* ind_s;                            // 3 per surf query   (pointSearchSurfInd1/2/3)
  int* ind_c;                            // 2 per corner query (pointSearchCornerInd1/2)
  float* sel_s; float* sel_c;            // optional dense traces (3 per query)
  float* coeff_s; float* coeff_c;        // optional (4 per query)
  unsigned char* mask_s; unsigned char* mask_c;  // optional
  double* accum;                         // n x 32 : 28 accumulators + m_surf + m_corner (modes 2,3)
  float4* az_s; float4* az_c;            // global scratch for the (ring, azimuth)-sorted copies when they do not fit
                                         // shared memory; same per-scan offsets as ts / tc
  int cap_s, cap_c;                      // points of the sorted copies held in shared memory (0 = use the scratch)
  int* work_counter;
  long long* timers;                     // optional: per-phase SM cycles summed over CTAs (diagnostics), 32 slots
  int qtile;                             // queries staged per pass
};

struct KParams {
  int num_iter, icp_freq, force_all_iters, mode, iter0;
  double nearest_sq, lidar_std, lidar_scale, scan_period;
};

__device__ __forceinline__ unsigned long long pack_key(float d, unsigned int lo) {
  return ((unsigned long long)__float_as_uint(d) << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    unsigned long long o = shfl_xor_u64(v, m);
    v = o < v ? o : v;
  }
  return v;
}

// shared-memory carve-up ---------------------------------------------------------------------------------
struct Smem {
  // iteration-invariant
  double prior[20];
  double P[324];        // prior covariance, row-major
  // iterate
  double lin[20];
  double phi[3];        // Quat2axis(lin q)
  double R[9];          // toRotationMatrix(lin q)
  double Rinv[9];       // Rinvleft(-phi)
  double dvec[18];      // filterState (-) linState
  double acc[kNAcc + 4];
  double wacc[kWarps][kNAcc];
  int wcnt[kWarps][2];
  int cnt[2];
  double A6[36];
  double y6[6];
  double M6[36];        // A6 P_cc + sig2 I  (6x6 gain system, see form_M6)
  double X6[6 * 12];    // right-hand sides / solutions of the 6x6 system
  double U[18 * 6];     // P[:,c] M^-1 A6
  double V[18 * 6];     // P[:,c] M^-1
  double X[18 * 18];    // (I - K H) P
  double upd[18];
  double residualNorm;
  int flags[4];         // 0 converged 1 diverged 2 has_nan 3 stop
  int scan;
  int sortedS, sortedC;
  int rsS[kRingTab];
  int rsC[kRingTab];
  int azTabS[kAzTabS + 1];  // (ring, azimuth bin) -> first slot of the sorted surf copy
  int azTabC[kAzTabC + 1];
  int scan_tmp[kThreads];
  int dbg[2];               // diagnostics (only touched when phase timers are enabled)
  int wl_n[2], wl_head[2];  // work lists of the closest-point / walk phases (entries, next entry to hand out)
  int az_ok;                // 1: both clouds ring-sorted + indexed (fast path); 0: legacy brute-force / plain walks
  int nbS, nbC, nringsS, nringsC;
  unsigned long long mbar;
  unsigned int phase;
  long long tlast;
};

// phase timers (thread 0 of each CTA; compiled in, enabled when bv.timers != nullptr)
#define LINS_TICK(k)                                                                   \
  do {                                                                                 \
    if (bv.timers && threadIdx.x == 0) {                                               \
      const long long _t = clock64();                                                  \
      atomicAdd((unsigned long long*)&bv.timers[k], (unsigned long long)(_t - sm.tlast)); \
      sm.tlast = _t;                                                                   \
    }                                                                                  \
  } while (0)

// same, and the interval is also added to slot kf when `first` holds (first pass of a scan)
#define LINS_TICK_F(k, kf, first)                                                       \
  do {                                                                                 \
    if (bv.timers && threadIdx.x == 0) {                                               \
      const long long _t = clock64();                                                  \
      atomicAdd((unsigned long long*)&bv.timers[k], (unsigned long long)(_t - sm.tlast)); \
      if (first) atomicAdd((unsigned long long*)&bv.timers[kf], (unsigned long long)(_t - sm.tlast)); \
      sm.tlast = _t;                                                                   \
    }                                                                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// A2 transformToStart (StateEstimator.hpp:1066-1080) -------------------------------------------------------
__device__ __forceinline__ float4 transform_to_start(const float4 p, const Smem& sm, double scan_period) {
  float fi = p.w - (float)((int)p.w);
  double s = (1.f / scan_period) * fi;
  d3 v = mk3(s * sm.phi[0], s * sm.phi[1], s * sm.phi[2]);
  q4 r = axis2Quat(v);
  d3 P2 = mk3((double)p.x, (double)p.y, (double)p.z);
  d3 rp = qrot(r, P2);
  d3 t = mk3(s * sm.lin[0], s * sm.lin[1], s * sm.lin[2]);
  float4 o;
  o.x = (float)(rp.x + t.x); o.y = (float)(rp.y + t.y); o.z = (float)(rp.z + t.z); o.w = p.w;
  return o;
}

// ring-start table: rs[r] = first index with ring >= r (valid only if the cloud is ring-sorted, rings in
// [0, kMaxRing)).  sorted flag cleared otherwise.
__device__ void build_ring_table(const float4* __restrict__ tgt, int T, int* rs, int* sorted_flag) {
  for (int r = threadIdx.x; r < kRingTab; r += kThreads) rs[r] = T;
  if (threadIdx.x == 0) *sorted_flag = 1;
  __syncthreads();
  for (int j = threadIdx.x; j < T; j += kThreads) {
    int rj = (int)__ldg(&tgt[j]).w;
    int rp = j > 0 ? (int)__ldg(&tgt[j - 1]).w : -1;
    if (rj < 0 || rj >= kMaxRing || (j > 0 && rj < rp)) { *sorted_flag = 0; continue; }
    if (j == 0) { for (int r = 0; r <= rj; ++r) rs[r] = 0; }
    else if (rj != rp) { for (int r = (rp < 0 ? 0 : rp + 1); r <= rj; ++r) rs[r] = j; }
  }
  __syncthreads();
}

// A3/A4 exact brute-force 1-NN: every (32-query group, kNNChunk-target chunk) pair is one warp work item; all
// lanes of a warp read the same target (broadcast 16-B load), each lane keeps its own query's best.
__device__ void nn_brute(const float4* sel, unsigned long long* key, int nq, const float4* __restrict__ tgt, int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqg = (nq + 31) >> 5, nch = (T + kNNChunk - 1) / kNNChunk;
  for (int item = warp; item < nqg * nch; item += kWarps) {
    const int qg = item % nqg, ch = item / nqg;
    const int qi = qg * 32 + lane;
    const bool valid = qi < nq;
    const float4 s = valid ? sel[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int j0 = ch * kNNChunk, j1 = min(T, j0 + kNNChunk);
    float best = __int_as_float(0x7f800000);
    int bi = -1;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const float4 t = __ldg(&tgt[j]);
      const float d = sqdist_f32(s.x, s.y, s.z, t.x, t.y, t.z);
      if (d < best) { best = d; bi = j; }
    }
    if (valid && bi >= 0) atomicMin(&key[qi], pack_key(best, (unsigned)bi));
  }
}


// walk keys: high 32 = f32 distance bits, low 32 = visiting order (forward walk ascending j first, then the
// backward walk descending j) so that u64 min == "first strictly smaller wins" of the reference loops.
__device__ __forceinline__ unsigned int order_fwd(int j) { return (unsigned)j; }
__device__ __forceinline__ unsigned int order_bwd(int j) { return 0x80000000u | (unsigned)(0x7fffffff - j); }
__device__ __forceinline__ int order_decode(unsigned int o) { return (o & 0x80000000u) ? (0x7fffffff - (int)(o & 0x7fffffffu)) : (int)o; }

// Ring walks of one query by one warp, ring-sorted fast path (StateEstimator.hpp:859-910 / :983-1024).
template <bool SURF>
__device__ __forceinline__ void walk_warp(const float4 s, int c, const float4* __restrict__ tgt, int T, const int* rs,
                                          int fwdBound, float nearf, int& i2, int& i3) {
  const int lane = threadIdx.x & 31;
  const int cr = (int)__ldg(&tgt[c]).w;
  const int rlo = cr - 2, rhi = cr + 3;
  const int lo = rlo <= 0 ? 0 : (rlo >= kRingTab ? T : rs[rlo]);
  int hi = rhi >= kRingTab ? T : (rhi <= 0 ? 0 : rs[rhi]);
  hi = min(hi, fwdBound);
  const unsigned long long init = pack_key(nearf, 0u);
  unsigned long long k2 = init, k3 = init;
  for (int j = c + 1 + lane; j < hi; j += 32) {  // forward
    const float4 t = __ldg(&tgt[j]);
    const float d = sqdist_f32(t.x, t.y, t.z, s.x, s.y, s.z);
    const int rj = (int)t.w;
    const unsigned long long k = pack_key(d, order_fwd(j));
    if (SURF) {
      if (rj <= cr) { if (d < nearf && k < k2) k2 = k; }
      else { if (d < nearf && k < k3) k3 = k; }
    } else {
      if (rj > cr) { if (d < nearf && k < k2) k2 = k; }
    }
  }
  for (int j = lo + lane; j < c; j += 32) {  // backward
    const float4 t = __ldg(&tgt[j]);
    const float d = sqdist_f32(t.x, t.y, t.z, s.x, s.y, s.z);
    const int rj = (int)t.w;
    const unsigned long long k = pack_key(d, order_bwd(j));
    if (SURF) {
      if (rj >= cr) { if (d < nearf && k < k2) k2 = k; }
      else { if (d < nearf && k < k3) k3 = k; }
    } else {
      if (rj < cr) { if (d < nearf && k < k2) k2 = k; }
    }
  }
  k2 = warp_min_u64(k2);
  i2 = k2 == init ? -1 : order_decode((unsigned)(k2 & 0xffffffffu));
  if (SURF) {
    k3 = warp_min_u64(k3);
    i3 = k3 == init ? -1 : order_decode((unsigned)(k3 & 0xffffffffu));
  } else {
    i3 = -1;
  }
}

// Sequential walk, literal restatement (any ring order / values).  One thread.
template <bool SURF>
__device__ void walk_seq(const float4 s, int c, const float4* __restrict__ tgt, int T, int fwdBound, float nearf,
                         int& i2, int& i3) {
  const int cr = (int)__ldg(&tgt[c]).w;
  float m2 = nearf, m3 = nearf;
  i2 = -1; i3 = -1;
  for (int j = c + 1; j < fwdBound; ++j) {
    const float4 t = __ldg(&tgt[j]);
    const int rj = (int)t.w;
    if ((double)rj > (double)cr + 2.5) break;
    const float d = sqdist_f32(t.x, t.y, t.z, s.x, s.y, s.z);
    if (SURF) {
      if (rj <= cr) { if (d < m2) { m2 = d; i2 = j; } }
      else { if (d < m3) { m3 = d; i3 = j; } }
    } else {
      if (rj > cr) { if (d < m2) { m2 = d; i2 = j; } }
    }
  }
  for (int j = c - 1; j >= 0; --j) {
    const float4 t = __ldg(&tgt[j]);
    const int rj = (int)t.w;
    if ((double)rj < (double)cr - 2.5) break;
    const float d = sqdist_f32(t.x, t.y, t.z, s.x, s.y, s.z);
    if (SURF) {
      if (rj >= cr) { if (d < m2) { m2 = d; i2 = j; } }
      else { if (d < m3) { m3 = d; i3 = j; } }
    } else {
      if (rj < cr) { if (d < m2) { m2 = d; i2 = j; } }
    }
  }
}

// A5 plane residual (StateEstimator.hpp:917-951).  Returns accept mask; coeff = (s*jac, s*res).
__device__ __forceinline__ bool plane_residual(const float4 sel, const float4 t1, const float4 t2, const float4 t3,
                                               bool weighted, float4& coeff) {
  d3 P0 = mk3(sel.x, sel.y, sel.z), P1 = mk3(t1.x, t1.y, t1.z), P2 = mk3(t2.x, t2.y, t2.z), P3 = mk3(t3.x, t3.y, t3.z);
  d3 M = cross3(sub3(P1, P2), sub3(P1, P3));
  double r = dot3(sub3(P0, P1), M);
  double m = norm3(M);
  float res = (float)(r / m);
  d3 jac = div3(M, m);
  float s = 1.f;
  if (weighted) {
    float n2 = __fadd_rn(__fadd_rn(__fmul_rn(sel.x, sel.x), __fmul_rn(sel.y, sel.y)), __fmul_rn(sel.z, sel.z));
    float rt = sqrtf(sqrtf(n2));
    s = (float)(1.0 - 1.8 * (double)fabsf(res) / (double)rt);
  }
  if ((double)s > 0.1 && res != 0.f) {
    coeff.x = (float)((double)s * jac.x); coeff.y = (float)((double)s * jac.y); coeff.z = (float)((double)s * jac.z);
    coeff.w = __fmul_rn(s, res);
    return true;
  }
  return false;
}
// A6 line residual (StateEstimator.hpp:1031-1060)
__device__ __forceinline__ bool line_residual(const float4 sel, const float4 t1, const float4 t2, bool weighted,
                                              float4& coeff) {
  d3 P0 = mk3(sel.x, sel.y, sel.z), P1 = mk3(t1.x, t1.y, t1.z), P2 = mk3(t2.x, t2.y, t2.z);
  d3 P = cross3(sub3(P0, P1), sub3(P0, P2));
  float r = (float)norm3(P);
  float d12 = (float)norm3(sub3(P1, P2));
  float res = __fdiv_rn(r, d12);
  d3 a = sub3(P2, P1);
  d3 num = mk3(P.y * a.z + P.z * (-a.y), P.x * (-a.z) + P.z * a.x, P.x * a.y + P.y * (-a.x));
  double den = (double)__fmul_rn(d12, r);
  d3 jac = div3(num, den);
  float s = 1.f;
  if (weighted) s = (float)(1.0 - 1.8 * (double)fabsf(res));
  if ((double)s > 0.1 && res != 0.f) {
    coeff.x = (float)((double)s * jac.x); coeff.y = (float)((double)s * jac.y); coeff.z = (float)((double)s * jac.z);
    coeff.w = __fmul_rn(s, res);
    return true;
  }
  return false;
}

// A7-A9: fold one accepted measurement into the per-thread accumulators.
// g = [c ; P2 x (R^T c)]  (the Jacobian row is h = [c ; Rinv^T g_att], applied once per scan at solve time)
__device__ __forceinline__ void accumulate_row(const float4 kp, const float4 coeff, const double* R, double lidar_scale,
                                               double* acc) {
  const double cx = coeff.x, cy = coeff.y, cz = coeff.z;
  const double r = lidar_scale * (double)coeff.w;
  const double ux = R[0] * cx + R[3] * cy + R[6] * cz;
  const double uy = R[1] * cx + R[4] * cy + R[7] * cz;
  const double uz = R[2] * cx + R[5] * cy + R[8] * cz;
  const double px = kp.x, py = kp.y, pz = kp.z;
  double g[6];
  g[0] = cx; g[1] = cy; g[2] = cz;
  g[3] = py * uz - pz * uy; g[4] = pz * ux - px * uz; g[5] = px * uy - py * ux;
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) acc[k++] += g[a] * g[b];
#pragma unroll
  for (int a = 0; a < 6; ++a) acc[21 + a] += g[a] * r;
  acc[27] += r * r;
}
// ICP fallback variant (StateEstimator.hpp:1228-1258): interpolated rotation per point, rotation block first.
__device__ __forceinline__ void accumulate_row_icp(const float4 kp, const float4 coeff, const Smem& sm, double scan_period,
                                                   double* acc) {
  float fi = kp.w - (float)((int)kp.w);
  double s = (1.f / scan_period) * fi;
  q4 rq = axis2Quat(mk3(s * sm.phi[0], s * sm.phi[1], s * sm.phi[2]));
  m3 R = qtoR(rq);
  const double cx = coeff.x, cy = coeff.y, cz = coeff.z;
  const double b = -0.05 * (double)coeff.w;
  const double ux = R.m[0] * cx + R.m[3] * cy + R.m[6] * cz;
  const double uy = R.m[1] * cx + R.m[4] * cy + R.m[7] * cz;
  const double uz = R.m[2] * cx + R.m[5] * cy + R.m[8] * cz;
  const double px = kp.x, py = kp.y, pz = kp.z;
  double g[6];
  g[0] = py * uz - pz * uy; g[1] = pz * ux - px * uz; g[2] = px * uy - py * ux;
  g[3] = cx; g[4] = cy; g[5] = cz;
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = a; c < 6; ++c) acc[k++] += g[a] * g[c];
#pragma unroll
  for (int a = 0; a < 6; ++a) acc[21 + a] += g[a] * b;
  acc[27] += b * b;
}

// fixed-tree block reduction of the per-thread accumulators into sm.acc / sm.cnt
// finish = false: stop after the per-warp partial sums are in shared memory (one barrier); the caller's warp 0
// adds them with finish_acc_warp0.
__device__ void block_reduce_acc(Smem& sm, double* acc, int cs, int cc, bool finish = true) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kNAcc; ++k) {
    double v = acc[k];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) sm.wacc[warp][k] = v;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) { cs += __shfl_xor_sync(0xffffffffu, cs, m); cc += __shfl_xor_sync(0xffffffffu, cc, m); }
  if (lane == 0) { sm.wcnt[warp][0] = cs; sm.wcnt[warp][1] = cc; }
  __syncthreads();
  if (!finish) return;
  if (threadIdx.x < kNAcc) {  // accumulates: callers zero sm.acc / sm.cnt at the start of a pass
    double v = sm.acc[threadIdx.x];
    for (int w = 0; w < kWarps; ++w) v += sm.wacc[w][threadIdx.x];
    sm.acc[threadIdx.x] = v;
  }
  if (threadIdx.x == 32) {
    int a = sm.cnt[0], b = sm.cnt[1];
    for (int w = 0; w < kWarps; ++w) { a += sm.wcnt[w][0]; b += sm.wcnt[w][1]; }
    sm.cnt[0] = a; sm.cnt[1] = b;
  }
  __syncthreads();
}

__device__ __forceinline__ int col6(int a) { return a < 3 ? a : a + 3; }  // {0,1,2,6,7,8}

// M = A6 P_cc + sig2 I6: the 6x6 system of the gain in push-through form.  With H = H6 E_c^T (only the 6
// structural columns c = {0,1,2,6,7,8} are non-zero) and R = sig2 I:
//   K = P H^T (H P H^T + R)^-1 = P[:,c] (A6 P[c,c] + sig2 I6)^-1 H6^T,   A6 = H6^T H6
// so  K (r + H d) = P[:,c] M^-1 (b6 + A6 d_c)  and nothing larger than 6x6 is ever factorised.
__device__ void form_M6(Smem& sm, double sig2, int idx, int stride) {
  for (int t = idx; t < 36; t += stride) {
    const int a = t / 6, c = t % 6;
    double s = 0;
    for (int k = 0; k < 6; ++k) s += sm.A6[a * 6 + k] * sm.P[col6(k) * 18 + col6(c)];
    if (a == c) s += sig2;
    sm.M6[t] = s;
  }
}

// linState (+) updateVec (KalmanFilter.hpp:71-81), thread 0
__device__ void box_plus(Smem& sm) {
  double* l = sm.lin; const double* u = sm.upd;
  for (int i = 0; i < 3; ++i) {
    l[0 + i] += u[0 + i]; l[3 + i] += u[3 + i]; l[10 + i] += u[9 + i]; l[13 + i] += u[12 + i]; l[16 + i] += u[15 + i];
  }
  q4 q; q.x = l[6]; q.y = l[7]; q.z = l[8]; q.w = l[9];
  q4 dq = axis2Quat(mk3(u[6], u[7], u[8]));
  q4 r = qnormalized(qmul(q, dq));
  l[6] = r.x; l[7] = r.y; l[8] = r.z; l[9] = r.w;
}


// ---------------------------------------------------------------------------------------------------------
// The serial tail of an iteration, executed by warp 0 alone between two block barriers.  Everything below is
// written so that the 32 lanes work on different entries of the same small formula in lockstep.

// N x N LU with partial pivoting, one COLUMN per lane held in registers: lanes [0, N) own the columns of S
// (row-major in shared memory), lanes [N, N + nrhs) the right-hand sides B (N x nrhs row-major, overwritten by
// the solution).  Pivot = first maximum of |column| from the diagonal down (what a sequential scan picks; a NaN
// on the diagonal poisons the step); one division per step; only the initial loads and the final stores touch
// shared memory.  nrhs <= 32 - N.
template <int N>
__device__ bool warp_lu_cols(const double* S, double* B, int nrhs) {
  const int lane = threadIdx.x & 31;
  const bool is_rhs = lane >= N && lane < N + nrhs;
  double c[N];
#pragma unroll
  for (int r = 0; r < N; ++r) c[r] = lane < N ? S[r * N + lane] : (is_rhs ? B[r * nrhs + (lane - N)] : 0.0);
  double dinv = 0.0;  // lane k keeps 1 / U[k][k]
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = fabs(c[k]);
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      const double v = fabs(c[r]);
      if (v > best) { best = v; piv = r; }
    }
    piv = __shfl_sync(0xffffffffu, piv, k);
    best = __shfl_sync(0xffffffffu, best, k);
    if (!(best > 0.0)) { ok = false; break; }  // singular or NaN pivot (uniform: broadcast value)
#pragma unroll
    for (int r = k + 1; r < N; ++r)
      if (piv == r) { const double t = c[k]; c[k] = c[r]; c[r] = t; }
    const double inv = 1.0 / c[k];  // meaningful in lane k
    if (lane == k) dinv = inv;
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      const double f = __shfl_sync(0xffffffffu, c[r] * inv, k);
      if (f != 0.0 && lane > k) c[r] -= f * c[k];
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    double sacc = c[k];
#pragma unroll
    for (int m = k + 1; m < N; ++m) {
      const double u = __shfl_sync(0xffffffffu, c[k], m);  // U[k][m] lives in lane m
      sacc -= u * c[m];
    }
    const double inv = __shfl_sync(0xffffffffu, dinv, k);
    if (is_rhs) c[k] = sacc * inv;
  }
  if (is_rhs) {
#pragma unroll
    for (int r = 0; r < N; ++r) B[r * nrhs + (lane - N)] = c[r];
  }
  return true;
}

// sm.acc / sm.cnt += the per-warp partial sums of the last tile (block_reduce_acc with finish = false)
__device__ __forceinline__ void finish_acc_warp0(Smem& sm) {
  const int lane = threadIdx.x & 31;
  if (lane < kNAcc) {
    double v = sm.acc[lane];
    for (int w = 0; w < kWarps; ++w) v += sm.wacc[w][lane];
    sm.acc[lane] = v;
  } else if (lane < kNAcc + 2) {
    const int k = lane - kNAcc;
    int a = sm.cnt[k];
    for (int w = 0; w < kWarps; ++w) a += sm.wcnt[w][k];
    sm.cnt[k] = a;
  }
  __syncwarp();
}

// A6 = T Ag T^T and y6 = T b with T = blockdiag(I3, Rinv^T), entry by entry with the structural zeros of T
// skipped (adding those zero products would not change a finite result).  Two lockstep rounds: the 27 entries
// that need arithmetic + 5 copies, then 4 copies + y6.
__device__ __forceinline__ void build_A6_warp0(Smem& sm) {
  const int lane = threadIdx.x & 31;
  auto Ag = [&](int i, int j) -> double {
    if (i > j) { const int tmp = i; i = j; j = tmp; }
    return sm.acc[i * 6 - (i * (i - 1)) / 2 + (j - i)];
  };
  auto entry = [&](int a, int c) -> double {
    if (a < 3 && c < 3) return Ag(a, c);
    if (a < 3) {  // Ag12 Rinv
      double u = 0;
      for (int j = 0; j < 3; ++j) u += Ag(a, 3 + j) * sm.Rinv[j * 3 + (c - 3)];
      return u;
    }
    if (c < 3) {  // Rinv^T Ag21
      double v = 0;
      for (int i = 0; i < 3; ++i) v += sm.Rinv[i * 3 + (a - 3)] * Ag(3 + i, c);
      return v;
    }
    double v = 0;  // Rinv^T Ag22 Rinv
    for (int i = 0; i < 3; ++i) {
      double u = 0;
      for (int j = 0; j < 3; ++j) u += Ag(3 + i, 3 + j) * sm.Rinv[j * 3 + (c - 3)];
      v += sm.Rinv[i * 3 + (a - 3)] * u;
    }
    return v;
  };
  // entry order: [3..5]x[3..5] (9), [0..2]x[3..5] (9), [3..5]x[0..2] (9), [0..2]x[0..2] (9)
  auto coords = [&](int e, int& a, int& c) {
    const int blk = e / 9, w = e % 9;
    a = w / 3 + ((blk == 0 || blk == 2) ? 3 : 0);
    c = w % 3 + ((blk == 0 || blk == 1) ? 3 : 0);
  };
  {
    int a, c;
    coords(lane, a, c);
    sm.A6[a * 6 + c] = entry(a, c);
  }
  if (lane < 4) {
    int a, c;
    coords(32 + lane, a, c);
    sm.A6[a * 6 + c] = entry(a, c);
  } else if (lane < 10) {
    const int a = lane - 4;
    double v = 0;
    if (a < 3) v = sm.acc[21 + a];
    else for (int i = 0; i < 3; ++i) v += sm.Rinv[i * 3 + (a - 3)] * sm.acc[24 + i];
    sm.y6[a] = v;  // b_h (A6 d_c is added by the caller)
  }
  __syncwarp();
}

// Start-of-iteration constants and filterState (-) linState in one lockstep pass: lane 0 derives phi, R, Rinv
// from the iterate, lane 1 the attitude part of boxMinus — both are Quat2axis of a quaternion, so the expensive
// part runs once for the two lanes.  (compute_iter_consts + box_minus, KalmanFilter.hpp:84-94.)
__device__ __forceinline__ void iter_consts_warp0(Smem& sm) {
  const int lane = threadIdx.x & 31;
  const double* f = sm.prior; const double* l = sm.lin;
  if (lane >= 2 && lane < 17) {  // the 15 vector-space components of boxMinus
    const int e = lane - 2, blk = e / 3, i = e % 3;
    const int so = blk == 0 ? 0 : blk == 1 ? 3 : 10 + 3 * (blk - 2);   // state offsets 0,3,10,13,16
    const int eo = blk == 0 ? 0 : blk == 1 ? 3 : 9 + 3 * (blk - 2);    // error-state offsets 0,3,9,12,15
    sm.dvec[eo + i] = f[so + i] - l[so + i];
  }
  if (lane < 2) {
    q4 ql; ql.x = l[6]; ql.y = l[7]; ql.z = l[8]; ql.w = l[9];
    q4 qf; qf.x = f[6]; qf.y = f[7]; qf.z = f[8]; qf.w = f[9];
    const q4 qd = qmul(qinverse(ql), qf);
    const q4 qin = lane == 0 ? ql : qd;
    const d3 ax = Quat2axis(qin);
    if (lane == 1) { sm.dvec[6] = ax.x; sm.dvec[7] = ax.y; sm.dvec[8] = ax.z; }
    else {
      sm.phi[0] = ax.x; sm.phi[1] = ax.y; sm.phi[2] = ax.z;
      const m3 R = qtoR(ql);
      const m3 Ri = Rinvleft(mk3(-ax.x, -ax.y, -ax.z));
      for (int i = 0; i < 9; ++i) { sm.R[i] = R.m[i]; sm.Rinv[i] = Ri.m[i]; }
    }
  }
  __syncwarp();
}

}  // namespace lins_dev
