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

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRingTab = 260;   // ring-start table: first target index with ring >= r, r = 0..259
constexpr int kMaxRing = 256;   // rings outside [0, kMaxRing) or unsorted clouds take the sequential walk
constexpr int kNAcc = 28;       // 21 (sym 6x6) + 6 (g*r) + 1 (r*r)
constexpr int kNNChunk = 1024;  // targets per brute-force work item
// exact 1-NN index: 2-D (x,y) spatial hash, cell edge kCell metres, rebuilt per scan on device
constexpr int kHashS = 4096;    // buckets, surf targets
constexpr int kHashC = 1024;    // buckets, corner targets
constexpr float kCell = 0.5f;
constexpr float kInvCell = 2.0f;
constexpr int kGridRounds = 3;  // Chebyshev rings searched before the brute-force fallback
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
  float4* grid_s; float4* grid_c;        // scratch: bucket-sorted copies (x,y,z,bits(orig index)) of the 1-NN clouds,
                                         // same per-scan offsets as the clouds they index
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
  int bstartS[kHashS + 1];  // bucket -> first slot in grid_s (exclusive scan of the bucket histogram)
  int bstartC[kHashC + 1];
  int scan_tmp[kThreads];
  int nlist;                // queries whose 1-NN is not yet proven exact
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

// ---------------------------------------------------------------------------------------------------------
// per-iteration constants (thread 0)
__device__ void compute_iter_consts(Smem& sm) {
  q4 q; q.x = sm.lin[6]; q.y = sm.lin[7]; q.z = sm.lin[8]; q.w = sm.lin[9];
  d3 phi = Quat2axis(q);
  sm.phi[0] = phi.x; sm.phi[1] = phi.y; sm.phi[2] = phi.z;
  m3 R = qtoR(q);
  m3 Ri = Rinvleft(mk3(-phi.x, -phi.y, -phi.z));
  for (int i = 0; i < 9; ++i) { sm.R[i] = R.m[i]; sm.Rinv[i] = Ri.m[i]; }
}

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


// ---- exact 1-NN through a 2-D spatial hash ------------------------------------------------------------------
// cell(p) = (floor(x / kCell), floor(y / kCell)); z is ignored, so a cell is a vertical column.  Any target within
// (real) distance < r * kCell of the query lies in a cell whose coordinates differ by at most r from the
// query's, so after all cells of the Chebyshev rings 0..r have been examined the running minimum d is the
// exact 1-NN as soon as d <= (r * kCell)^2 * (1 - 1e-4)  (the margin dwarfs every f32 rounding involved).
// Hash collisions only add candidates; every candidate is evaluated with the exact f32 distance expression
// and packed as (distance bits, original index), so the u64 minimum is "smallest distance, lowest index"
// regardless of visiting order.  Queries still unproven after kGridRounds rings take the brute-force scan.
__device__ __forceinline__ int cell_of(float v) {
  float c = floorf(v * kInvCell);
  c = fminf(fmaxf(c, -1.0e6f), 1.0e6f);  // monotone clamp; NaN -> -1e6 (such points are never indexed)
  return (int)c;
}
__device__ __forceinline__ unsigned int cell_hash(int ix, int iy) {
  return ((unsigned)ix * 73856093u) ^ ((unsigned)iy * 19349663u);
}

// Build: histogram -> exclusive scan -> scatter.  src = the cloud the reference's kd-tree was built on.
template <int H>
__device__ void grid_build(const float4* __restrict__ src, int T, float4* __restrict__ sorted, int* bstart, int* scan_tmp) {
  for (int b = threadIdx.x; b <= H; b += kThreads) bstart[b] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < T; j += kThreads) {
    const float4 t = __ldg(&src[j]);
    if (!(isfinite(t.x) && isfinite(t.y) && isfinite(t.z))) continue;  // PCL drops non-finite points
    atomicAdd(&bstart[cell_hash(cell_of(t.x), cell_of(t.y)) & (H - 1)], 1);
  }
  __syncthreads();
  // exclusive scan of H counters: each thread owns H / kThreads consecutive buckets
  constexpr int PER = H / kThreads;
  int loc[PER];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) { loc[k] = bstart[threadIdx.x * PER + k]; sum += loc[k]; }
  scan_tmp[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {  // 256 partial sums, 8 per lane
    int part[kThreads / 32];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) { part[k] = scan_tmp[threadIdx.x * (kThreads / 32) + k]; s += part[k]; }
    int incl = s;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, m); if ((int)threadIdx.x >= m) incl += o; }
    int run = incl - s;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) { int v = part[k]; scan_tmp[threadIdx.x * (kThreads / 32) + k] = run; run += v; }
    if (threadIdx.x == 31) bstart[H] = incl;
  }
  __syncthreads();
  {
    int run = scan_tmp[threadIdx.x];
#pragma unroll
    for (int k = 0; k < PER; ++k) { bstart[threadIdx.x * PER + k] = run; run += loc[k]; }
  }
  __syncthreads();
  // scatter with per-bucket cursors (the counters double as cursors, restored afterwards)
  for (int j = threadIdx.x; j < T; j += kThreads) {
    const float4 t = __ldg(&src[j]);
    if (!(isfinite(t.x) && isfinite(t.y) && isfinite(t.z))) continue;
    const int pos = atomicAdd(&bstart[cell_hash(cell_of(t.x), cell_of(t.y)) & (H - 1)], 1);
    sorted[pos] = make_float4(t.x, t.y, t.z, __int_as_float(j));
  }
  __syncthreads();
  // cursors now hold the END of each bucket == start of the next: shift back
  {
    int prev[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int b = threadIdx.x * PER + k;
      prev[k] = b == 0 ? 0 : bstart[b - 1];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) bstart[threadIdx.x * PER + k] = prev[k];
  }
  __syncthreads();
}

// (dx,dy) of the k-th cell of Chebyshev ring r (r >= 1: 8r cells; r == 0: the centre)
__device__ __forceinline__ void ring_cell(int r, int k, int& dx, int& dy) {
  if (r == 0) { dx = 0; dy = 0; return; }
  const int side = k / (2 * r), pos = k - side * 2 * r;
  if (side == 0) { dx = -r + pos; dy = -r; }
  else if (side == 1) { dx = r; dy = -r + pos; }
  else if (side == 2) { dx = r - pos; dy = r; }
  else { dx = -r; dy = r - pos; }
}

// brute-force scan of ALL targets for the listed queries: one warp per query, coalesced 16-B loads
__device__ void nn_brute_listed(const float4* sel, unsigned long long* key, const int* list, int nlist,
                                const float4* __restrict__ tgt, int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int li = warp; li < nlist; li += kWarps) {
    const int qi = list[li];
    const float4 s = sel[qi];
    unsigned long long best = kKeyMax;
    for (int j = lane; j < T; j += 32) {
      const float4 t = __ldg(&tgt[j]);
      const float d = sqdist_f32(s.x, s.y, s.z, t.x, t.y, t.z);
      if (d < __int_as_float(0x7f800000)) { const unsigned long long k = pack_key(d, (unsigned)j); best = k < best ? k : best; }
    }
    best = warp_min_u64(best);
    if (lane == 0) key[qi] = best;  // the full scan supersedes whatever the grid rounds found
  }
}

template <int H>
__device__ void nn_grid(Smem& sm, const float4* sel, unsigned long long* key, int nq, const float4* sorted,
                        const int* bstart, const float4* __restrict__ tgt, int T, int* list) {
  // round 0+1: centre cell + ring 1 (9 cells) for every query
  for (int item = threadIdx.x; item < nq * 9; item += kThreads) {
    const int qi = item / 9, k = item - qi * 9;
    const float4 s = sel[qi];
    const int cx = cell_of(s.x) + (k % 3) - 1, cy = cell_of(s.y) + (k / 3) - 1;
    const unsigned b = cell_hash(cx, cy) & (H - 1);
    unsigned long long best = kKeyMax;
    for (int p = bstart[b], e = bstart[b + 1]; p < e; ++p) {
      const float4 t = sorted[p];
      const float d = sqdist_f32(s.x, s.y, s.z, t.x, t.y, t.z);
      if (d < __int_as_float(0x7f800000)) { const unsigned long long kk = pack_key(d, (unsigned)__float_as_int(t.w)); best = kk < best ? kk : best; }
    }
    if (best != kKeyMax) atomicMin(&key[qi], best);
  }
  if (threadIdx.x == 0) sm.nlist = 0;
  __syncthreads();
  // classify: proven iff d <= (1 * kCell)^2 * (1 - 1e-4)
  {
    const float thr = kCell * kCell * 0.9999f;
    for (int qi = threadIdx.x; qi < nq; qi += kThreads) {
      const unsigned long long k1 = key[qi];
      const float d = __uint_as_float((unsigned)(k1 >> 32));
      if (!(k1 != kKeyMax && d <= thr)) list[atomicAdd(&sm.nlist, 1)] = qi;
    }
  }
  __syncthreads();
  for (int r = 2; r <= kGridRounds && sm.nlist > 0; ++r) {
    const int nl = sm.nlist, ncell = 8 * r;
    for (int item = threadIdx.x; item < nl * ncell; item += kThreads) {
      const int li = item / ncell, k = item - li * ncell;
      const int qi = list[li];
      const float4 s = sel[qi];
      int dx, dy;
      ring_cell(r, k, dx, dy);
      const unsigned b = cell_hash(cell_of(s.x) + dx, cell_of(s.y) + dy) & (H - 1);
      unsigned long long best = kKeyMax;
      for (int p = bstart[b], e = bstart[b + 1]; p < e; ++p) {
        const float4 t = sorted[p];
        const float d = sqdist_f32(s.x, s.y, s.z, t.x, t.y, t.z);
        if (d < __int_as_float(0x7f800000)) { const unsigned long long kk = pack_key(d, (unsigned)__float_as_int(t.w)); best = kk < best ? kk : best; }
      }
      if (best != kKeyMax) atomicMin(&key[qi], best);
    }
    __syncthreads();
    // compact the list in place (order irrelevant): one thread does it serially (lists are short)
    if (threadIdx.x == 0) {
      const float thr = (r * kCell) * (r * kCell) * 0.9999f;
      int w = 0;
      for (int li = 0; li < nl; ++li) {
        const int qi = list[li];
        const unsigned long long k1 = key[qi];
        const float d = __uint_as_float((unsigned)(k1 >> 32));
        if (!(k1 != kKeyMax && d <= thr)) list[w++] = qi;
      }
      sm.nlist = w;
    }
    __syncthreads();
  }
  if (sm.nlist > 0) nn_brute_listed(sel, key, list, sm.nlist, tgt, T);
  __syncthreads();
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
__device__ void block_reduce_acc(Smem& sm, double* acc, int cs, int cc) {
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
  if (threadIdx.x < kNAcc) {
    double v = 0;
    for (int w = 0; w < kWarps; ++w) v += sm.wacc[w][threadIdx.x];
    sm.acc[threadIdx.x] = v;
  }
  if (threadIdx.x == 32) {
    int a = 0, b = 0;
    for (int w = 0; w < kWarps; ++w) { a += sm.wcnt[w][0]; b += sm.wcnt[w][1]; }
    sm.cnt[0] = a; sm.cnt[1] = b;
  }
  __syncthreads();
}

// N x N (N <= 32) LU with partial pivoting by one warp; S row-major (destroyed), B = N x nrhs row-major
// (-> solution).  Pivot choice = first maximum of |column| from the diagonal down (same as a sequential scan).
template <int N>
__device__ bool warp_lu_solve(double* S, double* B, int nrhs) {
  const int lane = threadIdx.x & 31;
  bool ok = true;
  for (int k = 0; k < N; ++k) {
    double v = -1.0;
    if (lane >= k && lane < N) {
      v = fabs(S[lane * N + k]);
      if (v != v) v = (lane == k) ? __longlong_as_double(0x7ff8000000000000ll) : -1.0;
    }
    // NaN on the diagonal poisons the step (a sequential scan would keep best = NaN)
    unsigned nanmask = __ballot_sync(0xffffffffu, v != v);
    int piv = lane;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, v, m);
      int op = __shfl_xor_sync(0xffffffffu, piv, m);
      if (ov > v || (ov == v && op < piv)) { v = ov; piv = op; }
    }
    if (nanmask != 0u || !(v > 0.0)) { ok = false; break; }
    if (piv != k) {
      if (lane < N) { double t = S[k * N + lane]; S[k * N + lane] = S[piv * N + lane]; S[piv * N + lane] = t; }
      for (int j = lane; j < nrhs; j += 32) { double t = B[k * nrhs + j]; B[k * nrhs + j] = B[piv * nrhs + j]; B[piv * nrhs + j] = t; }
    }
    __syncwarp();
    const double inv = 1.0 / S[k * N + k];
    if (lane > k && lane < N) {
      const double f = S[lane * N + k] * inv;
      if (f != 0.0) {
        for (int j = k + 1; j < N; ++j) S[lane * N + j] -= f * S[k * N + j];
        for (int j = 0; j < nrhs; ++j) B[lane * nrhs + j] -= f * B[k * nrhs + j];
      }
    }
    __syncwarp();
  }
  if (!ok) return false;
  for (int k = N - 1; k >= 0; --k) {
    const double inv = 1.0 / S[k * N + k];
    for (int j = lane; j < nrhs; j += 32) {
      double s = B[k * nrhs + j];
      for (int c = k + 1; c < N; ++c) s -= S[k * N + c] * B[c * nrhs + j];
      B[k * nrhs + j] = s * inv;
    }
    __syncwarp();
  }
  return true;
}

__device__ __forceinline__ int col6(int a) { return a < 3 ? a : a + 3; }  // {0,1,2,6,7,8}

// Build A6 = T Ag T^T (information matrix on the 6 structural columns) from the reduced accumulators.
// T = blockdiag(I3, Rinv^T).  36 threads.
__device__ void build_A6(Smem& sm) {
  const int t = threadIdx.x;
  if (t < 36) {
    const int a = t / 6, c = t % 6;
    // Ag full from packed upper triangle
    auto Ag = [&](int i, int j) -> double {
      if (i > j) { int tmp = i; i = j; j = tmp; }
      int idx = i * 6 - (i * (i - 1)) / 2 + (j - i);
      return sm.acc[idx];
    };
    auto T = [&](int i, int j) -> double {
      if (i < 3 || j < 3) return (i == j) ? 1.0 : 0.0;
      return sm.Rinv[(j - 3) * 3 + (i - 3)];  // (Rinv^T)[i-3][j-3]
    };
    double s = 0;
    for (int i = 0; i < 6; ++i) {
      double ti = T(a, i);
      if (ti == 0.0) continue;
      double u = 0;
      for (int j = 0; j < 6; ++j) u += Ag(i, j) * T(c, j);
      s += ti * u;
    }
    sm.A6[t] = s;
  }
  if (t >= 64 && t < 70) {
    const int a = t - 64;
    double s = 0;
    if (a < 3) s = sm.acc[21 + a];
    else for (int i = 0; i < 3; ++i) s += sm.Rinv[i * 3 + (a - 3)] * sm.acc[24 + i];
    sm.y6[a] = s;  // b_h (b + A d added later)
  }
}

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

// filterState (-) linState  (KalmanFilter.hpp:84-94), thread 0
__device__ void box_minus(Smem& sm) {
  const double* f = sm.prior; const double* l = sm.lin;
  for (int i = 0; i < 3; ++i) {
    sm.dvec[0 + i] = f[0 + i] - l[0 + i];
    sm.dvec[3 + i] = f[3 + i] - l[3 + i];
    sm.dvec[9 + i] = f[10 + i] - l[10 + i];
    sm.dvec[12 + i] = f[13 + i] - l[13 + i];
    sm.dvec[15 + i] = f[16 + i] - l[16 + i];
  }
  q4 ql; ql.x = l[6]; ql.y = l[7]; ql.z = l[8]; ql.w = l[9];
  q4 qf; qf.x = f[6]; qf.y = f[7]; qf.z = f[8]; qf.w = f[9];
  d3 da = Quat2axis(qmul(qinverse(ql), qf));
  sm.dvec[6] = da.x; sm.dvec[7] = da.y; sm.dvec[8] = da.z;
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
// One association + reduction pass over all queries of the scan at the current linearisation point.
// Fills sm.acc (28 sums) and sm.cnt (accepted surf / corner counts).
template <int MODE>
__device__ void association_pass(Smem& sm, const BatchView& bv, const KParams& kp, int scan, int iter, float4* qpt,
                                 float4* selbuf, unsigned long long* key, int* indbuf) {
  const int qs0 = bv.qs_off[scan], ns = bv.qs_off[scan + 1] - qs0;
  const int qc0 = bv.qc_off[scan], nc = bv.qc_off[scan + 1] - qc0;
  const int ts0 = bv.ts_off[scan], Ts = bv.ts_off[scan + 1] - ts0;
  const int tc0 = bv.tc_off[scan], Tc = bv.tc_off[scan + 1] - tc0;
  const float4* __restrict__ tgtS = bv.ts + ts0;
  const float4* __restrict__ tgtC = bv.tc + tc0;
  const float4* __restrict__ nnS = bv.nn_s ? bv.nn_s + bv.nn_s_off[scan] : tgtS;
  const float4* __restrict__ nnC = bv.nn_c ? bv.nn_c + bv.nn_c_off[scan] : tgtC;
  const int TnS = bv.nn_s ? bv.nn_s_off[scan + 1] - bv.nn_s_off[scan] : Ts;
  const int TnC = bv.nn_c ? bv.nn_c_off[scan + 1] - bv.nn_c_off[scan] : Tc;
  const float4* gridS = bv.grid_s + (bv.nn_s ? bv.nn_s_off[scan] : ts0);
  const float4* gridC = bv.grid_c + (bv.nn_c ? bv.nn_c_off[scan] : tc0);
  const bool search = (iter % kp.icp_freq) == 0;
  const bool weighted = iter >= kp.icp_freq;
  const float nearf = (float)kp.nearest_sq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  double acc[kNAcc];
#pragma unroll
  for (int k = 0; k < kNAcc; ++k) acc[k] = 0.0;
  int cntS = 0, cntC = 0;

  const int ntot = ns + nc;
  for (int q0 = 0; q0 < ntot; q0 += bv.qtile) {
    const int nq = min(bv.qtile, ntot - q0);
    // sub-lists of this tile: surf [0, nsT) then corner [nsT, nq)
    const int nsT = max(0, min(ns - q0, nq));
    // ---- stage the tile's queries (1-D TMA) --------------------------------------------------------------
    __syncthreads();  // previous tile fully consumed
    if (threadIdx.x == 0) {
      uint32_t bytes = 0;
      if (nsT > 0) bytes += (uint32_t)nsT * 16u;
      if (nq - nsT > 0) bytes += (uint32_t)(nq - nsT) * 16u;
      fence_proxy_async();
      mbar_expect_tx(&sm.mbar, bytes);
      if (nsT > 0) tma_load_1d(qpt, bv.qs + qs0 + q0, (uint32_t)nsT * 16u, &sm.mbar);
      if (nq - nsT > 0) tma_load_1d(qpt + nsT, bv.qc + qc0 + max(0, q0 - ns), (uint32_t)(nq - nsT) * 16u, &sm.mbar);
    }
    {
      const unsigned int ph = sm.phase;  // one mbarrier phase per staged tile
      mbar_wait(&sm.mbar, ph & 1u);
      __syncthreads();
      if (threadIdx.x == 0) sm.phase = ph + 1u;
    }
    LINS_TICK(2);
    // ---- A2: de-skew -------------------------------------------------------------------------------------
    for (int i = threadIdx.x; i < nq; i += kThreads) {
      selbuf[i] = transform_to_start(qpt[i], sm, kp.scan_period);
      key[i] = kKeyMax;
    }
    __syncthreads();
    LINS_TICK(3);
    if (search) {
      // ---- A3/A4: exact 1-NN ------------------------------------------------------------------------------
      if (nsT > 0 && TnS > 0) nn_grid<kHashS>(sm, selbuf, key, nsT, gridS, sm.bstartS, nnS, TnS, indbuf);
      if (nq - nsT > 0 && TnC > 0) nn_grid<kHashC>(sm, selbuf + nsT, key + nsT, nq - nsT, gridC, sm.bstartC, nnC, TnC, indbuf);
      __syncthreads();
      LINS_TICK(4);
      // ---- ring walks ---------------------------------------------------------------------------------------
      const int fwdS = min(ns, Ts), fwdC = min(nc, Tc);  // :859 / :983 loop-bound quirk (+ OOB clamp)
      for (int i = warp; i < nq; i += kWarps) {
        const bool surf = i < nsT;
        const unsigned long long k1 = key[i];
        const float d1 = __uint_as_float((unsigned)(k1 >> 32));
        const int c = (int)(unsigned)(k1 & 0xffffffffu);
        const bool found = (k1 != kKeyMax) && ((double)d1 < kp.nearest_sq) && c < (surf ? Ts : Tc);
        int i1 = -1, i2 = -1, i3 = -1;
        if (found) {
          i1 = c;
          const float4 s = selbuf[i];
          if (surf) {
            if (sm.sortedS) walk_warp<true>(s, c, tgtS, Ts, sm.rsS, fwdS, nearf, i2, i3);
            else { if (lane == 0) walk_seq<true>(s, c, tgtS, Ts, fwdS, nearf, i2, i3); }
          } else {
            if (sm.sortedC) walk_warp<false>(s, c, tgtC, Tc, sm.rsC, fwdC, nearf, i2, i3);
            else { if (lane == 0) walk_seq<false>(s, c, tgtC, Tc, fwdC, nearf, i2, i3); }
          }
        }
        if (lane == 0) {
          indbuf[3 * i] = i1; indbuf[3 * i + 1] = i2; indbuf[3 * i + 2] = i3;
          const int gq = q0 + i;
          if (surf) { int* o = bv.ind_s + 3 * (size_t)(qs0 + gq); o[0] = i1; o[1] = i2; o[2] = i3; }
          else { int* o = bv.ind_c + 2 * (size_t)(qc0 + gq - ns); o[0] = i1; o[1] = i2; }
        }
      }
    } else {
      // iter % ICP_FREQ != 0: reuse pointSearch*Ind (StateEstimator.hpp:844, :970)
      for (int i = threadIdx.x; i < nq; i += kThreads) {
        const int gq = q0 + i;
        if (i < nsT) { const int* o = bv.ind_s + 3 * (size_t)(qs0 + gq); indbuf[3 * i] = o[0]; indbuf[3 * i + 1] = o[1]; indbuf[3 * i + 2] = o[2]; }
        else { const int* o = bv.ind_c + 2 * (size_t)(qc0 + gq - ns); indbuf[3 * i] = o[0]; indbuf[3 * i + 1] = o[1]; indbuf[3 * i + 2] = -1; }
      }
    }
    __syncthreads();
    LINS_TICK(5);
    // ---- A5/A6 residuals + A7-A9 fold ----------------------------------------------------------------------
    for (int i = threadIdx.x; i < nq; i += kThreads) {
      const bool surf = i < nsT;
      const int i1 = indbuf[3 * i], i2 = indbuf[3 * i + 1], i3 = indbuf[3 * i + 2];
      const float4 s = selbuf[i];
      float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
      bool ok = false;
      if (surf) {
        if (i2 >= 0 && i3 >= 0 && i1 >= 0 && i1 < Ts && i2 < Ts && i3 < Ts)
          ok = plane_residual(s, __ldg(&tgtS[i1]), __ldg(&tgtS[i2]), __ldg(&tgtS[i3]), weighted, coeff);
      } else {
        if (i2 >= 0 && i1 >= 0 && i1 < Tc && i2 < Tc) ok = line_residual(s, __ldg(&tgtC[i1]), __ldg(&tgtC[i2]), weighted, coeff);
      }
      if (ok) {
        if (MODE == MODE_ICP_REDUCE) accumulate_row_icp(qpt[i], coeff, sm, kp.scan_period, acc);
        else accumulate_row(qpt[i], coeff, sm.R, kp.lidar_scale, acc);
        if (surf) ++cntS; else ++cntC;
      }
      if (MODE == MODE_ASSOC) {
        const int gq = q0 + i;
        if (surf) {
          const size_t o = (size_t)(qs0 + gq);
          if (bv.sel_s) { bv.sel_s[3 * o] = s.x; bv.sel_s[3 * o + 1] = s.y; bv.sel_s[3 * o + 2] = s.z; }
          if (bv.coeff_s) { bv.coeff_s[4 * o] = coeff.x; bv.coeff_s[4 * o + 1] = coeff.y; bv.coeff_s[4 * o + 2] = coeff.z; bv.coeff_s[4 * o + 3] = coeff.w; }
          if (bv.mask_s) bv.mask_s[o] = ok ? 1 : 0;
        } else {
          const size_t o = (size_t)(qc0 + gq - ns);
          if (bv.sel_c) { bv.sel_c[3 * o] = s.x; bv.sel_c[3 * o + 1] = s.y; bv.sel_c[3 * o + 2] = s.z; }
          if (bv.coeff_c) { bv.coeff_c[4 * o] = coeff.x; bv.coeff_c[4 * o + 1] = coeff.y; bv.coeff_c[4 * o + 2] = coeff.z; bv.coeff_c[4 * o + 3] = coeff.w; }
          if (bv.mask_c) bv.mask_c[o] = ok ? 1 : 0;
        }
      }
    }
  }
  LINS_TICK(6);
  block_reduce_acc(sm, acc, cntS, cntC);
  LINS_TICK(7);
}

}  // namespace lins_dev
