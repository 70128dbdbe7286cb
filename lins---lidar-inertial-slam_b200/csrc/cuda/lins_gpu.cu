// lins_gpu.cu — C-ABI implementation (include/lins_gpu.h) + the fused IESKF kernel entry points, sm_100a.
// Host side is plain C++ / CUDA runtime: no torch, no Eigen, no PCL in any signature.  There is NO CPU
// fallback: every entry point fails with LINS_E_NODEVICE / LINS_E_CUDA when the device path is unavailable.
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "../host/math_utils.hpp"
#include "../host/small_linalg.hpp"
#include "../host/host_pool.hpp"
#include "../host/lins_map_host.hpp"
#include "lins_assoc.cuh"
#include "lins_icp_step.cuh"
#include "lins_map.cuh"

using namespace lins_dev;

// lins_jacobian.cu
extern "C" int lins_jacobian_parts(int n_units, int sm_count);
extern "C" int lins_launch_jacobian_mma(const lins_dev::BatchView* bv, const lins_dev::KParams* kp, int n_units, int sm_count, int P, double* part_acc,
                                        int* part_cnt, cudaStream_t stream);

namespace {

// ---------------------------------------------------------------------------------------------------------
// dynamic shared memory layout:
//   [CtaMem][Smem x nslots][wacc f64 x nslots*nvw*28][wcnt int x nslots*nvw*2][per-query arrays x nslots*qtile]
// per-query arrays (kQueryBytes per query): qpt, sel, qa, qw, qref, qref2, qext (16 B each), key (8), pos (12), qccr (4),
// wl (4).  When even one slot does not fit shared memory they live in a per-CTA global scratch instead (bv.qscratch).
// ---------------------------------------------------------------------------------------------------------
constexpr size_t kQueryBytes = 16 * 7 + 8 + 12 + 4 + 4;
__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }
__host__ __device__ inline size_t query_array_bytes(int nslots, int qtile) { return align16((size_t)nslots * qtile * kQueryBytes); }
__host__ __device__ inline size_t smem_fixed_bytes(int nslots, int qtile) {
  const size_t nvw = (size_t)nslots * (qtile / 32);
  return align16(sizeof(CtaMem)) + (size_t)nslots * align16(sizeof(Smem)) + align16(nvw * kNAcc * sizeof(double)) + align16(nvw * 2 * sizeof(int));
}

// prologue of a freshly claimed unit: prior -> shared, the search index (≙ kdtree*->setInputCloud,
// StateEstimator.hpp:363-364 / :1158-1159) built on device, queries staged, first linearisation constants.  Block-wide.
template <int MODE>
__device__ void unit_prologue(CtaMem& cta, Smem& sm, const BatchView& bv, const KParams& kp, const PassBuffers& pb, int slot) {
  const int tid = threadIdx.x, warp = tid >> 5;
  const int scan = sm.scan;
  if (tid < 20) { const double v = tid < 19 ? bv.state_in[(size_t)scan * 20 + tid] : 0.0; sm.prior[tid] = v; sm.lin[tid] = v; }
  for (int e = tid; e < 108; e += kThreads) {  // P[:, c] (cov_in is column-major)
    const int a = e / 6, c = e % 6;
    sm.Pc[e] = bv.cov_in[(size_t)scan * 324 + col6(c) * 18 + a];
  }
  if (tid == 32) {
    sm.flags[0] = sm.flags[1] = sm.flags[2] = sm.flags[3] = 0; sm.residualNorm = 1e6;
    sm.cnt[0] = sm.cnt[1] = 0;
    sm.iter = MODE == MODE_IESKF ? 0 : kp.iter0;
    sm.fresh = 0; sm.finished = 0; sm.first_pass = 1; sm.pos_valid = 0; sm.pos_is_slot = 0;
    sm.run = 1;
    if (MODE == MODE_IESKF && kp.num_iter <= 0) { sm.run = 0; sm.finished = 1; cta.any_finished = 1; }
    sm.qs0 = bv.qs_off[scan]; sm.ns = bv.qs_off[scan + 1] - sm.qs0;
    sm.qc0 = bv.qc_off[scan]; sm.nc = bv.qc_off[scan + 1] - sm.qc0;
    sm.ts0 = bv.ts_off[scan]; sm.Ts = bv.ts_off[scan + 1] - sm.ts0;
    sm.tc0 = bv.tc_off[scan]; sm.Tc = bv.tc_off[scan + 1] - sm.tc0;
  }
  __syncthreads();
  check_ring_sorted(bv.ts + sm.ts0, sm.Ts, &sm.sortedS);
  check_ring_sorted(bv.tc + sm.tc0, sm.Tc, &sm.sortedC);
  if (tid == 0) {
    // slot payload = ring (7 bits) | index (24 bits); bucket tables hold 16-bit slots
    const bool ok = sm.sortedS && sm.sortedC && bv.nn_s == nullptr && bv.nn_c == nullptr && sm.Ts < 65536 && sm.Tc < 65536;
    sm.az_ok = ok ? 1 : 0;
    sm.nringsS = ok && sm.Ts > 0 ? (int)bv.ts[sm.ts0 + sm.Ts - 1].w + 1 : 0;  // ring-sorted: the last point has the largest ring
    sm.nringsC = ok && sm.Tc > 0 ? (int)bv.tc[sm.tc0 + sm.Tc - 1].w + 1 : 0;
    sm.nbS = az_bins_for(sm.nringsS, kAzTabS);
    sm.nbC = az_bins_for(sm.nringsC, kAzTabC);
  }
  __syncthreads();
  if (sm.az_ok) {
    az_build<kAzTabS>(bv.ts + sm.ts0, sm.Ts, bv.az_s + sm.ts0, sm.azTabS, cta.u.build_tab, cta.scan_tmp, sm.nbS, sm.elevS);
    az_build<kAzTabC>(bv.tc + sm.tc0, sm.Tc, bv.az_c + sm.tc0, sm.azTabC, cta.u.build_tab, cta.scan_tmp, sm.nbC, sm.elevC);
  }
  // queries: staged once per unit — 1-D TMA (cp.async.bulk + mbarrier) into shared memory, plain copies into the scratch
  float4* qdst = pb.qpt + (size_t)slot * bv.qtile;
  if (bv.qscratch == nullptr) {
    const uint32_t bytes = (uint32_t)(sm.ns + sm.nc) * 16u;
    if (bytes > 0) {
      if (tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(&cta.mbar, bytes);
        if (sm.ns > 0) tma_load_1d(qdst, bv.qs + sm.qs0, (uint32_t)sm.ns * 16u, &cta.mbar);
        if (sm.nc > 0) tma_load_1d(qdst + sm.ns, bv.qc + sm.qc0, (uint32_t)sm.nc * 16u, &cta.mbar);
      }
      const unsigned int ph = cta.phase;  // one mbarrier phase per staged unit
      mbar_wait(&cta.mbar, ph & 1u);
      __syncthreads();
      if (tid == 0) cta.phase = ph + 1u;
    }
  } else {
    for (int i = tid; i < sm.ns + sm.nc; i += kThreads) qdst[i] = i < sm.ns ? __ldg(bv.qs + sm.qs0 + i) : __ldg(bv.qc + sm.qc0 + i - sm.ns);
  }
  if (warp == 0) iter_consts_warp0(sm);
  __syncthreads();
}

// The serial tail of one unit's iteration (StateEstimator.hpp:535-580), run by ONE warp: finish the reduction, 6x6 gain
// system, update, convergence logic, and the constants of the unit's NEXT iteration.  The tails of the resident units
// run side by side on different warps.
__device__ void unit_tail(CtaMem& cta, Smem& sm, const BatchView& bv, const KParams& kp, const PassBuffers& pb, int slot, double sig2) {
  const int lane = threadIdx.x & 31;
  const int iter = sm.iter;
  lins_report* rep = bv.reports ? bv.reports + sm.scan : nullptr;
  const int nq = sm.ns + sm.nc;
  finish_acc_warp0(sm, pb.wacc + (size_t)slot * pb.nvw * kNAcc, pb.wcnt + (size_t)slot * pb.nvw * 2, (nq + 31) >> 5);
  build_A6_warp0(sm);
  if (slot == 0) LINS_TICK(18);
  if (lane < 6) {  // y = b6 + A6 d_c on the 6 structural rows
    double y = sm.y6[lane];
    for (int c = 0; c < 6; ++c) y += sm.A6[lane * 6 + c] * sm.dvec[col6(c)];
    sm.X6[lane] = y;
  }
  form_M6(sm, sig2, lane, 32);
  __syncwarp();
  const bool ok = warp_lu_cols<6>(sm.M6, sm.X6, 1);  // z = M^-1 (b6 + A6 d_c)
  __syncwarp();
  if (slot == 0) LINS_TICK(19);
  double u = 0.0;
  if (lane < 18) {  // K (r + H d) = P[:,c] z
    double kx = 0;
    for (int c = 0; c < 6; ++c) kx += sm.Pc[lane * 6 + c] * sm.X6[c];
    u = ok ? (-kx + sm.dvec[lane]) : __longlong_as_double(0x7ff8000000000000ll);
  }
  const bool hasNaN = __ballot_sync(0xffffffffu, u != u) != 0u;  // :553-558
  if (u != u) u = 0.0;
  if (lane < 18) sm.upd[lane] = u;
  __syncwarp();
  double nrm = 0.0;  // lane 0: ||update||^2 (sequential order), lane 1: ||residual||^2
  if (lane == 0) for (int a = 0; a < 18; ++a) nrm += sm.upd[a] * sm.upd[a];
  if (lane == 1) nrm = sm.acc[27];
  nrm = sqrt(nrm);
  const double un = __shfl_sync(0xffffffffu, nrm, 0), rnorm = __shfl_sync(0xffffffffu, nrm, 1);
  if (slot == 0) LINS_TICK(21);
  if (lane == 0) {
    if (rep) {
      rep->m_surf[iter] = sm.cnt[0]; rep->m_corner[iter] = sm.cnt[1];
      rep->residual_norm[iter] = rnorm; rep->update_norm[iter] = un;
    }
    if (hasNaN) {  // :559-563
      sm.flags[2] = 1; sm.flags[1] = 1; sm.flags[3] = 1;
    } else if (rnorm > sm.residualNorm * 10) {  // :566-570
      sm.flags[1] = 1; sm.flags[3] = 1;
    } else {
      box_plus(sm);  // :573
      if (un <= 1e-2 && !kp.force_all_iters) { sm.flags[0] = 1; sm.flags[3] = 1; }  // :576-578
      sm.residualNorm = rnorm;
    }
    const bool search = (iter % kp.icp_freq) == 0;
    if (search) sm.pos_is_slot = sm.az_ok; else if (!sm.pos_valid) sm.pos_is_slot = 0;
    sm.pos_valid = 1;
    sm.first_pass = 0;
    sm.iter = iter + 1;
    if (sm.flags[3] || iter + 1 >= kp.num_iter) { sm.finished = 1; sm.run = 0; cta.any_finished = 1; }
  }
  __syncwarp();
  if (slot == 0) LINS_TICK(23);
  if (!sm.finished) iter_consts_warp0(sm);
  if (slot == 0) LINS_TICK(25);
}

// exit of one finished unit: covariance + outputs (StateEstimator.hpp:585-599).  Block-wide.
__device__ void unit_exit(CtaMem& cta, Smem& sm, const BatchView& bv, double sig2) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int scan = sm.scan, iters = sm.iter;
  const bool diverged = sm.flags[1] != 0;
  auto& ex = cta.u.ex;
  if (!diverged && iters > 0) {
    // Joseph form with the LAST iteration's K, H, R (:595-596), all through the 6x6 system:
    //   K H = U E_c^T,  U = P[:,c] M^-1 A6 ;   K R K^T = sig2 U V^T,  V = P[:,c] M^-1
    for (int e = tid; e < 324; e += kThreads) {
      const int r = e / 18, c = e % 18;
      ex.P[e] = bv.cov_in[(size_t)scan * 324 + c * 18 + r];
    }
    form_M6(sm, sig2, tid, kThreads);
    for (int e = tid; e < 72; e += kThreads) {  // right-hand sides [A6 | I6]
      const int a = e / 12, c = e % 12;
      ex.X6[e] = c < 6 ? sm.A6[a * 6 + c] : (a == c - 6 ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      const bool ok = warp_lu_cols<6>(sm.M6, ex.X6, 12);
      if (!ok) for (int e = lane; e < 72; e += 32) ex.X6[e] = __longlong_as_double(0x7ff8000000000000ll);
    }
    __syncthreads();
    for (int t = tid; t < 216; t += kThreads) {
      const int which = t / 108, e = t % 108, a = e / 6, c = e % 6;
      double sacc = 0;
      for (int k = 0; k < 6; ++k) sacc += sm.Pc[a * 6 + k] * ex.X6[k * 12 + c + 6 * which];
      (which ? ex.V : ex.U)[e] = sacc;
    }
    __syncthreads();
    for (int e = tid; e < 324; e += kThreads) {  // X = (I - K H) P = P - U P[c,:]
      const int i = e / 18, j = e % 18;
      double sacc = ex.P[e];
      for (int c = 0; c < 6; ++c) sacc -= ex.U[i * 6 + c] * ex.P[col6(c) * 18 + j];
      ex.X[e] = sacc;
    }
    __syncthreads();
    for (int e = tid; e < 324; e += kThreads) {  // P <- X (I - K H)^T + sig2 U V^T
      const int i = e / 18, j = e % 18;
      double sacc = ex.X[e], t = 0;
      for (int c = 0; c < 6; ++c) { sacc -= ex.X[i * 18 + col6(c)] * ex.U[j * 6 + c]; t += ex.U[i * 6 + c] * ex.V[j * 6 + c]; }
      ex.P[e] = sacc + t * sig2;
    }
    __syncthreads();
  }
  // state_out / cov_out
  if (tid < 20) bv.state_out[(size_t)scan * 20 + tid] = diverged ? sm.prior[tid] : sm.lin[tid];
  for (int e = tid; e < 324; e += kThreads) {
    const int r = e / 18, c = e % 18;
    double v;
    if (diverged || iters == 0) v = bv.cov_in[(size_t)scan * 324 + c * 18 + r];
    else v = 0.5 * (ex.P[r * 18 + c] + ex.P[c * 18 + r]);  // enforceSymmetry (:597)
    bv.cov_out[(size_t)scan * 324 + c * 18 + r] = v;
  }
  if (tid == 0) {
    lins_scan_result& o = bv.results[scan];
    o.scan_id = scan;
    o.iters = (uint16_t)iters;
    o.flags = (uint16_t)((sm.flags[0] ? 1 : 0) | (sm.flags[1] ? 2 : 0) | (sm.flags[2] ? 4 : 0));
    const double* st = diverged ? sm.prior : sm.lin;
    o.pose[0] = st[0]; o.pose[1] = st[1]; o.pose[2] = st[2];
    o.pose[3] = st[6]; o.pose[4] = st[7]; o.pose[5] = st[8]; o.pose[6] = st[9];
    lins_report* rep = bv.reports ? bv.reports + scan : nullptr;
    if (rep) { rep->iters = iters; rep->converged = sm.flags[0]; rep->diverged = sm.flags[1]; rep->has_nan = sm.flags[2]; }
  }
  __syncthreads();  // (ex is reused by the next finished unit / the next prologue)
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, kMinCtas) lins_ieskf_kernel(const __grid_constant__ BatchView bv,
                                                                     const __grid_constant__ KParams kp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int S = bv.nslots, Q = bv.qtile, NQ = S * Q;
  CtaMem& cta = *reinterpret_cast<CtaMem*>(smem_raw);
  Smem* slots = reinterpret_cast<Smem*>(smem_raw + align16(sizeof(CtaMem)));
  const size_t slot_stride = align16(sizeof(Smem));
  auto slot_at = [&](int s) -> Smem& { return *reinterpret_cast<Smem*>(reinterpret_cast<unsigned char*>(slots) + s * slot_stride); };
  static_assert(sizeof(Smem) % 16 == 0, "slots are indexed as an array");
  const size_t nvw_all = (size_t)S * (Q / 32);
  unsigned char* p = smem_raw + align16(sizeof(CtaMem)) + S * slot_stride;
  PassBuffers pb;
  pb.nvw = Q / 32;
  pb.wacc = reinterpret_cast<double*>(p); p += align16(nvw_all * kNAcc * sizeof(double));
  pb.wcnt = reinterpret_cast<int*>(p); p += align16(nvw_all * 2 * sizeof(int));
  unsigned char* qbase = bv.qscratch ? bv.qscratch + (size_t)blockIdx.x * bv.qscratch_stride : p;
  pb.qpt = reinterpret_cast<float4*>(qbase);
  pb.sel = pb.qpt + NQ;
  pb.qa = pb.sel + NQ;
  pb.qw = reinterpret_cast<int4*>(pb.qa + NQ);
  pb.qref = reinterpret_cast<float4*>(pb.qw + NQ);
  pb.qref2 = pb.qref + NQ;
  pb.qext = pb.qref2 + NQ;
  pb.key = reinterpret_cast<unsigned long long*>(pb.qext + NQ);
  pb.pos = reinterpret_cast<int*>(pb.key + NQ);
  pb.qccr = pb.pos + 3 * NQ;
  pb.wl = pb.qccr + NQ;
  const int tid = threadIdx.x, warp = tid >> 5;
  const double sig2 = kp.lidar_std * kp.lidar_std;
  (void)slot_at;
  if (MODE == MODE_ICP_REDUCE && bv.icp_done && *bv.icp_done) return;  // the queued tail of a converged ICP loop (uniform)

  if (tid == 0) {
    mbar_init(&cta.mbar, 1);
    cta.phase = 0;
    cta.wl_n[0] = 0; cta.wl_n[1] = 0; cta.wl_tn[0] = 0; cta.wl_tn[1] = 0; cta.wl_thead[0] = 0; cta.wl_thead[1] = 0; cta.wl_head[0] = 0; cta.wl_head[1] = 0; cta.dbg[0] = 0; cta.dbg[1] = 0;  // (reset after every pass)
    cta.exhausted = 0; cta.any_finished = 0;
    for (int s = 0; s < S; ++s) { slots[s].scan = -1; slots[s].fresh = 0; slots[s].run = 0; slots[s].finished = 0; }
    fence_mbar_init();
    cta.tlast = clock64();
  }
  __syncthreads();
  const long long t_cta0 = clock64();

  for (;;) {
    // ---- claim units for the free slots ------------------------------------------------------------------------------
    if (tid == 0) {
      int nact = 0, nfresh = 0;
      for (int s = 0; s < S; ++s) {
        Smem& sm = slots[s];
        if (sm.scan < 0 && !cta.exhausted) {
          const int u = atomicAdd(bv.work_counter, 1);
          if (u < bv.n_scans) { sm.scan = u; sm.fresh = 1; } else cta.exhausted = 1;
        }
        if (sm.scan >= 0) ++nact;
        if (sm.fresh) ++nfresh;
      }
      cta.n_active = nact; cta.any_fresh = nfresh;
      cta.pass_first = 0;
      if (bv.timers) {  // diagnostics: does this pass contain a unit's first pass (every query searches)?
        int nrun = 0;
        for (int s = 0; s < S; ++s) { if (slots[s].fresh || (slots[s].run && slots[s].first_pass)) cta.pass_first = 1; if (slots[s].scan >= 0) ++nrun; }
        atomicAdd((unsigned long long*)&bv.timers[cta.pass_first ? 62 : 30], 1ull);
        atomicAdd((unsigned long long*)&bv.timers[cta.pass_first ? 63 : 31], (unsigned long long)nrun);
      }
    }
    __syncthreads();
    if (cta.n_active == 0) {
      if (bv.timers && tid == 0) {  // busy time of this CTA: sum and max over CTAs give the tail imbalance
        const unsigned long long busy = (unsigned long long)(clock64() - t_cta0);
        atomicAdd((unsigned long long*)&bv.timers[26], busy);
        atomicMax((unsigned long long*)&bv.timers[27], busy);
      }
      break;
    }
    LINS_TICK(1);
    if (cta.any_fresh) {
      for (int s = 0; s < S; ++s)
        if (slots[s].fresh) unit_prologue<MODE>(cta, slots[s], bv, kp, pb, s);
      if (tid == 0) {
        int leg = 0, idx = 0;
        for (int s = 0; s < S; ++s) if (slots[s].scan >= 0) { if (slots[s].az_ok) idx = 1; else leg = 1; }
        cta.any_legacy = leg; cta.any_indexed = idx;
      }
      __syncthreads();
      LINS_TICK(0);
    }

    // ---- one pass of every resident unit (StateEstimator.hpp:475-581) -----------------------------------------------
    association_pass<MODE>(cta, slots, bv, kp, pb);
    if (MODE == MODE_IESKF) {
      if (warp < S && slots[warp].run) unit_tail(cta, slots[warp], bv, kp, pb, warp, sig2);
    } else {
      if (warp < S && slots[warp].run) {  // single-pass modes: the 28 sums + counts are the result
        Smem& sm = slots[warp];
        const int lane = tid & 31;
        const int scan = sm.scan;
        finish_acc_warp0(sm, pb.wacc + (size_t)warp * pb.nvw * kNAcc, pb.wcnt + (size_t)warp * pb.nvw * 2, (sm.ns + sm.nc + 31) >> 5);
        if (bv.accum) {
          if (lane < kNAcc) bv.accum[(size_t)scan * 32 + lane] = sm.acc[lane];
          if (lane == 28) bv.accum[(size_t)scan * 32 + 28] = (double)sm.cnt[0];
          if (lane == 29) bv.accum[(size_t)scan * 32 + 29] = (double)sm.cnt[1];
        }
        __syncwarp();
        if (lane == 0) { sm.scan = -1; sm.run = 0; }
      }
    }
    __syncthreads();
    LINS_TICK(8);
    if (MODE == MODE_IESKF && cta.any_finished) {
      for (int s = 0; s < S; ++s) {
        if (!slots[s].finished) continue;  // (uniform: shared flag)
        unit_exit(cta, slots[s], bv, sig2);
        if (tid == 0) { slots[s].scan = -1; slots[s].finished = 0; slots[s].run = 0; }
      }
      if (tid == 0) cta.any_finished = 0;
      __syncthreads();
      LINS_TICK(9);
    }
  }
}


// Split "Jacobian kernel", shuffle-fold version (kept for A/B runs, LINS_JAC_VARIANT=2; the default is the tensor-core fold
// of lins_jacobian.cu).  SURVEY.md §8(d) unit U1; rows A5-A9 form B given the correspondence IDs.
// One warp per scan: streams the scan's queries (16 B each, coalesced) and IDs (12 / 8 B, coalesced), gathers
// the 3 / 2 matched targets (16 B each), recomputes de-skew, residual, weight and Jacobian row.  Every trip of 32
// queries is folded across the warp straight away (warp_fold_row: lane e owns sum e), so a lane carries ONE running
// sum instead of 28 accumulators: few registers, many resident warps to hide the dependent ID -> target gathers; the
// next trip's query and IDs are loaded before the current trip's arithmetic.  Nothing is staged in shared memory: with
// >= 4096 resident scans the working set exceeds L2 and the kernel is bound by HBM traffic + f64 issue.
template <int kJacThreads, int kJacMinBlocks>
__global__ void __launch_bounds__(kJacThreads, kJacMinBlocks) lins_jacobian_kernel(const __grid_constant__ BatchView bv,
                                                                                const __grid_constant__ KParams kp) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * kJacThreads) >> 5;
  for (int scan = (blockIdx.x * kJacThreads + threadIdx.x) >> 5; scan < bv.n_scans; scan += warps_per_grid) {
    // per-scan constants (every lane computes the same values)
    const double* st = bv.state_in + (size_t)scan * 20;
    const double rn0 = st[0], rn1 = st[1], rn2 = st[2];
    q4 q; q.x = st[6]; q.y = st[7]; q.z = st[8]; q.w = st[9];
    const d3 phi = Quat2axis(q);
    const m3 R = qtoR(q);
    const int qs0 = bv.qs_off[scan], ns = bv.qs_off[scan + 1] - qs0;
    const int qc0 = bv.qc_off[scan], nc = bv.qc_off[scan + 1] - qc0;
    const float4* __restrict__ tgtS = bv.ts + bv.ts_off[scan];
    const float4* __restrict__ tgtC = bv.tc + bv.tc_off[scan];
    const int Ts = bv.ts_off[scan + 1] - bv.ts_off[scan], Tc = bv.tc_off[scan + 1] - bv.tc_off[scan];
    const bool weighted = kp.iter0 >= kp.icp_freq;
    double total = 0.0;  // lane e: running sum of entry e
    int cs = 0, cc = 0;
    auto fetch = [&](int i, float4& p, int& i1, int& i2, int& i3) {
      p = make_float4(0.f, 0.f, 0.f, 0.f); i1 = -1; i2 = -1; i3 = -1;
      if (i < ns) {
        p = __ldg(bv.qs + qs0 + i);
        const int* id = bv.ind_s + 3 * (size_t)(qs0 + i);
        i1 = __ldg(id); i2 = __ldg(id + 1); i3 = __ldg(id + 2);
      } else if (i < ns + nc) {
        p = __ldg(bv.qc + qc0 + (i - ns));
        const int* id = bv.ind_c + 2 * (size_t)(qc0 + (i - ns));
        i1 = __ldg(id); i2 = __ldg(id + 1);
      }
    };
    float4 pn; int n1, n2, n3;
    fetch(lane, pn, n1, n2, n3);
    for (int i0 = 0; i0 < ns + nc; i0 += 32) {
      const int i = i0 + lane;
      const float4 p = pn;
      const int i1 = n1, i2 = n2, i3 = n3;
      const bool surf = i < ns;
      // gathers of this trip, then the next trip's streaming loads: both in flight during the arithmetic below
      float4 t1 = make_float4(0.f, 0.f, 0.f, 0.f), t2 = t1, t3 = t1;
      bool have = false;
      if (surf) {
        if (i2 >= 0 && i3 >= 0 && i1 >= 0 && i1 < Ts && i2 < Ts && i3 < Ts) { t1 = __ldg(&tgtS[i1]); t2 = __ldg(&tgtS[i2]); t3 = __ldg(&tgtS[i3]); have = true; }
      } else if (i < ns + nc) {
        if (i2 >= 0 && i1 >= 0 && i1 < Tc && i2 < Tc) { t1 = __ldg(&tgtC[i1]); t2 = __ldg(&tgtC[i2]); have = true; }
      }
      fetch(i + 32, pn, n1, n2, n3);
      double g[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, r = 0.0;
      bool ok = false;
      if (have) {
        // A2 de-skew
        const float fi = p.w - (float)((int)p.w);
        const double s = (1.f / kp.scan_period) * fi;
        const q4 rq = axis2Quat(mk3(s * phi.x, s * phi.y, s * phi.z));
        const d3 rp = qrot(rq, mk3(p.x, p.y, p.z));
        float4 sel;
        sel.x = (float)(rp.x + s * rn0); sel.y = (float)(rp.y + s * rn1); sel.z = (float)(rp.z + s * rn2); sel.w = p.w;
        float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
        ok = surf ? plane_residual(sel, t1, t2, t3, weighted, coeff) : line_residual(sel, t1, t2, weighted, coeff);
        if (ok) jacobian_row(p, coeff, R.m, kp.lidar_scale, g, r);
      }
      total += warp_fold_row(g, r);
      cs += __popc(__ballot_sync(0xffffffffu, ok && surf));
      cc += __popc(__ballot_sync(0xffffffffu, ok && !surf));
    }
    if (lane < kNAcc) bv.accum[(size_t)scan * 32 + lane] = total;
    if (lane == 28) bv.accum[(size_t)scan * 32 + 28] = (double)cs;
    if (lane == 29) bv.accum[(size_t)scan * 32 + 29] = (double)cc;
  }
}

// F1: transformToEnd (StateEstimator.hpp:1083-1101) of a packed cloud, in place on device.
__global__ void lins_transform_to_end_kernel(float4* __restrict__ pts, int n, const double* __restrict__ lin,
                                             double scan_period, lins_point* __restrict__ out32) {
  __shared__ double sphi[3], srn[3], sq[4];
  if (threadIdx.x == 0) {
    q4 q; q.x = lin[6]; q.y = lin[7]; q.z = lin[8]; q.w = lin[9];
    d3 phi = Quat2axis(q);
    sphi[0] = phi.x; sphi[1] = phi.y; sphi[2] = phi.z;
    srn[0] = lin[0]; srn[1] = lin[1]; srn[2] = lin[2];
    sq[0] = q.x; sq[1] = q.y; sq[2] = q.z; sq[3] = q.w;
  }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  float fi = p.w - (float)((int)p.w);
  double s = (1.f / scan_period) * fi;
  q4 r = axis2Quat(mk3(s * sphi[0], s * sphi[1], s * sphi[2]));
  d3 P1 = add3(qrot(r, mk3(p.x, p.y, p.z)), mk3(s * srn[0], s * srn[1], s * srn[2]));
  q4 q; q.x = sq[0]; q.y = sq[1]; q.z = sq[2]; q.w = sq[3];
  d3 P2 = qrot(qinverse(q), sub3(P1, mk3(srn[0], srn[1], srn[2])));
  p.x = (float)P2.x; p.y = (float)P2.y; p.z = (float)P2.z;
  pts[i] = p;
  if (out32) {  // full pcl::PointXYZI records for a straight D2H into the caller's cloud
    float4* o = reinterpret_cast<float4*>(out32 + i);
    o[0] = make_float4(p.x, p.y, p.z, 1.0f);
    o[1] = make_float4(p.w, 0.f, 0.f, 0.f);
  }
}

// pcl::PointXYZI records (32 B) -> packed (x, y, z, intensity) (16 B) on the device: the upload path of clouds that
// sit in caller-pinned host memory (DMA of the raw records, no host pass over the points).
__global__ void lins_pack_points_kernel(const float4* __restrict__ raw, float4* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(raw + 2 * i);       // x y z pad
    const float b = __ldcs(&raw[2 * i + 1].x);  // intensity
    out[i] = make_float4(a.x, a.y, a.z, b);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = std::max<size_t>(n, 16);
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = std::max<size_t>(n, 16);
    cudaError_t e = cudaMallocHost(&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Resident {  // one resident batch (device) + its pinned staging (host)
  int n = 0;
  size_t nqs = 0, nqc = 0, nts = 0, ntc = 0;
  int max_q = 0;
  DevBuf<float4> qs, qc, ts, tc, az_s, az_c;
  DevBuf<float4> raw;              // raw 32-B records of clouds uploaded straight from caller-pinned memory (2 float4 per point)
  DevBuf<unsigned char> qscratch;  // per-CTA per-query arrays of units too large for shared memory
  int max_ts = 0, max_tc = 0;
  DevBuf<int> qs_off, qc_off, ts_off, tc_off, ind_s, ind_c, counter;
  DevBuf<long long> timers;
  DevBuf<double> jac_part;         // split Jacobian kernel: per-(unit, part) sums and arrival counters
  DevBuf<int> jac_cnt;
  DevBuf<IcpState> icp;            // pose-update state of the ICP fallback loop (lins_gpu_estimate_transform)
  PinBuf<IcpState> h_icp;
  DevBuf<double> state_in, cov_in, state_out, cov_out, accum;
  DevBuf<lins_scan_result> results;
  DevBuf<lins_report> reports;
  DevBuf<float> sel_s, sel_c, coeff_s, coeff_c;
  DevBuf<unsigned char> mask_s, mask_c;
  PinBuf<float4> h_pts;       // staging for all four clouds, back to back
  PinBuf<int> h_off;          // 4 x (n+1)
  PinBuf<double> h_state, h_cov, h_state_out, h_cov_out, h_accum;
  PinBuf<lins_scan_result> h_results;
  PinBuf<lins_report> h_reports;
  void release() {
    qs.release(); qc.release(); ts.release(); tc.release(); az_s.release(); az_c.release(); raw.release(); qscratch.release(); qs_off.release(); qc_off.release(); ts_off.release();
    tc_off.release(); ind_s.release(); ind_c.release(); counter.release(); state_in.release(); cov_in.release();
    state_out.release(); cov_out.release(); accum.release(); results.release(); reports.release(); sel_s.release();
    sel_c.release(); coeff_s.release(); coeff_c.release(); mask_s.release(); mask_c.release(); h_pts.release();
    h_off.release(); h_state.release(); h_cov.release(); h_state_out.release(); h_cov_out.release();
    h_accum.release(); h_results.release(); h_reports.release(); icp.release(); h_icp.release(); jac_part.release(); jac_cnt.release();
  }
};

}  // namespace

struct lins_ctx {
  HostPool pool;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  lins_params prm;
  std::string err;
  int sm_count = 148;
  int max_smem_optin = 0;
  int64_t launches = 0;
  int64_t upload_raw_points = 0, upload_packed_points = 0;  // cumulative split of lins_gpu_batch_upload (raw DMA vs host pack)
  Resident batch;   // lins_gpu_batch_* working set
  Resident single;  // lins_gpu_ieskf / associate / estimate_transform (n = 1)
  // the single-scan map: "last" clouds (walks + tripods) and the clouds the 1-NN index was built on
  DevBuf<float4> map_s, map_c, tree_s, tree_c;
  DevBuf<int> map_off;  // 4 x 2 ints: [0,ns][0,nc][0,tns][0,tnc]
  int map_ns = -1, map_nc = -1, tree_ns = -1, tree_nc = -1;
  bool tree_is_map = true;
  bool timers_on = false;
  bool verbose = false;  // LINS_VERBOSE: print the launch configuration
  int force_slots = 0;  // tuning knob (LINS_SLOTS): resident units per CTA
  DevBuf<float4> tmp_pts;
  DevBuf<lins_point> tmp_out;  // transformed clouds as full PointXYZI records (update_map read-back)
  cudaEvent_t tmp_ev = nullptr; // recorded after the last H2D copies that read h_tmp
  PinBuf<lins_point> h_out;     // pinned staging of the update_map read-back
  DevBuf<double> tmp_lin;
  PinBuf<float4> h_tmp;
  // row F2 (scan-to-map refinement): the map clouds, the current feature clouds, search / reduction scratch
  struct MapState {
    DevBuf<float4> map_c, map_s, q_c, q_s;
    int n_map_c = -1, n_map_s = -1;
    struct Grid {  // hashed uniform grid of one map cloud (lins_map.cuh)
      DevBuf<float4> sorted;
      DevBuf<int> start, count, cursor;
      lins_map::GridIndex index;
      int n = 0;
      void release() { sorted.release(); start.release(); count.release(); cursor.release(); }
    } grid_c, grid_s;
    DevBuf<lins_map::MapLoopState> loop;   // device-resident state of one scan2map call
    PinBuf<lins_map::MapLoopState> h_loop;
    DevBuf<lins_map::PassConsts> consts;   // sin / cos + translation of the pass being run
    DevBuf<float> part_d;
    DevBuf<int> part_i;
    DevBuf<double> partial;
    PinBuf<double> h_partial;
    DevBuf<int32_t> knn_c, knn_s;
    DevBuf<float> coeff_c, coeff_s;
    DevBuf<uint8_t> mask_c, mask_s;
    void release() {
      grid_c.release(); grid_s.release(); loop.release(); h_loop.release(); consts.release();
      map_c.release(); map_s.release(); q_c.release(); q_s.release(); part_d.release(); part_i.release(); partial.release();
      h_partial.release(); knn_c.release(); knn_s.release(); coeff_c.release(); coeff_s.release(); mask_c.release(); mask_s.release();
    }
  } mp;
};

namespace {

int fail(lins_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
  }
  return code;
}
#define CK(call)                                                     \
  do {                                                               \
    cudaError_t _e = (call);                                         \
    if (_e != cudaSuccess) return fail(ctx, LINS_E_CUDA, #call, _e); \
  } while (0)

// pcl::PointXYZI (32 B) -> (x, y, z, intensity) (16 B).  The destination is pinned staging that the copy engine
// reads next and the host never reads back: SSE2 (x86-64 baseline) with non-temporal stores.
inline void pack_into(float4* dst, const lins_point* src, int n) {
#if defined(__SSE2__)
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    for (int i = 0; i < n; ++i) {
      const __m128 a = _mm_loadu_ps(&src[i].x);                         // x y z pad
      const __m128 b = _mm_load_ss(&src[i].intensity);                  // i 0 0 0
      const __m128 t = _mm_shuffle_ps(a, b, _MM_SHUFFLE(0, 0, 2, 2));   // z z i i
      _mm_stream_ps(reinterpret_cast<float*>(dst + i), _mm_shuffle_ps(a, t, _MM_SHUFFLE(2, 0, 1, 0)));  // x y z i
    }
    _mm_sfence();
    return;
  }
#endif
  for (int i = 0; i < n; ++i) dst[i] = make_float4(src[i].x, src[i].y, src[i].z, src[i].intensity);
}

KParams make_kparams(const lins_params& p, int mode, int iter0) {
  KParams k;
  k.num_iter = p.num_iter; k.icp_freq = p.icp_freq < 1 ? 1 : p.icp_freq; k.force_all_iters = p.force_all_iters;
  k.mode = mode; k.iter0 = iter0;
  k.nearest_sq = p.nearest_feature_search_sq_dist; k.lidar_std = p.lidar_std; k.lidar_scale = p.lidar_scale;
  k.scan_period = p.scan_period;
  return k;
}

int validate_params(const lins_params* p) {
  if (!p) return LINS_E_INVALID;
  if (p->num_iter < 0 || p->num_iter > LINS_MAX_ITER) return LINS_E_INVALID;
  if (p->icp_freq < 1) return LINS_E_INVALID;
  if (!(p->scan_period > 0)) return LINS_E_INVALID;
  return LINS_OK;
}

template <int MODE>
int launch_mode(lins_ctx* ctx, Resident& r, const BatchView& bv_in, const KParams& kp) {
  BatchView bv = bv_in;
  // resident units per CTA: as many as fit shared memory (per-query arrays of every slot + fixed part), but no more
  // than the batch can fill on every SM
  const int limit = ctx->max_smem_optin;
  int want = std::min(kMaxSlots, std::max(1, (bv.n_scans + ctx->sm_count - 1) / ctx->sm_count));
  if (ctx->force_slots > 0) want = std::min(kMaxSlots, ctx->force_slots);
  int S = 0;
  for (int s = want; s >= 1; --s)
    if ((int)(smem_fixed_bytes(s, bv.qtile) + query_array_bytes(s, bv.qtile)) <= limit) { S = s; break; }
  size_t smem;
  bv.qscratch = nullptr; bv.qscratch_stride = 0;
  if (S == 0) {  // even one unit's per-query arrays exceed shared memory: keep them in a per-CTA global scratch
    S = 1;
    smem = smem_fixed_bytes(1, bv.qtile);
    if ((int)smem > limit) return fail(ctx, LINS_E_TOOBIG, "unit too large for the fused kernel");
    bv.qscratch_stride = query_array_bytes(1, bv.qtile);
  } else {
    smem = smem_fixed_bytes(S, bv.qtile) + query_array_bytes(S, bv.qtile);
  }
  bv.nslots = S;
  CK(cudaFuncSetAttribute(lins_ieskf_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lins_ieskf_kernel<MODE>, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = std::min((bv.n_scans + S - 1) / S, ctx->sm_count * per_sm);
  if (grid < 1) grid = 1;
  if (bv.qscratch_stride) {
    CK(r.qscratch.reserve(bv.qscratch_stride * (size_t)grid));
    bv.qscratch = r.qscratch.p;
  }
  if (ctx->verbose) std::fprintf(stderr, "[lins_gpu] mode %d: %d threads, %zu B shared, %d CTA/SM, grid %d, slots %d qtile %d%s\n", MODE, kThreads, smem, per_sm, grid, S, bv.qtile, bv.qscratch ? " (per-query arrays in global scratch)" : "");
  CK(cudaMemsetAsync(bv.work_counter, 0, sizeof(int), ctx->stream));
  lins_ieskf_kernel<MODE><<<grid, kThreads, smem, ctx->stream>>>(bv, kp);
  CK(cudaGetLastError());
  ctx->launches += 1;
  return LINS_OK;
}

int launch(lins_ctx* ctx, Resident& r, const BatchView& bv, const KParams& kp) {
  switch (kp.mode) {
    case MODE_IESKF: return launch_mode<MODE_IESKF>(ctx, r, bv, kp);
    case MODE_ASSOC: return launch_mode<MODE_ASSOC>(ctx, r, bv, kp);
    case MODE_ICP_REDUCE: return launch_mode<MODE_ICP_REDUCE>(ctx, r, bv, kp);
  }
  return fail(ctx, LINS_E_INVALID, "bad kernel mode");
}

int choose_qtile(int max_q) {
  int q = ((max_q + 31) / 32) * 32;
  if (q < 32) q = 32;
  return q;
}

// allocate the per-batch outputs / scratch for n scans with the given query totals
int reserve_outputs(lins_ctx* ctx, Resident& r, bool want_reports, bool want_trace) {
  CK(r.state_out.reserve((size_t)r.n * 20));
  CK(r.cov_out.reserve((size_t)r.n * 324));
  CK(r.results.reserve(r.n));
  CK(r.accum.reserve((size_t)r.n * 32));
  CK(r.az_s.reserve(r.nts + 4));
  CK(r.az_c.reserve(r.ntc + 4));
  CK(r.ind_s.reserve(3 * r.nqs + 4));
  CK(r.ind_c.reserve(2 * r.nqc + 4));
  CK(r.counter.reserve(4));
  if (want_reports) CK(r.reports.reserve(r.n));
  if (want_trace) {
    CK(r.sel_s.reserve(3 * r.nqs + 4)); CK(r.sel_c.reserve(3 * r.nqc + 4));
    CK(r.coeff_s.reserve(4 * r.nqs + 4)); CK(r.coeff_c.reserve(4 * r.nqc + 4));
    CK(r.mask_s.reserve(r.nqs + 4)); CK(r.mask_c.reserve(r.nqc + 4));
  }
  return LINS_OK;
}

BatchView view_of(const Resident& r, bool reports, bool trace) {
  BatchView bv;
  std::memset(&bv, 0, sizeof(bv));
  bv.n_scans = r.n;
  bv.qs = r.qs.p; bv.qs_off = r.qs_off.p; bv.qc = r.qc.p; bv.qc_off = r.qc_off.p;
  bv.ts = r.ts.p; bv.ts_off = r.ts_off.p; bv.tc = r.tc.p; bv.tc_off = r.tc_off.p;
  bv.state_in = r.state_in.p; bv.cov_in = r.cov_in.p; bv.state_out = r.state_out.p; bv.cov_out = r.cov_out.p;
  bv.results = r.results.p; bv.reports = reports ? r.reports.p : nullptr;
  bv.ind_s = r.ind_s.p; bv.ind_c = r.ind_c.p;
  bv.az_s = r.az_s.p; bv.az_c = r.az_c.p;
  if (trace) {
    bv.sel_s = r.sel_s.p; bv.sel_c = r.sel_c.p; bv.coeff_s = r.coeff_s.p; bv.coeff_c = r.coeff_c.p;
    bv.mask_s = r.mask_s.p; bv.mask_c = r.mask_c.p;
  }
  bv.accum = r.accum.p;
  bv.work_counter = r.counter.p;
  bv.qtile = choose_qtile(r.max_q);
  return bv;
}

// Upload the queries + prior of ONE scan into ctx->single and point its targets at the resident map.
int stage_single(lins_ctx* ctx, const lins_point* surf_flat, int ns, const lins_point* corner_sharp, int nc,
                 const double* state_in, const double* cov_in, bool trace) {
  if (ctx->map_ns < 0) return fail(ctx, LINS_E_NOMAP, "lins_gpu_set_map has not been called");
  if (ns < 0 || nc < 0 || (ns > 0 && !surf_flat) || (nc > 0 && !corner_sharp)) return fail(ctx, LINS_E_INVALID, "bad query cloud");
  Resident& r = ctx->single;
  r.n = 1; r.nqs = ns; r.nqc = nc; r.max_q = ns + nc;
  CK(r.qs.reserve(ns + 1)); CK(r.qc.reserve(nc + 1));
  CK(r.qs_off.reserve(2)); CK(r.qc_off.reserve(2));
  CK(r.state_in.reserve(20)); CK(r.cov_in.reserve(324));
  CK(r.h_pts.reserve((size_t)ns + nc + 1)); CK(r.h_off.reserve(4)); CK(r.h_state.reserve(20)); CK(r.h_cov.reserve(324));
  pack_into(r.h_pts.p, surf_flat, ns);
  pack_into(r.h_pts.p + ns, corner_sharp, nc);
  r.h_off.p[0] = 0; r.h_off.p[1] = ns; r.h_off.p[2] = 0; r.h_off.p[3] = nc;
  for (int i = 0; i < 19; ++i) r.h_state.p[i] = state_in[i];
  r.h_state.p[19] = 0.0;
  if (cov_in) std::memcpy(r.h_cov.p, cov_in, sizeof(double) * 324); else std::memset(r.h_cov.p, 0, sizeof(double) * 324);
  if (ns) CK(cudaMemcpyAsync(r.qs.p, r.h_pts.p, sizeof(float4) * ns, cudaMemcpyHostToDevice, ctx->stream));
  if (nc) CK(cudaMemcpyAsync(r.qc.p, r.h_pts.p + ns, sizeof(float4) * nc, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(r.qs_off.p, r.h_off.p, sizeof(int) * 2, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(r.qc_off.p, r.h_off.p + 2, sizeof(int) * 2, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(r.state_in.p, r.h_state.p, sizeof(double) * 20, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(r.cov_in.p, r.h_cov.p, sizeof(double) * 324, cudaMemcpyHostToDevice, ctx->stream));
  r.nts = (size_t)std::max(std::max(ctx->map_ns, ctx->tree_ns), 0);
  r.ntc = (size_t)std::max(std::max(ctx->map_nc, ctx->tree_nc), 0);
  r.max_ts = std::max(ctx->map_ns, 0); r.max_tc = std::max(ctx->map_nc, 0);
  int rc = reserve_outputs(ctx, r, true, trace);
  return rc;
}

BatchView single_view(lins_ctx* ctx, bool trace) {
  BatchView bv = view_of(ctx->single, true, trace);
  bv.ts = ctx->map_s.p; bv.ts_off = ctx->map_off.p + 0;
  bv.tc = ctx->map_c.p; bv.tc_off = ctx->map_off.p + 2;
  if (!ctx->tree_is_map) {
    bv.nn_s = ctx->tree_s.p; bv.nn_s_off = ctx->map_off.p + 4;
    bv.nn_c = ctx->tree_c.p; bv.nn_c_off = ctx->map_off.p + 6;
  }
  return bv;
}

// ctx->h_tmp is pinned staging that asynchronous H2D copies read: before it is rewritten, wait for the event recorded
// after the last copies that read it (already complete in steady state: no stream synchronisation)
cudaError_t tmp_staging_wait(lins_ctx* ctx) {
  if (!ctx->tmp_ev) return cudaSuccess;
  return cudaEventSynchronize(ctx->tmp_ev);
}
cudaError_t tmp_staging_mark(lins_ctx* ctx) {
  if (!ctx->tmp_ev) { cudaError_t e = cudaEventCreateWithFlags(&ctx->tmp_ev, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
  return cudaEventRecord(ctx->tmp_ev, ctx->stream);
}

int upload_map_offsets(lins_ctx* ctx) {
  CK(ctx->map_off.reserve(8));
  int h[8] = {0, ctx->map_ns, 0, ctx->map_nc, 0, ctx->tree_ns, 0, ctx->tree_nc};
  // (h is pageable stack memory: cudaMemcpyAsync returns once it has been copied to the driver's staging buffer)
  CK(cudaMemcpyAsync(ctx->map_off.p, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  return LINS_OK;
}

}  // namespace

// =========================================================================================================
// C-ABI
// =========================================================================================================
extern "C" {

int lins_gpu_abi_version(void) { return 1; }

int lins_gpu_create(const lins_params* params, int device, void* stream, lins_ctx** out) {
  if (!out) return LINS_E_INVALID;
  *out = nullptr;
  if (validate_params(params) != LINS_OK) return LINS_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LINS_E_NODEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return LINS_E_NODEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LINS_E_NODEVICE;
  if (prop.major != 10) return LINS_E_NODEVICE;  // the fatbin holds sm_100a code only
  lins_ctx* ctx = new (std::nothrow) lins_ctx();
  if (!ctx) return LINS_E_INVALID;
  ctx->device = device;
  ctx->prm = *params;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  ctx->verbose = std::getenv("LINS_VERBOSE") != nullptr;
  if (const char* e = std::getenv("LINS_SLOTS")) ctx->force_slots = std::atoi(e);
  if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LINS_E_CUDA; }
    ctx->own_stream = true;
  }
  *out = ctx;
  return LINS_OK;
}

void lins_gpu_destroy(lins_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->batch.release(); ctx->single.release();
  ctx->map_s.release(); ctx->map_c.release(); ctx->tree_s.release(); ctx->tree_c.release(); ctx->map_off.release();
  ctx->tmp_pts.release(); ctx->tmp_lin.release(); ctx->tmp_out.release(); ctx->h_out.release(); ctx->h_tmp.release(); ctx->mp.release();
  if (ctx->tmp_ev) cudaEventDestroy(ctx->tmp_ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* lins_gpu_last_error(const lins_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int lins_gpu_set_params(lins_ctx* ctx, const lins_params* params) {
  if (!ctx) return LINS_E_INVALID;
  if (validate_params(params) != LINS_OK) return fail(ctx, LINS_E_INVALID, "bad params");
  ctx->prm = *params;
  return LINS_OK;
}

int64_t lins_gpu_launch_count(const lins_ctx* ctx) { return ctx ? ctx->launches : 0; }

int lins_gpu_set_map(lins_ctx* ctx, const lins_point* surf, int ns, const lins_point* corner, int nc) {
  if (!ctx) return LINS_E_INVALID;
  if (ns < 0 || nc < 0 || (ns > 0 && !surf) || (nc > 0 && !corner)) return fail(ctx, LINS_E_INVALID, "bad map cloud");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->map_s.reserve(ns + 1)); CK(ctx->map_c.reserve(nc + 1));
  CK(tmp_staging_wait(ctx));
  CK(ctx->h_tmp.reserve((size_t)ns + nc + 1));
  pack_into(ctx->h_tmp.p, surf, ns);
  pack_into(ctx->h_tmp.p + ns, corner, nc);
  if (ns) CK(cudaMemcpyAsync(ctx->map_s.p, ctx->h_tmp.p, sizeof(float4) * ns, cudaMemcpyHostToDevice, ctx->stream));
  if (nc) CK(cudaMemcpyAsync(ctx->map_c.p, ctx->h_tmp.p + ns, sizeof(float4) * nc, cudaMemcpyHostToDevice, ctx->stream));
  CK(tmp_staging_mark(ctx));
  ctx->map_ns = ns; ctx->map_nc = nc; ctx->tree_ns = ns; ctx->tree_nc = nc; ctx->tree_is_map = true;
  return upload_map_offsets(ctx);
}

int lins_gpu_ieskf(lins_ctx* ctx, const lins_point* surf_flat, int ns, const lins_point* corner_sharp, int nc,
                   const double* state_in, const double* cov_in, double* state_out, double* cov_out, lins_report* rep) {
  if (!ctx) return LINS_E_INVALID;
  if (!state_in || !cov_in) return fail(ctx, LINS_E_INVALID, "null prior");
  CK(cudaSetDevice(ctx->device));
  int rc = stage_single(ctx, surf_flat, ns, corner_sharp, nc, state_in, cov_in, false);
  if (rc != LINS_OK) return rc;
  Resident& r = ctx->single;
  BatchView bv = single_view(ctx, false);
  CK(cudaMemsetAsync(r.reports.p, 0, sizeof(lins_report), ctx->stream));
  rc = launch(ctx, r, bv, make_kparams(ctx->prm, MODE_IESKF, 0));
  if (rc != LINS_OK) return rc;
  CK(r.h_state_out.reserve(20)); CK(r.h_cov_out.reserve(324)); CK(r.h_reports.reserve(1));
  CK(cudaMemcpyAsync(r.h_state_out.p, r.state_out.p, sizeof(double) * 20, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(r.h_cov_out.p, r.cov_out.p, sizeof(double) * 324, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(r.h_reports.p, r.reports.p, sizeof(lins_report), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (state_out) std::memcpy(state_out, r.h_state_out.p, sizeof(double) * 19);
  if (cov_out) std::memcpy(cov_out, r.h_cov_out.p, sizeof(double) * 324);
  if (rep) *rep = r.h_reports.p[0];
  return LINS_OK;
}

int lins_gpu_associate(lins_ctx* ctx, const lins_point* surf_flat, int ns, const lins_point* corner_sharp, int nc,
                       const double* lin_state, int iter, int32_t* surf_ind, int32_t* corner_ind, float* surf_coeff,
                       float* corner_coeff, uint8_t* surf_mask, uint8_t* corner_mask, float* surf_sel, float* corner_sel) {
  if (!ctx) return LINS_E_INVALID;
  if (!lin_state || iter < 0) return fail(ctx, LINS_E_INVALID, "bad lin_state / iter");
  CK(cudaSetDevice(ctx->device));
  // pointSearch*Ind persist on device between calls (buffers are only re-allocated when they must grow)
  int rc = stage_single(ctx, surf_flat, ns, corner_sharp, nc, lin_state, nullptr, true);
  if (rc != LINS_OK) return rc;
  Resident& r = ctx->single;
  BatchView bv = single_view(ctx, true);
  rc = launch(ctx, r, bv, make_kparams(ctx->prm, MODE_ASSOC, iter));
  if (rc != LINS_OK) return rc;
  auto d2h = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (!dst || bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  };
  CK(d2h(surf_ind, r.ind_s.p, sizeof(int) * 3 * ns));
  CK(d2h(corner_ind, r.ind_c.p, sizeof(int) * 2 * nc));
  CK(d2h(surf_coeff, r.coeff_s.p, sizeof(float) * 4 * ns));
  CK(d2h(corner_coeff, r.coeff_c.p, sizeof(float) * 4 * nc));
  CK(d2h(surf_mask, r.mask_s.p, ns));
  CK(d2h(corner_mask, r.mask_c.p, nc));
  CK(d2h(surf_sel, r.sel_s.p, sizeof(float) * 3 * ns));
  CK(d2h(corner_sel, r.sel_c.p, sizeof(float) * 3 * nc));
  CK(cudaStreamSynchronize(ctx->stream));
  return LINS_OK;
}

// ---- batched mode ----------------------------------------------------------------------------------------
int lins_gpu_batch_upload(lins_ctx* ctx, const lins_batch_desc* b) {
  if (!ctx) return LINS_E_INVALID;
  if (!b || b->n_scans < 0) return fail(ctx, LINS_E_INVALID, "bad batch");
  CK(cudaSetDevice(ctx->device));
  Resident& r = ctx->batch;
  const int n = b->n_scans;
  CK(cudaStreamSynchronize(ctx->stream));  // the pinned staging of a previous upload may still be in flight
  r.n = n;
  if (n == 0) return LINS_OK;
  const int32_t* offs[4] = {b->surf_flat_off, b->corner_sharp_off, b->surf_less_flat_off, b->corner_less_sharp_off};
  const lins_point* pts[4] = {b->surf_flat, b->corner_sharp, b->surf_less_flat, b->corner_less_sharp};
  for (int k = 0; k < 4; ++k) {
    if (!offs[k]) return fail(ctx, LINS_E_INVALID, "null offsets");
    if (offs[k][0] != 0) return fail(ctx, LINS_E_INVALID, "offsets must start at 0");
    for (int i = 0; i < n; ++i) if (offs[k][i + 1] < offs[k][i]) return fail(ctx, LINS_E_INVALID, "offsets must be non-decreasing");
    if (offs[k][n] > 0 && !pts[k]) return fail(ctx, LINS_E_INVALID, "null cloud");
  }
  if (!b->state_in || !b->cov_in) return fail(ctx, LINS_E_INVALID, "null prior");
  if (b->point_format != LINS_POINTS_XYZI32 && b->point_format != LINS_POINTS_PACKED16) return fail(ctx, LINS_E_INVALID, "bad point_format");
  const bool packed16 = b->point_format == LINS_POINTS_PACKED16;  // the clouds are already (x, y, z, intensity) float4 records
  r.nqs = offs[0][n]; r.nqc = offs[1][n]; r.nts = offs[2][n]; r.ntc = offs[3][n];
  r.max_q = 0; r.max_ts = 0; r.max_tc = 0;
  for (int i = 0; i < n; ++i) {
    r.max_q = std::max(r.max_q, (offs[0][i + 1] - offs[0][i]) + (offs[1][i + 1] - offs[1][i]));
    r.max_ts = std::max(r.max_ts, offs[2][i + 1] - offs[2][i]);
    r.max_tc = std::max(r.max_tc, offs[3][i + 1] - offs[3][i]);
  }
  const size_t total = r.nqs + r.nqc + r.nts + r.ntc;
  CK(r.h_pts.reserve(total + 1)); CK(r.h_off.reserve(4 * (size_t)(n + 1)));
  CK(r.h_state.reserve((size_t)n * 20)); CK(r.h_cov.reserve((size_t)n * 324));
  CK(r.qs.reserve(r.nqs + 1)); CK(r.qc.reserve(r.nqc + 1)); CK(r.ts.reserve(r.nts + 1)); CK(r.tc.reserve(r.ntc + 1));
  CK(r.qs_off.reserve(n + 1)); CK(r.qc_off.reserve(n + 1)); CK(r.ts_off.reserve(n + 1)); CK(r.tc_off.reserve(n + 1));
  CK(r.state_in.reserve((size_t)n * 20)); CK(r.cov_in.reserve((size_t)n * 324));
  // pack 32-B PointXYZI -> 16-B float4 while copying into pinned staging (the copy is needed anyway: user
  // buffers are pageable), so PCIe moves half the bytes.  The pack is spread over host threads in 64 K-point
  // slices; each slice's H2D copy is queued as soon as the slice is packed, so packing and PCIe overlap.
  float4* hp = r.h_pts.p;
  size_t seg[5] = {0, r.nqs, r.nqs + r.nqc, r.nqs + r.nqc + r.nts, total};
  float4* dsts[4] = {r.qs.p, r.qc.p, r.ts.p, r.tc.p};
  int* doffs[4] = {r.qs_off.p, r.qc_off.p, r.ts_off.p, r.tc_off.p};
  // Clouds in caller-PINNED host memory (cudaHostAlloc / cudaHostRegister) can also go the other way: the copy engine reads
  // the raw 32-B records straight from the caller's buffer and a device kernel packs them — twice the PCIe bytes, but no host
  // pass over the points.  The clouds are cut into 64 K-point slices; pack threads take slices from the front of the list
  // (pack -> pinned staging -> 16-B H2D), a feeder hands slices from the back to the copy engine as raw records.
  // LINS_UPLOAD=pack: host pack only; =direct: raw DMA for every pinned cloud; =pinned: as direct, and an unpinned cloud
  // fails the call; =hybrid: both ends at once, the feeder never more than two slices ahead (kept for experiments: it did
  // not beat the better of the two pure modes).
  // pack threads: LINS_PACK_THREADS when set (a job that runs several contexts / ranks per host divides the cores
  // among them), else half the hardware threads, at most 32
  int want_threads;
  {
    unsigned hw = std::thread::hardware_concurrency();
    want_threads = (int)std::min<unsigned>(hw ? hw / 2 : 4, 32);
    if (const char* e = std::getenv("LINS_PACK_THREADS")) { const int v = std::atoi(e); if (v >= 1) want_threads = std::min(v, 64); }
  }
  const char* mode = std::getenv("LINS_UPLOAD");
  // default: host pack when this context has >= 8 pack threads to itself (measured, 2 GPUs x 3 contexts: 10 threads each
  // 6.1 M it/s per GPU; raw DMA 3.8 M whatever the threads; the two-ended split with 2-5 threads 3.3-3.6 M), raw DMA otherwise
  const bool p16 = b->point_format == LINS_POINTS_PACKED16;  // (16-B records from pinned memory: always straight DMA)
  const bool mode_pack = mode ? std::strcmp(mode, "pack") == 0 : (want_threads >= 8 && !p16);
  const bool mode_direct = mode ? (std::strcmp(mode, "direct") == 0 || std::strcmp(mode, "pinned") == 0) : (want_threads < 8 || p16);
  bool pinned[4] = {false, false, false, false};
  bool any_pinned = false;
  for (int k = 0; k < 4 && !mode_pack; ++k) {
    if (seg[k + 1] == seg[k]) continue;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, pts[k]) == cudaSuccess && at.type == cudaMemoryTypeHost) { pinned[k] = true; any_pinned = true; }
    else cudaGetLastError();
    if (!pinned[k] && mode && std::strcmp(mode, "pinned") == 0) return fail(ctx, LINS_E_INVALID, "LINS_UPLOAD=pinned but a cloud is not in pinned host memory");
  }
  if (any_pinned && !packed16) CK(r.raw.reserve(2 * total + 2));
  {
    struct Slice { int k; size_t a, b; };
    std::vector<Slice> slices;  // unpinned clouds first: only the pack threads may take those
    const size_t SL = 1u << 16;
    size_t n_unpinned = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < 4; ++k) {
        if ((pass == 0) == pinned[k]) continue;
        for (size_t a = 0; a < seg[k + 1] - seg[k]; a += SL) slices.push_back(Slice{k, a, std::min(a + SL, seg[k + 1] - seg[k])});
        if (pass == 0) n_unpinned = slices.size();
      }
    // one list, two ends: lo = next slice for the pack threads, hi = one past the last slice not yet taken by the DMA feeder
    std::mutex mu;
    size_t lo = 0, hi = slices.size();
    std::atomic<int> cuda_err(0);
    const int device = ctx->device;
    cudaStream_t stream = ctx->stream;
    auto take_front = [&](size_t& i) { std::lock_guard<std::mutex> g(mu); if (lo >= hi) return false; i = lo++; return true; };
    auto take_back = [&](size_t& i) { std::lock_guard<std::mutex> g(mu); if (lo >= hi || hi - 1 < n_unpinned) return false; i = --hi; return true; };
    auto worker = [&]() {
      cudaSetDevice(device);
      size_t i;
      while (take_front(i)) {
        const Slice& sl = slices[i];
        if (packed16) std::memcpy(hp + seg[sl.k] + sl.a, reinterpret_cast<const float4*>(pts[sl.k]) + sl.a, sizeof(float4) * (sl.b - sl.a));
        else pack_into(hp + seg[sl.k] + sl.a, pts[sl.k] + sl.a, (int)(sl.b - sl.a));
        cudaError_t e = cudaMemcpyAsync(dsts[sl.k] + sl.a, hp + seg[sl.k] + sl.a, sizeof(float4) * (sl.b - sl.a), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) cuda_err.store((int)e);
      }
    };
    const bool feed_raw = any_pinned && !mode_pack;
    int nthr = mode_direct && n_unpinned == 0 ? 0 : (int)std::min<size_t>((size_t)want_threads, std::max<size_t>(slices.size(), 1));
    // the DMA feeder (this thread): raw slices from the back, at most two in flight (all of them at once with LINS_UPLOAD=direct)
    auto feeder = [&]() {
      if (!feed_raw) return;
      cudaEvent_t ev[2] = {nullptr, nullptr};
      if (!mode_direct) for (auto& e : ev) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cuda_err.store((int)cudaGetLastError()); return; }
      size_t i;
      int turn = 0;
      while (take_back(i)) {
        const Slice& sl = slices[i];
        const size_t cnt = sl.b - sl.a;
        if (!mode_direct) cudaEventSynchronize(ev[turn]);  // (a never-recorded event is complete)
        cudaError_t e;
        if (packed16) {  // already in the device's format: one DMA, nothing else
          e = cudaMemcpyAsync(dsts[sl.k] + sl.a, reinterpret_cast<const float4*>(pts[sl.k]) + sl.a, sizeof(float4) * cnt, cudaMemcpyHostToDevice, stream);
        } else {
          float4* rawk = r.raw.p + 2 * (seg[sl.k] + sl.a);
          e = cudaMemcpyAsync(rawk, pts[sl.k] + sl.a, sizeof(lins_point) * cnt, cudaMemcpyHostToDevice, stream);
          const int blocks = (int)std::min<size_t>((cnt + 255) / 256, (size_t)ctx->sm_count * 4);
          lins_pack_points_kernel<<<blocks, 256, 0, stream>>>(rawk, dsts[sl.k] + sl.a, cnt);
          if (e == cudaSuccess) e = cudaGetLastError();
          if (e == cudaSuccess) ctx->launches += 1;
        }
        if (e != cudaSuccess) { cuda_err.store((int)e); break; }
        if (!mode_direct) { cudaEventRecord(ev[turn], stream); turn ^= 1; }
      }
      for (auto& e : ev) if (e) cudaEventDestroy(e);
    };
    if (nthr >= 1 && !slices.empty()) {
      // the pool runs the workers; the calling thread feeds the copy engine meanwhile (HostPool::run blocks, so the feeder is
      // the pool's first worker's prologue when only one thread is available)
      std::atomic<int> first(0);
      ctx->pool.run(nthr + (feed_raw ? 1 : 0), [&]() { if (feed_raw && first.fetch_add(1) == 0) { cudaSetDevice(device); feeder(); } else worker(); });
    } else {
      feeder();
    }
    if (cuda_err.load() != 0) return fail(ctx, LINS_E_CUDA, "H2D copy of a slice", (cudaError_t)cuda_err.load());
    // what went which way (lins_gpu_batch_upload_stats): slices [n_unpinned.., lo) were packed by the host, [hi, end) went raw
    size_t raw_pts = 0;
    for (size_t i = hi; i < slices.size(); ++i) raw_pts += slices[i].b - slices[i].a;
    if (packed16) raw_pts = 0;  // (16-B records either way)
    ctx->upload_raw_points += (int64_t)raw_pts;
    ctx->upload_packed_points += (int64_t)(total - raw_pts);
  }
  for (int k = 0; k < 4; ++k) std::memcpy(r.h_off.p + (size_t)k * (n + 1), offs[k], sizeof(int) * (n + 1));
  for (int i = 0; i < n; ++i) {
    std::memcpy(r.h_state.p + (size_t)i * 20, b->state_in + (size_t)i * 19, sizeof(double) * 19);
    r.h_state.p[(size_t)i * 20 + 19] = 0.0;
  }
  std::memcpy(r.h_cov.p, b->cov_in, sizeof(double) * 324 * (size_t)n);
  for (int k = 0; k < 4; ++k)
    CK(cudaMemcpyAsync(doffs[k], r.h_off.p + (size_t)k * (n + 1), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(r.state_in.p, r.h_state.p, sizeof(double) * 20 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(r.cov_in.p, r.h_cov.p, sizeof(double) * 324 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  return reserve_outputs(ctx, r, true, false);
}

int lins_gpu_batch_upload_stats(lins_ctx* ctx, int64_t* packed_points, int64_t* raw_points) {
  if (!ctx) return LINS_E_INVALID;
  if (packed_points) *packed_points = ctx->upload_packed_points;
  if (raw_points) *raw_points = ctx->upload_raw_points;
  return LINS_OK;
}

int lins_gpu_batch_run(lins_ctx* ctx) {
  if (!ctx) return LINS_E_INVALID;
  Resident& r = ctx->batch;
  if (r.n <= 0) return fail(ctx, LINS_E_INVALID, "no resident batch");
  CK(cudaSetDevice(ctx->device));
  BatchView bv = view_of(r, true, false);
  if (ctx->timers_on) {
    CK(r.timers.reserve(64));
    CK(cudaMemsetAsync(r.timers.p, 0, sizeof(long long) * 64, ctx->stream));
    bv.timers = r.timers.p;
  }
  return launch(ctx, r, bv, make_kparams(ctx->prm, MODE_IESKF, 0));
}

int lins_gpu_batch_download(lins_ctx* ctx, double* state_out, double* cov_out, lins_scan_result* results,
                            lins_report* reports) {
  if (!ctx) return LINS_E_INVALID;
  Resident& r = ctx->batch;
  if (r.n <= 0) return fail(ctx, LINS_E_INVALID, "no resident batch");
  CK(cudaSetDevice(ctx->device));
  const size_t n = r.n;
  if (state_out) { CK(r.h_state_out.reserve(n * 20)); CK(cudaMemcpyAsync(r.h_state_out.p, r.state_out.p, sizeof(double) * 20 * n, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (cov_out) { CK(r.h_cov_out.reserve(n * 324)); CK(cudaMemcpyAsync(r.h_cov_out.p, r.cov_out.p, sizeof(double) * 324 * n, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (results) { CK(r.h_results.reserve(n)); CK(cudaMemcpyAsync(r.h_results.p, r.results.p, sizeof(lins_scan_result) * n, cudaMemcpyDeviceToHost, ctx->stream)); }
  if (reports) { CK(r.h_reports.reserve(n)); CK(cudaMemcpyAsync(r.h_reports.p, r.reports.p, sizeof(lins_report) * n, cudaMemcpyDeviceToHost, ctx->stream)); }
  CK(cudaStreamSynchronize(ctx->stream));
  if (state_out) for (size_t i = 0; i < n; ++i) std::memcpy(state_out + i * 19, r.h_state_out.p + i * 20, sizeof(double) * 19);
  if (cov_out) std::memcpy(cov_out, r.h_cov_out.p, sizeof(double) * 324 * n);
  if (results) std::memcpy(results, r.h_results.p, sizeof(lins_scan_result) * n);
  if (reports) std::memcpy(reports, r.h_reports.p, sizeof(lins_report) * n);
  return LINS_OK;
}

int lins_gpu_batch_download_indices(lins_ctx* ctx, int32_t* surf_ind, int32_t* corner_ind) {
  if (!ctx) return LINS_E_INVALID;
  Resident& r = ctx->batch;
  if (r.n <= 0) return fail(ctx, LINS_E_INVALID, "no resident batch");
  CK(cudaSetDevice(ctx->device));
  if (surf_ind && r.nqs) CK(cudaMemcpyAsync(surf_ind, r.ind_s.p, sizeof(int32_t) * 3 * r.nqs, cudaMemcpyDeviceToHost, ctx->stream));
  if (corner_ind && r.nqc) CK(cudaMemcpyAsync(corner_ind, r.ind_c.p, sizeof(int32_t) * 2 * r.nqc, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LINS_OK;
}

int lins_gpu_ieskf_batch(lins_ctx* ctx, const lins_batch_desc* batch, double* state_out, double* cov_out,
                         lins_scan_result* results) {
  int rc = lins_gpu_batch_upload(ctx, batch);
  if (rc != LINS_OK) return rc;
  if (batch->n_scans == 0) return LINS_OK;
  rc = lins_gpu_batch_run(ctx);
  if (rc != LINS_OK) return rc;
  return lins_gpu_batch_download(ctx, state_out, cov_out, results, nullptr);
}

int lins_gpu_batch_results_device(lins_ctx* ctx, void** dev_ptr, int* n_scans) {
  if (!ctx || !dev_ptr) return LINS_E_INVALID;
  *dev_ptr = ctx->batch.results.p;
  if (n_scans) *n_scans = ctx->batch.n;
  return LINS_OK;
}

// Diagnostics: enable per-phase cycle counters for lins_gpu_batch_run (enable != 0), and/or read the 32
// counters of the last run (out may be NULL).  Slot meaning: see LINS_TICK sites in lins_gpu.cu / lins_kernels.cuh.
int lins_gpu_debug_phase_cycles(lins_ctx* ctx, int enable, long long* out) {
  if (!ctx) return LINS_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  if (out) {
    if (!ctx->batch.timers.p) return fail(ctx, LINS_E_INVALID, "timers were not enabled for the last run");
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, ctx->batch.timers.p, sizeof(long long) * 64, cudaMemcpyDeviceToHost));
  }
  ctx->timers_on = enable != 0;
  return LINS_OK;
}


// ---- row F2: the mapping node's scan-to-map refinement (lidar_mapping_node.cpp:1635-1652) --------------------------------
namespace {

// pack + queue the H2D of two clouds through the pinned staging (no stream synchronisation: see tmp_staging_wait)
int map_upload2(lins_ctx* ctx, DevBuf<float4>& dst_a, const lins_point* a, int na, DevBuf<float4>& dst_b, const lins_point* b, int nb) {
  CK(dst_a.reserve((size_t)na + 1)); CK(dst_b.reserve((size_t)nb + 1));
  if (na + nb == 0) return LINS_OK;
  CK(tmp_staging_wait(ctx));
  CK(ctx->h_tmp.reserve((size_t)na + nb + 1));
  pack_into(ctx->h_tmp.p, a, na);
  pack_into(ctx->h_tmp.p + na, b, nb);
  if (na) CK(cudaMemcpyAsync(dst_a.p, ctx->h_tmp.p, sizeof(float4) * (size_t)na, cudaMemcpyHostToDevice, ctx->stream));
  if (nb) CK(cudaMemcpyAsync(dst_b.p, ctx->h_tmp.p + na, sizeof(float4) * (size_t)nb, cudaMemcpyHostToDevice, ctx->stream));
  CK(tmp_staging_mark(ctx));
  return LINS_OK;
}

lins_map::PassConsts host_pass_consts(const float* T) {  // libm sin / cos in f32, like the reference (:579-592, :1527-1532)
  lins_map::PassConsts pc;
  pc.cRoll = std::cos(T[0]); pc.sRoll = std::sin(T[0]); pc.cPitch = std::cos(T[1]); pc.sPitch = std::sin(T[1]);
  pc.cYaw = std::cos(T[2]); pc.sYaw = std::sin(T[2]); pc.tX = T[3]; pc.tY = T[4]; pc.tZ = T[5];
  pc.srx = std::sin(T[0]); pc.crx = std::cos(T[0]); pc.sry = std::sin(T[1]); pc.cry = std::cos(T[1]);
  pc.srz = std::sin(T[2]); pc.crz = std::cos(T[2]);
  return pc;
}

// which 5-NN search a pass uses: the hashed grid (exact for every point that can be accepted) or the brute-force slices
// (exact for every point).  LINS_MAP_KNN=grid|brute overrides the caller's default.
bool map_use_grid(bool dflt) {
  const char* e = std::getenv("LINS_MAP_KNN");  // (read per call: tests flip it between calls)
  return !e || !*e ? dflt : std::strcmp(e, "grid") == 0;
}

// bucket-sort one map cloud into its grid (≙ kdtree*FromMap->setInputCloud, :1637-1638)
int map_build_grid(lins_ctx* ctx, lins_ctx::MapState::Grid& g, const float4* map, int n, const lins_point* host_pts) {
  using namespace lins_map;
  g.n = n;
  if (n <= 0) return LINS_OK;
  float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f};
  for (int i = 0; i < n; ++i) {  // origin = the finite minimum (cells are addressed by hash: the extent does not matter)
    const float v[3] = {host_pts[i].x, host_pts[i].y, host_pts[i].z};
    for (int k = 0; k < 3; ++k) if (v[k] == v[k] && std::fabs(v[k]) < 1.0e30f && v[k] < mn[k]) mn[k] = v[k];
  }
  for (int k = 0; k < 3; ++k) if (!(mn[k] < 3.0e38f)) mn[k] = 0.f;
  unsigned nb = 4096;
  while (nb < 2u * (unsigned)n && nb < (1u << 24)) nb <<= 1;
  CK(g.start.reserve((size_t)nb + 2)); CK(g.count.reserve((size_t)nb + 2)); CK(g.cursor.reserve((size_t)nb + 2)); CK(g.sorted.reserve((size_t)n + 1));
  GridIndex gi;
  gi.pts = g.sorted.p; gi.start = g.start.p; gi.mask = nb - 1; gi.ox = mn[0]; gi.oy = mn[1]; gi.oz = mn[2];
  g.index = gi;
  CK(cudaMemsetAsync(g.count.p, 0, sizeof(int) * nb, ctx->stream));
  lins_grid_count_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(map, n, gi, g.count.p);
  lins_grid_scan_kernel<<<1, 1024, 0, ctx->stream>>>(g.count.p, g.start.p, g.cursor.p, (int)nb);
  lins_grid_scatter_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(map, n, gi, g.cursor.p, g.sorted.p);
  CK(cudaGetLastError());
  ctx->launches += 3;
  return LINS_OK;
}

// queue one cornerOptimization + surfOptimization pass (5-NN, fits, block partials) that reads its constants from
// m.consts (device); nothing is synchronised.  Returns the number of partial blocks.
int map_queue_pass(lins_ctx* ctx, int nc, int ns, bool dense, bool grid, const int* done, int* nblocks_out) {
  using namespace lins_map;
  lins_ctx::MapState& m = ctx->mp;
  const int qb[2] = {(nc + kKnnThreads - 1) / kKnnThreads, (ns + kKnnThreads - 1) / kKnnThreads};
  const int nq[2] = {nc, ns}, nm[2] = {std::max(m.n_map_c, 0), std::max(m.n_map_s, 0)};
  const float4* q[2] = {m.q_c.p, m.q_s.p};
  const float4* mp[2] = {m.map_c.p, m.map_s.p};
  const lins_ctx::MapState::Grid* gr[2] = {&m.grid_c, &m.grid_s};
  int slices[2], slice_len[2];
  size_t part = 0;
  for (int k = 0; k < 2; ++k) {
    // brute force: enough (query block, map slice) pairs for ~8 CTAs per SM (the scan is latency bound: profiles/r01_map_*);
    // a slice is at least 256 map points.  grid: one list per query
    int S = qb[k] > 0 ? (8 * ctx->sm_count + qb[k] - 1) / qb[k] : 1;
    S = std::max(1, std::min(S, std::min(64, (nm[k] + 255) / 256)));
    if (grid) S = 1;
    slices[k] = S;
    slice_len[k] = std::max(1, (nm[k] + S - 1) / S);
    part = std::max(part, (size_t)nq[k] * S * 5);
  }
  CK(m.part_d.reserve(part + 1)); CK(m.part_i.reserve(part + 1));
  const int nblocks = qb[0] + qb[1];
  CK(m.partial.reserve((size_t)(nblocks + 1) * (kRowAcc + 1))); CK(m.h_partial.reserve((size_t)(nblocks + 1) * (kRowAcc + 1)));
  if (dense) {
    CK(m.knn_c.reserve(5 * (size_t)nc + 1)); CK(m.knn_s.reserve(5 * (size_t)ns + 1)); CK(m.coeff_c.reserve(4 * (size_t)nc + 1));
    CK(m.coeff_s.reserve(4 * (size_t)ns + 1)); CK(m.mask_c.reserve((size_t)nc + 1)); CK(m.mask_s.reserve((size_t)ns + 1));
  }
  for (int k = 0; k < 2; ++k) {
    if (nq[k] == 0) continue;
    if (grid && nm[k] > 0)
      lins_map_knn_grid_kernel<<<(nq[k] + kGridKnnWarps - 1) / kGridKnnWarps, kGridKnnWarps * 32, 0, ctx->stream>>>(q[k], nq[k], gr[k]->index, m.consts.p, done, m.part_d.p, m.part_i.p);
    else
      lins_map_knn_kernel<<<dim3(qb[k], slices[k]), kKnnThreads, 0, ctx->stream>>>(q[k], nq[k], mp[k], nm[k], slice_len[k], m.consts.p, done, m.part_d.p, m.part_i.p);
    double* partial = m.partial.p + (size_t)(k == 0 ? 0 : qb[0]) * (kRowAcc + 1);
    if (k == 0)
      lins_map_fit_kernel<true><<<qb[k], kFitThreads, 0, ctx->stream>>>(q[k], nq[k], mp[k], slices[k], m.part_d.p, m.part_i.p, m.consts.p, done, dense ? m.knn_c.p : nullptr,
                                                                       dense ? m.coeff_c.p : nullptr, dense ? m.mask_c.p : nullptr, partial);
    else
      lins_map_fit_kernel<false><<<qb[k], kFitThreads, 0, ctx->stream>>>(q[k], nq[k], mp[k], slices[k], m.part_d.p, m.part_i.p, m.consts.p, done, dense ? m.knn_s.p : nullptr,
                                                                        dense ? m.coeff_s.p : nullptr, dense ? m.mask_s.p : nullptr, partial);
    CK(cudaGetLastError());
    ctx->launches += 2;
  }
  *nblocks_out = nblocks;
  return LINS_OK;
}

int map_stage_queries(lins_ctx* ctx, const lins_point* corner, int nc, const lins_point* surf, int ns) {
  if (nc < 0 || ns < 0 || (nc > 0 && !corner) || (ns > 0 && !surf)) return fail(ctx, LINS_E_INVALID, "bad feature clouds");
  if (ctx->mp.n_map_c < 0) return fail(ctx, LINS_E_NOMAP, "lins_gpu_map_set has not been called");
  return map_upload2(ctx, ctx->mp.q_c, corner, nc, ctx->mp.q_s, surf, ns);
}

}  // namespace

int lins_gpu_map_set(lins_ctx* ctx, const lins_point* corner, int nc, const lins_point* surf, int ns) {
  if (!ctx) return LINS_E_INVALID;
  if (nc < 0 || ns < 0 || (nc > 0 && !corner) || (ns > 0 && !surf)) return fail(ctx, LINS_E_INVALID, "bad map clouds");
  CK(cudaSetDevice(ctx->device));
  int rc = map_upload2(ctx, ctx->mp.map_c, corner, nc, ctx->mp.map_s, surf, ns);
  if (rc != LINS_OK) return rc;
  ctx->mp.n_map_c = nc; ctx->mp.n_map_s = ns;
  rc = map_build_grid(ctx, ctx->mp.grid_c, ctx->mp.map_c.p, nc, corner);
  if (rc != LINS_OK) return rc;
  return map_build_grid(ctx, ctx->mp.grid_s, ctx->mp.map_s.p, ns, surf);
}

int lins_gpu_scan2map(lins_ctx* ctx, const lins_point* corner, int nc, const lins_point* surf, int ns, float* T, lins_map_report* rep) {
  if (!ctx) return LINS_E_INVALID;
  if (!T) return fail(ctx, LINS_E_INVALID, "null transform");
  CK(cudaSetDevice(ctx->device));
  using namespace lins_map;
  lins_map_report r;
  std::memset(&r, 0, sizeof(r));
  lins_ctx::MapState& m = ctx->mp;
  if (m.n_map_c < 0) return fail(ctx, LINS_E_NOMAP, "lins_gpu_map_set has not been called");
  if (!(m.n_map_c > 10 && m.n_map_s > 100)) {  // :1636
    r.skipped = 1;
    if (rep) *rep = r;
    return LINS_OK;
  }
  int rc = map_stage_queries(ctx, corner, nc, surf, ns);
  if (rc != LINS_OK) return rc;
  // The whole iteration loop (:1640-1648) is queued up front: transformTobeMapped, matP / isDegenerate and the report live
  // on the device (MapLoopState); the first pass uses libm sin / cos of the caller's transform (bit-identical to the
  // reference's first pass), later ones the constants the LM kernel derived on the device.
  if (!m.loop.p) {  // isDegenerate / matP are members of the reference's mapping node (:226-227, :395-396): they survive the
    CK(m.loop.reserve(1));  // calls — a call whose first pass selects < 50 points keeps using the previous scan's values
    CK(cudaMemsetAsync(m.loop.p, 0, sizeof(MapLoopState), ctx->stream));
  }
  CK(m.h_loop.reserve(1)); CK(m.consts.reserve(1));
  const PassConsts pc0 = host_pass_consts(T);
  CK(cudaMemcpyAsync(m.loop.p, T, sizeof(float) * 6, cudaMemcpyHostToDevice, ctx->stream));  // (pageable sources: staged before the call returns)
  CK(cudaMemsetAsync(reinterpret_cast<char*>(m.loop.p) + offsetof(MapLoopState, done), 0, sizeof(MapLoopState) - offsetof(MapLoopState, done), ctx->stream));
  CK(cudaMemcpyAsync(m.consts.p, &pc0, sizeof(pc0), cudaMemcpyHostToDevice, ctx->stream));
  const bool grid = map_use_grid(true);
  for (int iter = 0; iter < LINS_MAP_MAX_ITER; ++iter) {
    int nblocks = 0;
    rc = map_queue_pass(ctx, nc, ns, false, grid, &m.loop.p->done, &nblocks);
    if (rc != LINS_OK) return rc;
    lins_map_lm_kernel<<<1, 32, 0, ctx->stream>>>(m.partial.p, nblocks, iter, m.loop.p, m.consts.p);
    CK(cudaGetLastError());
    ctx->launches += 1;
  }
  CK(cudaMemcpyAsync(m.h_loop.p, m.loop.p, sizeof(MapLoopState), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const MapLoopState& st = *m.h_loop.p;
  for (int i = 0; i < 6; ++i) T[i] = st.T[i];
  r.iters = st.iters; r.converged = st.converged; r.degenerate = st.isDegenerate;
  for (int i = 0; i < LINS_MAP_MAX_ITER; ++i) { r.n_sel[i] = st.n_sel[i]; r.delta_r[i] = st.delta_r[i]; r.delta_t[i] = st.delta_t[i]; }
  if (rep) *rep = r;
  return LINS_OK;
}

int lins_gpu_map_associate(lins_ctx* ctx, const lins_point* corner, int nc, const lins_point* surf, int ns, const float* T,
                           int32_t* cknn, int32_t* sknn, float* ccoeff, float* scoeff, uint8_t* cmask, uint8_t* smask) {
  if (!ctx) return LINS_E_INVALID;
  if (!T) return fail(ctx, LINS_E_INVALID, "null transform");
  CK(cudaSetDevice(ctx->device));
  int rc = map_stage_queries(ctx, corner, nc, surf, ns);
  if (rc != LINS_OK) return rc;
  lins_ctx::MapState& m = ctx->mp;
  CK(m.consts.reserve(1));
  const lins_map::PassConsts pc = host_pass_consts(T);
  CK(cudaMemcpyAsync(m.consts.p, &pc, sizeof(pc), cudaMemcpyHostToDevice, ctx->stream));
  int nblocks = 0;
  // the parity hook: brute force by default (exact neighbours for EVERY point, also those the 1 m gate rejects)
  rc = map_queue_pass(ctx, nc, ns, true, map_use_grid(false), nullptr, &nblocks);
  if (rc != LINS_OK) return rc;
  if (cknn && nc) CK(cudaMemcpyAsync(cknn, m.knn_c.p, sizeof(int32_t) * 5 * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (sknn && ns) CK(cudaMemcpyAsync(sknn, m.knn_s.p, sizeof(int32_t) * 5 * (size_t)ns, cudaMemcpyDeviceToHost, ctx->stream));
  if (ccoeff && nc) CK(cudaMemcpyAsync(ccoeff, m.coeff_c.p, sizeof(float) * 4 * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (scoeff && ns) CK(cudaMemcpyAsync(scoeff, m.coeff_s.p, sizeof(float) * 4 * (size_t)ns, cudaMemcpyDeviceToHost, ctx->stream));
  if (cmask && nc) CK(cudaMemcpyAsync(cmask, m.mask_c.p, (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
  if (smask && ns) CK(cudaMemcpyAsync(smask, m.mask_s.p, (size_t)ns, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LINS_OK;
}

int lins_gpu_host_register(void* ptr, size_t bytes) {
  if (!ptr || bytes == 0) return LINS_E_INVALID;
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? LINS_E_NODEVICE : LINS_E_CUDA; }
  return LINS_OK;
}
int lins_gpu_host_unregister(void* ptr) {
  if (!ptr) return LINS_E_INVALID;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return LINS_E_CUDA; }
  return LINS_OK;
}

int lins_gpu_sync(lins_ctx* ctx) {
  if (!ctx) return LINS_E_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return LINS_OK;
}


int lins_gpu_batch_jacobian_pass(lins_ctx* ctx, double* accum_out) {
  if (!ctx) return LINS_E_INVALID;
  Resident& r = ctx->batch;
  if (r.n <= 0) return fail(ctx, LINS_E_INVALID, "no resident batch");
  CK(cudaSetDevice(ctx->device));
  BatchView bv = view_of(r, false, false);
  bv.state_in = r.state_out.p;  // linearise at the updated state; IDs = the last iteration's
  KParams kp = make_kparams(ctx->prm, MODE_JACOBIAN, 1);
  // default: lins_jacobian.cu (FP64 tensor-core fold, multiply-add contraction); LINS_JAC_VARIANT=2: the shuffle-fold
  // kernel above (this TU, -fmad=false)
  static const int variant = [] { const char* e = std::getenv("LINS_JAC_VARIANT"); return e ? std::atoi(e) : 0; }();
  if (variant == 2) {
    const int wpb = 128 / 32;
    int grid = std::min((r.n + wpb - 1) / wpb, ctx->sm_count * 5);
    if (grid < 1) grid = 1;
    lins_jacobian_kernel<128, 5><<<grid, 128, 0, ctx->stream>>>(bv, kp);  // measured: 83 us per 5000 units (96 registers, 20 warps per SM)
  } else {
    const int P = lins_jacobian_parts(r.n, ctx->sm_count);
    if (P > 1) { CK(r.jac_part.reserve((size_t)r.n * P * 32)); CK(r.jac_cnt.reserve((size_t)r.n)); }
    const int e = lins_launch_jacobian_mma(&bv, &kp, r.n, ctx->sm_count, P, r.jac_part.p, r.jac_cnt.p, ctx->stream);
    if (e != 0) return fail(ctx, LINS_E_CUDA, "jacobian kernel launch", (cudaError_t)e);
  }
  CK(cudaGetLastError());
  ctx->launches += 1;
  if (accum_out) {
    CK(r.h_accum.reserve((size_t)r.n * 32));
    CK(cudaMemcpyAsync(r.h_accum.p, r.accum.p, sizeof(double) * 32 * (size_t)r.n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::memcpy(accum_out, r.h_accum.p, sizeof(double) * 32 * (size_t)r.n);
  }
  return LINS_OK;
}

// (A/B only, LINS_ICP_HOST_LOOP=1: the round-1 flow — one reduction launch, D2H and host 6x6 step per iteration)
static int estimate_transform_host_loop(lins_ctx* ctx, const lins_point* surf_flat, int ns, const lins_point* corner_sharp,
                                int nc, double* pose_io, int* iters_out, int* converged_out) {
  if (!ctx) return LINS_E_INVALID;
  if (!pose_io) return fail(ctx, LINS_E_INVALID, "null pose");
  CK(cudaSetDevice(ctx->device));
  using namespace lins;
  V3D t(pose_io[0], pose_io[1], pose_io[2]);
  Q4D q(pose_io[6], pose_io[3], pose_io[4], pose_io[5]);
  double lin[19];
  std::memset(lin, 0, sizeof(lin));
  linalg::Mat<6> matP{};
  bool conv = false;
  int it = 0;
  Resident& r = ctx->single;
  for (int iter = 0; iter < ctx->prm.num_iter; ++iter) {
    it = iter + 1;
    lin[0] = t.x(); lin[1] = t.y(); lin[2] = t.z();
    lin[6] = q.x(); lin[7] = q.y(); lin[8] = q.z(); lin[9] = q.w();
    int rc = stage_single(ctx, surf_flat, ns, corner_sharp, nc, lin, nullptr, false);
    if (rc != LINS_OK) return rc;
    BatchView bv = single_view(ctx, false);
    rc = launch(ctx, r, bv, make_kparams(ctx->prm, MODE_ICP_REDUCE, iter));
    if (rc != LINS_OK) return rc;
    CK(r.h_accum.reserve(32));
    CK(cudaMemcpyAsync(r.h_accum.p, r.accum.p, sizeof(double) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const double* a = r.h_accum.p;
    if (ctx->verbose) { std::fprintf(stderr, "[lins_gpu] icp iter %d:", iter); for (int k = 0; k < 30; ++k) std::fprintf(stderr, " %.6g", a[k]); std::fprintf(stderr, "\n"); }
    if (a[28] < 10) continue;  // "Insufficient matched surfs..." (:1175-1178)
    if (a[29] < 5) continue;   // "Insufficient matched corners..." (:1181-1184)
    linalg::Mat<6> JTJ;
    linalg::Vec<6> JTb;
    int k = 0;
    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { JTJ[i][j] = a[k]; JTJ[j][i] = a[k]; ++k; }
    for (int i = 0; i < 6; ++i) JTb[i] = a[21 + i];
    linalg::Vec<6> x = linalg::colPivQrSolve<6>(JTJ, JTb);
    if (ctx->verbose) std::fprintf(stderr, "[lins_gpu] icp iter %d x (before projection): %.12g %.12g %.12g %.12g %.12g %.12g\n", iter, x[0], x[1], x[2], x[3], x[4], x[5]);
    bool isDegenerate = false;
    if (iter == 0) {  // :1269-1296
      linalg::Vec<6> matE;
      linalg::Mat<6> matV, matV2, Vinv;
      linalg::symmetricEigen<6>(JTJ, matE, matV);
      if (ctx->verbose) std::fprintf(stderr, "[lins_gpu] icp E: %.9g %.9g %.9g %.9g %.9g %.9g\n", matE[0], matE[1], matE[2], matE[3], matE[4], matE[5]);
      matV2 = matV;
      for (int i = 0; i < 6; ++i) {
        if (matE[i] < 10.) { for (int j = 0; j < 6; ++j) matV2[i][j] = 0; isDegenerate = true; }
        else break;
      }
      if (!linalg::inverse<6>(matV, Vinv)) for (auto& row : Vinv) for (auto& e : row) e = std::numeric_limits<double>::quiet_NaN();
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double sacc = 0; for (int m = 0; m < 6; ++m) sacc += Vinv[i][m] * matV2[m][j]; matP[i][j] = sacc; }
    }
    if (ctx->verbose) std::fprintf(stderr, "[lins_gpu] icp iter %d degenerate %d\n", iter, (int)isDegenerate);
    if (isDegenerate) {
      linalg::Vec<6> x2 = x;
      for (int i = 0; i < 6; ++i) { double sacc = 0; for (int j = 0; j < 6; ++j) sacc += matP[i][j] * x2[j]; x[i] = sacc; }
    }
    Q4D dq = math_utils::rpy2Quat(V3D(x[0], x[1], x[2]));
    q = (q * dq).normalized();
    t = t + V3D(x[3], x[4], x[5]);
    double deltaR = V3D(math_utils::rad2deg(x[0]), math_utils::rad2deg(x[1]), math_utils::rad2deg(x[2])).norm();
    double deltaT = V3D(100 * x[3], 100 * x[4], 100 * x[5]).norm();
    if (deltaR < 0.1 && deltaT < 0.1) { conv = true; break; }
  }
  pose_io[0] = t.x(); pose_io[1] = t.y(); pose_io[2] = t.z();
  pose_io[3] = q.x(); pose_io[4] = q.y(); pose_io[5] = q.z(); pose_io[6] = q.w();
  if (iters_out) *iters_out = it;
  if (converged_out) *converged_out = conv ? 1 : 0;
  return LINS_OK;
}

// ≙ estimateTransform (StateEstimator.hpp:1163-1196): the association + J^T J / J^T b reduction of every
// Gauss-Newton step runs on device (MODE_ICP_REDUCE), and so do the 6x6 solve / degeneracy projection / pose update of
// calculateTransformation (:1260-1320): lins_icp_step.cuh.
int lins_gpu_estimate_transform(lins_ctx* ctx, const lins_point* surf_flat, int ns, const lins_point* corner_sharp,
                                int nc, double* pose_io, int* iters_out, int* converged_out) {
  if (!ctx) return LINS_E_INVALID;
  if (!pose_io) return fail(ctx, LINS_E_INVALID, "null pose");
  CK(cudaSetDevice(ctx->device));
  if (std::getenv("LINS_ICP_HOST_LOOP")) return estimate_transform_host_loop(ctx, surf_flat, ns, corner_sharp, nc, pose_io, iters_out, converged_out);
  // queries + initial pose are staged ONCE; every Gauss-Newton iteration is a reduction launch (MODE_ICP_REDUCE, linearised
  // at the pose block on the device) + the one-thread step kernel that updates that block; once the step kernel sets
  // `done` the remaining queued launches return immediately.  One D2H + one synchronisation at the end.
  double lin[19];
  std::memset(lin, 0, sizeof(lin));
  lin[0] = pose_io[0]; lin[1] = pose_io[1]; lin[2] = pose_io[2];
  lin[6] = pose_io[3]; lin[7] = pose_io[4]; lin[8] = pose_io[5]; lin[9] = pose_io[6];
  int rc = stage_single(ctx, surf_flat, ns, corner_sharp, nc, lin, nullptr, false);
  if (rc != LINS_OK) return rc;
  Resident& r = ctx->single;
  CK(r.icp.reserve(1)); CK(r.h_icp.reserve(1)); CK(r.h_state_out.reserve(20));
  CK(cudaMemsetAsync(r.icp.p, 0, sizeof(IcpState), ctx->stream));
  for (int iter = 0; iter < ctx->prm.num_iter; ++iter) {
    BatchView bv = single_view(ctx, false);
    bv.icp_done = &r.icp.p->done;
    rc = launch(ctx, r, bv, make_kparams(ctx->prm, MODE_ICP_REDUCE, iter));
    if (rc != LINS_OK) return rc;
    lins_icp_step_kernel<<<1, 32, 0, ctx->stream>>>(r.accum.p, r.state_in.p, r.icp.p, iter);
    CK(cudaGetLastError());
    ctx->launches += 1;
  }
  CK(cudaMemcpyAsync(r.h_state_out.p, r.state_in.p, sizeof(double) * 20, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(r.h_icp.p, r.icp.p, sizeof(IcpState), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const double* so = r.h_state_out.p;
  pose_io[0] = so[0]; pose_io[1] = so[1]; pose_io[2] = so[2];
  pose_io[3] = so[6]; pose_io[4] = so[7]; pose_io[5] = so[8]; pose_io[6] = so[9];
  if (iters_out) *iters_out = r.h_icp.p->iters;
  if (converged_out) *converged_out = r.h_icp.p->converged;
  return LINS_OK;
}

// ≙ updatePointCloud (StateEstimator.hpp:1116-1161), XYZ part: transformToEnd on the device, then the map swap + the
// guarded index refresh.  The clouds stay on the device as the next map; copying them back is optional (the reference
// transforms scan_new_'s clouds in place because the mapping node consumes them) and, when asked for, is one D2H of
// full PointXYZI records straight into the caller's clouds — one stream synchronisation per call, none without read-back.
int lins_gpu_update_map_ex(lins_ctx* ctx, const lins_point* surf, int ns, const lins_point* corner, int nc, const double* lin_state,
                           lins_point* surf_out, lins_point* corner_out, int* map_replaced) {
  if (!ctx) return LINS_E_INVALID;
  if (ns < 0 || nc < 0 || (ns > 0 && !surf) || (nc > 0 && !corner)) return fail(ctx, LINS_E_INVALID, "bad update_map args");
  CK(cudaSetDevice(ctx->device));
  const double* lin_dev = nullptr;
  if (!lin_state) {  // the posterior the last lins_gpu_ieskf left on the device (linState_ of a run that did not diverge)
    if (ctx->single.n != 1 || !ctx->single.state_out.p) return fail(ctx, LINS_E_INVALID, "no device-resident state: call lins_gpu_ieskf first or pass lin_state");
    lin_dev = ctx->single.state_out.p;
  }
  const bool rebuild = nc >= 5 && ns >= 20;  // :1156-1157
  const bool want_out = (ns > 0 && surf_out) || (nc > 0 && corner_out);
  // every allocation first: the context is only changed once nothing can fail any more
  const bool keep_tree = !rebuild && ctx->tree_is_map;  // keep the old map alive as the 1-NN cloud if the guard fails
  DevBuf<float4>& dst_s = keep_tree ? ctx->tree_s : ctx->map_s;  // (after the swap below these are the new map buffers)
  DevBuf<float4>& dst_c = keep_tree ? ctx->tree_c : ctx->map_c;
  CK(dst_s.reserve(ns + 1)); CK(dst_c.reserve(nc + 1));
  CK(ctx->map_s.reserve(1)); CK(ctx->map_c.reserve(1)); CK(ctx->tree_s.reserve(1)); CK(ctx->tree_c.reserve(1));
  CK(tmp_staging_wait(ctx));  // (the H2D of a previous call may still be reading the pinned staging)
  CK(ctx->h_tmp.reserve((size_t)ns + nc + 1));
  CK(ctx->tmp_lin.reserve(20));
  if (want_out) CK(ctx->tmp_out.reserve((size_t)ns + nc + 1));
  if (keep_tree) {
    std::swap(ctx->tree_s, ctx->map_s); std::swap(ctx->tree_c, ctx->map_c);
    ctx->tree_ns = std::max(ctx->map_ns, 0); ctx->tree_nc = std::max(ctx->map_nc, 0);
    ctx->tree_is_map = false;
  }
  pack_into(ctx->h_tmp.p, surf, ns);
  pack_into(ctx->h_tmp.p + ns, corner, nc);
  if (lin_state) {
    double lin[20];
    std::memcpy(lin, lin_state, sizeof(double) * 19);
    lin[19] = 0;
    CK(cudaMemcpyAsync(ctx->tmp_lin.p, lin, sizeof(lin), cudaMemcpyHostToDevice, ctx->stream));  // (pageable source: staged before the call returns)
    lin_dev = ctx->tmp_lin.p;
  }
  if (ns) CK(cudaMemcpyAsync(ctx->map_s.p, ctx->h_tmp.p, sizeof(float4) * ns, cudaMemcpyHostToDevice, ctx->stream));
  if (nc) CK(cudaMemcpyAsync(ctx->map_c.p, ctx->h_tmp.p + ns, sizeof(float4) * nc, cudaMemcpyHostToDevice, ctx->stream));
  CK(tmp_staging_mark(ctx));
  lins_point* o_s = surf_out && ns ? ctx->tmp_out.p : nullptr;
  lins_point* o_c = corner_out && nc ? ctx->tmp_out.p + ns : nullptr;
  if (ns) { lins_transform_to_end_kernel<<<(ns + 255) / 256, 256, 0, ctx->stream>>>(ctx->map_s.p, ns, lin_dev, ctx->prm.scan_period, o_s); ctx->launches += 1; }
  if (nc) { lins_transform_to_end_kernel<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(ctx->map_c.p, nc, lin_dev, ctx->prm.scan_period, o_c); ctx->launches += 1; }
  CK(cudaGetLastError());
  ctx->map_ns = ns; ctx->map_nc = nc;
  if (rebuild) { ctx->tree_is_map = true; ctx->tree_ns = ns; ctx->tree_nc = nc; }
  if (map_replaced) *map_replaced = rebuild ? 1 : 0;
  int rc = upload_map_offsets(ctx);
  if (rc != LINS_OK) return rc;
  if (want_out) {  // full records through pinned staging (a D2H into pageable memory is staged chunk by chunk by the driver)
    CK(ctx->h_out.reserve((size_t)ns + nc + 1));
    if (o_s) CK(cudaMemcpyAsync(ctx->h_out.p, o_s, sizeof(lins_point) * ns, cudaMemcpyDeviceToHost, ctx->stream));
    if (o_c) CK(cudaMemcpyAsync(ctx->h_out.p + ns, o_c, sizeof(lins_point) * nc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (o_s) std::memcpy(surf_out, ctx->h_out.p, sizeof(lins_point) * ns);
    if (o_c) std::memcpy(corner_out, ctx->h_out.p + ns, sizeof(lins_point) * nc);
  }
  return LINS_OK;
}

int lins_gpu_update_map(lins_ctx* ctx, lins_point* surf, int ns, lins_point* corner, int nc, const double* lin_state,
                        int* map_replaced) {
  if (!ctx) return LINS_E_INVALID;
  if (!lin_state) return fail(ctx, LINS_E_INVALID, "bad update_map args");
  return lins_gpu_update_map_ex(ctx, surf, ns, corner, nc, lin_state, surf, corner, map_replaced);
}

}  // extern "C"
