// One association + reduction pass over all queries of a scan at the current linearisation point
// (rows A2-A9 of SURVEY.md §8; reference findCorrespondingSurfFeatures / findCorrespondingCornerFeatures,
// lins/include/StateEstimator.hpp:829-1063, + the measurement assembly :499-532 folded into 28 sums).
//
// Fast path (sm.az_ok): targets are ring-sorted, the (ring, azimuth) index of lins_assoc_az.cuh is valid.
// Legacy path: any ring order / ring values / a 1-NN cloud that differs from the walk cloud (the stale-index
// quirk of :1156-1160): brute-force exact 1-NN + the plain (or literal sequential) walks over global memory.
#pragma once
#include "lins_assoc_az.cuh"

namespace lins_dev {

// Probe-first search (lins_assoc_az.cuh: az_probe_window) is used for the closest-point search of a scan's FIRST pass
// only.  Measured on the VLP-16 workload: there every query is unseeded, all warps are busy and the phase is issue
// bound (176 K -> 104 K cycles per scan); in later passes and for the walks a search is a lone warp bound by latency,
// which a second scan only lengthens (tried, slower).
#ifndef LINS_SEARCH_DIAG
#define LINS_SEARCH_DIAG 0
#endif
constexpr bool kSearchDiag = LINS_SEARCH_DIAG != 0;  // per-search cycle counters for tools/phase_profile.py

struct PassBuffers {
  float4* qpt;               // staged queries (x, y, z, intensity) of the current tile
  float4* sel;               // de-skewed queries (pointSel)
  unsigned long long* key;   // legacy path: 1-NN keys
  int* pos;                  // 3 per query.  fast path: slots in the sorted copies (closest, Ind2, Ind3);
                             //               legacy path: original indices
  float4* qa;                // fast path: (azimuth, rho, -, -) of the de-skewed query
  int4* qw;                  // fast path: search windows w1 / w2 / w3 and (ring << 24 | index) of the closest point
  float4* qref;              // fast path: pointSel at the last closest-point search + the slack (m) of its answer
  float4* qref2;             // fast path: pointSel at the last walk search + the slack of Ind2 / Ind3
  int* qccr;                 // fast path: (ring << 24 | original index) of the closest point, -1 = none
  float4* qext;              // fast path: (bound / slack of the Ind3 search, runner-up slots of closest / Ind2 / Ind3 as int bits)
  int* wl;                   // fast path: work list of the queries that need a closest-point search / ring walks this pass
  const float4* azS;         // sorted copies (shared or global)
  const float4* azC;
};

template <int MODE>
__device__ void association_pass(Smem& sm, const BatchView& bv, const KParams& kp, int scan, int iter, const PassBuffers& pb,
                                 bool first_pass_of_scan) {
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
  const bool search = (iter % kp.icp_freq) == 0;
  const bool weighted = iter >= kp.icp_freq;
  const float nearf = (float)kp.nearest_sq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool fast = sm.az_ok != 0;
  const int fwdS = min(ns, Ts), fwdC = min(nc, Tc);  // :859 / :983 loop-bound quirk (+ OOB clamp)
  const int ntot = ns + nc;
  // previous answers are usable as bounds only if this scan has a single query tile (slots stay in shared memory)
  const bool seeded = fast && !first_pass_of_scan && ntot <= bv.qtile;

  AzIndex ixS, ixC;
  ixS.sbase = Ts <= bv.cap_s ? smem_u32(pb.azS) : 0u;
  ixC.sbase = Tc <= bv.cap_c ? smem_u32(pb.azC) : 0u;
  ixS.pts = pb.azS; ixS.bstart = sm.azTabS; ixS.nb = sm.nbS; ixS.nrings = sm.nringsS; ixS.T = Ts;
  ixC.pts = pb.azC; ixC.bstart = sm.azTabC; ixC.nb = sm.nbC; ixC.nrings = sm.nringsC; ixC.T = Tc;

  if (threadIdx.x < kNAcc) sm.acc[threadIdx.x] = 0.0;
  if (threadIdx.x == 32) { sm.cnt[0] = 0; sm.cnt[1] = 0; }
  // (ordered before the first block_reduce_acc by the barriers of the tile loop)

  for (int q0 = 0; q0 < ntot; q0 += bv.qtile) {
    const int nq = min(bv.qtile, ntot - q0);
    const int nsT = max(0, min(ns - q0, nq));  // sub-lists of this tile: surf [0, nsT) then corner [nsT, nq)
    // ---- stage the tile's queries (1-D TMA) --------------------------------------------------------------------
    __syncthreads();  // previous tile fully consumed
    if (ntot > bv.qtile || first_pass_of_scan) {  // a single-tile scan keeps its queries staged across iterations
      if (threadIdx.x == 0) {
        uint32_t bytes = 0;
        if (nsT > 0) bytes += (uint32_t)nsT * 16u;
        if (nq - nsT > 0) bytes += (uint32_t)(nq - nsT) * 16u;
        fence_proxy_async();
        mbar_expect_tx(&sm.mbar, bytes);
        if (nsT > 0) tma_load_1d(pb.qpt, bv.qs + qs0 + q0, (uint32_t)nsT * 16u, &sm.mbar);
        if (nq - nsT > 0) tma_load_1d(pb.qpt + nsT, bv.qc + qc0 + max(0, q0 - ns), (uint32_t)(nq - nsT) * 16u, &sm.mbar);
      }
      const unsigned int ph = sm.phase;  // one mbarrier phase per staged tile
      mbar_wait(&sm.mbar, ph & 1u);
      __syncthreads();
      if (threadIdx.x == 0) sm.phase = ph + 1u;
    }
    LINS_TICK(2);
    // ---- A2: de-skew ---------------------------------------------------------------------------------------------
    // ... fused with phase P1 of the association (same thread-per-query mapping, P1 touches only its own query's
    // state and the read-only index): per-query level -2 everything certified, -3 closest point certified (walks
    // only), >= 0 full search (= window), -1 cannot match.  The work-list counters were reset after the previous pass.
    const bool p1_here = search && fast;
    for (int i = threadIdx.x; i < nq; i += kThreads) {
      const float4 s = transform_to_start(pb.qpt[i], sm, kp.scan_period);
      pb.sel[i] = s;
      pb.key[i] = kKeyMax;
      if (!p1_here) continue;
      const bool surf = i < nsT;
      float4 qa = make_float4(0.f, 0.f, -1.f, 0.f);
      if (seeded) {  // certificates (lins_assoc_az.cuh: cert_accepted / cert_rejected): the stored answers still hold
        const float4 r1 = pb.qref[i], r2 = pb.qref2[i], ex = pb.qext[i];
        const AzIndex& ixq = surf ? ixS : ixC;
        const unsigned nearbits = __float_as_uint(nearf);
        const float moved1 = sqrtf(sqdist_f32(s.x, s.y, s.z, r1.x, r1.y, r1.z));
        const int w1s = pb.pos[3 * i];
        const bool ok1 = w1s >= 0 ? cert_accepted<false>(ixq, s, w1s, __float_as_int(ex.y), r1.w, moved1, nearbits, 0) : cert_rejected(r1.w, moved1);
        bool ok2 = false;
        const int ccr0 = pb.qccr[i];
        if (ok1 && ccr0 >= 0) {  // (the walks' candidate sets are defined by the closest point: only meaningful while it stands)
          const int c0 = ccr0 & 0x00ffffff;
          const float moved2 = sqrtf(sqdist_f32(s.x, s.y, s.z, r2.x, r2.y, r2.z));
          const int w2s = pb.pos[3 * i + 1], w3s = pb.pos[3 * i + 2];
          ok2 = w2s >= 0 ? cert_accepted<true>(ixq, s, w2s, __float_as_int(ex.z), r2.w, moved2, nearbits, c0) : cert_rejected(r2.w, moved2);
          if (ok2 && surf) ok2 = w3s >= 0 ? cert_accepted<true>(ixq, s, w3s, __float_as_int(ex.w), ex.x, moved2, nearbits, c0) : cert_rejected(ex.x, moved2);
        }
        if (ok1 && (ok2 || ccr0 < 0)) { pb.qw[i] = make_int4(-2, 0, 0, 0); continue; }
        if (ok1) { az_polar(s, qa); pb.qa[i] = qa; pb.qw[i] = make_int4(-3, 0, 0, 0); continue; }
      }
      const int w1 = az_prepare_nn(surf ? ixS : ixC, s, nearf, seeded ? pb.pos[3 * i] : -1, qa);
      pb.qa[i] = qa;
      pb.qw[i] = make_int4(w1, 0, 0, 0);
      pb.wl[atomicAdd(&sm.wl_n[0], 1)] = i;  // (list order does not matter: every query's result goes to its own slot)
      if (bv.timers && w1 >= 0) atomicAdd(&sm.dbg[0], w1 & 0xffff);
    }
    __syncthreads();
    LINS_TICK(3);
    if (search && fast) {
      // ---- A3/A4 fast path.  Scalar preparation (certificates, atan2f, asinf, bounds: P1 above, P3 below) runs one
      // THREAD per query so that all queries proceed in parallel; the memory scans (P2, P4) run one WARP per query.
      if (bv.timers && threadIdx.x == 0) {
        atomicAdd((unsigned long long*)&bv.timers[10], (unsigned long long)sm.wl_n[0]);
        atomicAdd((unsigned long long*)&bv.timers[12], (unsigned long long)sm.dbg[0]);
        if (first_pass_of_scan) atomicAdd((unsigned long long*)&bv.timers[14], (unsigned long long)sm.dbg[0]);
      }
      const float gate = sqrtf(nearf);
      // P2: warps pull queries from the work list (the per-query cost is heavy-tailed; a static split leaves warps idle)
      for (;;) {
        const bool diag = kSearchDiag && bv.timers != nullptr;  // per-search clocks: compiled out unless -DLINS_SEARCH_DIAG=1
        const long long t_f0 = diag ? clock64() : 0;
        long long tmv[4] = {0, 0, 0, 0};
        long long* tm = diag ? tmv : nullptr;
        int k = 0;
        if (lane == 0) k = atomicAdd(&sm.wl_head[0], 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= sm.wl_n[0]) break;
        const int i = pb.wl[k];
        int w1 = pb.qw[i].x;
        const long long t_s0 = (diag && lane == 0) ? clock64() : 0;
        unsigned long long k1 = kKeyMax;
        int p1 = -1;
        Top3 top;
        top.init();
        const bool surf = i < nsT;
        const AzIndex& ix = surf ? ixS : ixC;
        const float4 s = pb.sel[i];
        float4 qa = pb.qa[i];
        if (w1 >= 0) {
          const int wp = first_pass_of_scan ? az_probe_window(ix, qa, w1) : -1;
          if (wp >= 0) {  // wide window (no usable previous answer): probe first, then search inside the implied window
            const Top3 pr = az_scan_nn(ix, s, wp);
            if (pr.p1 >= 0) {
              w1 = az_nn_window(ix, az_seed_bound(ix, s, pr.p1, nearf), qa);
              if (lane == 0) pb.qa[i] = qa;
            }
          }
          top = az_scan_nn(ix, s, w1, tm);
          k1 = top.k1; p1 = top.p1;
        }
        if (lane == 0) {
          const float d1 = __uint_as_float((unsigned)(k1 >> 32));
          const bool acc1 = k1 != kKeyMax && p1 >= 0 && (double)d1 < kp.nearest_sq;
          // accepted: what everything but the two front-runners exceeded; else the slack of "nothing within the gate"
          const float bound1 = w1 < 0 ? -1.f : acc1 ? cert_bound(top.d3, qa.w) : rejected_slack((unsigned)(k1 >> 32), qa.w, gate);
          pb.qref[i] = make_float4(s.x, s.y, s.z, bound1);
          pb.qext[i].y = __int_as_float(acc1 ? top.p2 : -1);
          pb.pos[3 * i] = acc1 ? p1 : -1;
          pb.qccr[i] = acc1 ? ((slot_ring(ix.pts[p1].w) << 24) | (int)(unsigned)(k1 & 0xffffffffu)) : -1;
          if (diag && !first_pass_of_scan) {
            const unsigned long long dt = (unsigned long long)(clock64() - t_s0);
            atomicAdd((unsigned long long*)&bv.timers[20], dt);
            atomicMax((unsigned long long*)&bv.timers[22], dt);
            const long long t_e = clock64();
            if (w1 >= 0) {  // fetch+loads | setup+loop | arg-min | epilogue
              atomicAdd((unsigned long long*)&bv.timers[32], (unsigned long long)(t_s0 - t_f0));
              atomicAdd((unsigned long long*)&bv.timers[33], (unsigned long long)(tmv[0] - t_s0));
              atomicAdd((unsigned long long*)&bv.timers[34], (unsigned long long)(tmv[1] - tmv[0]));
              atomicAdd((unsigned long long*)&bv.timers[35], (unsigned long long)(t_e - tmv[1]));
              atomicAdd((unsigned long long*)&bv.timers[36], 1ull);
              atomicAdd((unsigned long long*)&bv.timers[37], (unsigned long long)tmv[2]);
              atomicAdd((unsigned long long*)&bv.timers[38], (unsigned long long)tmv[3]);
            }
            if (w1 >= 0 && (w1 & 0xffff) >= 64) { atomicAdd((unsigned long long*)&bv.timers[16], dt); atomicAdd((unsigned long long*)&bv.timers[17], 1ull); }
          }
        }
      }
      __syncthreads();
      LINS_TICK_F(4, 28, first_pass_of_scan);
      for (int i = threadIdx.x; i < nq; i += kThreads) {  // P3
        const int lvl = pb.qw[i].x;
        if (lvl == -2) continue;
        const bool surf = i < nsT;
        const int gq = q0 + i;
        const int ccr = pb.qccr[i];
        if (ccr < 0) {  // no closest point within the gate: nothing to walk
          pb.pos[3 * i + 1] = -1; pb.pos[3 * i + 2] = -1;
          pb.qw[i] = make_int4(-1, 0, 0, 0);
          if (surf) { int* o = bv.ind_s + 3 * (size_t)(qs0 + gq); o[0] = -1; o[1] = -1; o[2] = -1; }
          else { int* o = bv.ind_c + 2 * (size_t)(qc0 + gq - ns); o[0] = -1; o[1] = -1; }
          continue;
        }
        int w2 = 0, w3 = 0;
        float B2 = 0.f, B3 = 0.f;
        const int sd2 = seeded ? pb.pos[3 * i + 1] : -1, sd3 = seeded ? pb.pos[3 * i + 2] : -1;
        const int c = ccr & 0x00ffffff, cr = (int)((unsigned)ccr >> 24);
        if (surf) az_prepare_walk<true>(ixS, pb.sel[i], pb.qa[i], c, cr, sd2, sd3, fwdS, nearf, w2, w3, B2, B3);
        else az_prepare_walk<false>(ixC, pb.sel[i], pb.qa[i], c, cr, sd2, sd3, fwdC, nearf, w2, w3, B2, B3);
        pb.qw[i] = make_int4(1, w2, w3, ccr);
        reinterpret_cast<float2*>(pb.key)[i] = make_float2(B2, B3);
        pb.wl[atomicAdd(&sm.wl_n[1], 1)] = i;
        if (bv.timers) atomicAdd(&sm.dbg[1], (w2 & 0xffff) + 4 * (w3 & 0xffff));
      }
      __syncthreads();
      if (bv.timers && threadIdx.x == 0) {
        atomicAdd((unsigned long long*)&bv.timers[11], (unsigned long long)sm.wl_n[1]);
        atomicAdd((unsigned long long*)&bv.timers[13], (unsigned long long)sm.dbg[1]);
        if (first_pass_of_scan) atomicAdd((unsigned long long*)&bv.timers[15], (unsigned long long)sm.dbg[1]);
      }
      for (;;) {  // P4: same work-list scheme
        int k = 0;
        if (lane == 0) k = atomicAdd(&sm.wl_head[1], 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= sm.wl_n[1]) break;
        const int i = pb.wl[k];
        const int4 w = pb.qw[i];
        const bool diag = kSearchDiag && bv.timers != nullptr;
        const long long t_s0 = (diag && lane == 0) ? clock64() : 0;
        const bool surf = i < nsT;
        const int gq = q0 + i;
        const float2 B = reinterpret_cast<const float2*>(pb.key)[i];
        const int w2 = w.y, w3 = w.z;
        const float4 s = pb.sel[i];
        const WalkOut wo = surf ? az_scan_walk<true>(ixS, s, w.w, w2, w3, fwdS, nearf, B.x, B.y) : az_scan_walk<false>(ixC, s, w.w, w2, w3, fwdC, nearf, B.x, B.y);
        const int i2 = wo.i2, i3 = wo.i3;
        if (lane == 0) {
          pb.pos[3 * i + 1] = wo.pos2; pb.pos[3 * i + 2] = wo.pos3;
          pb.qref2[i] = make_float4(s.x, s.y, s.z, wo.bound2);
          float4 ex = pb.qext[i];  // (.y = the closest point's runner-up, written by P2 or kept from an earlier pass)
          ex.x = wo.bound3; ex.z = __int_as_float(wo.run2); ex.w = __int_as_float(wo.run3);
          pb.qext[i] = ex;
          if (diag && !first_pass_of_scan) {
            const unsigned long long dt = (unsigned long long)(clock64() - t_s0);
            atomicAdd((unsigned long long*)&bv.timers[30], dt);
            if (max(w2 & 0xffff, w3 & 0xffff) >= 64) atomicAdd((unsigned long long*)&bv.timers[31], dt);
          }
          const int i1 = w.w & 0x00ffffff;
          if (surf) { int* o = bv.ind_s + 3 * (size_t)(qs0 + gq); o[0] = i1; o[1] = i2; o[2] = i3; }
          else { int* o = bv.ind_c + 2 * (size_t)(qc0 + gq - ns); o[0] = i1; o[1] = i2; }
        }
      }
    } else if (search) {
      // ---- legacy: brute-force exact 1-NN ----------------------------------------------------------------------------
      if (nsT > 0 && TnS > 0) nn_brute(pb.sel, pb.key, nsT, nnS, TnS);
      if (nq - nsT > 0 && TnC > 0) nn_brute(pb.sel + nsT, pb.key + nsT, nq - nsT, nnC, TnC);
      __syncthreads();
      LINS_TICK(4);
      for (int i = warp; i < nq; i += kWarps) {
        const bool surf = i < nsT;
        const unsigned long long k1 = pb.key[i];
        const float d1 = __uint_as_float((unsigned)(k1 >> 32));
        const int c = (int)(unsigned)(k1 & 0xffffffffu);
        const bool found = (k1 != kKeyMax) && ((double)d1 < kp.nearest_sq) && c < (surf ? Ts : Tc);
        int i1 = -1, i2 = -1, i3 = -1;
        if (found) {
          i1 = c;
          const float4 s = pb.sel[i];
          if (surf) {
            if (sm.sortedS) walk_warp<true>(s, c, tgtS, Ts, sm.rsS, fwdS, nearf, i2, i3);
            else { if (lane == 0) walk_seq<true>(s, c, tgtS, Ts, fwdS, nearf, i2, i3); }
          } else {
            if (sm.sortedC) walk_warp<false>(s, c, tgtC, Tc, sm.rsC, fwdC, nearf, i2, i3);
            else { if (lane == 0) walk_seq<false>(s, c, tgtC, Tc, fwdC, nearf, i2, i3); }
          }
        }
        if (lane == 0) {
          pb.pos[3 * i] = i1; pb.pos[3 * i + 1] = i2; pb.pos[3 * i + 2] = i3;
          const int gq = q0 + i;
          if (surf) { int* o = bv.ind_s + 3 * (size_t)(qs0 + gq); o[0] = i1; o[1] = i2; o[2] = i3; }
          else { int* o = bv.ind_c + 2 * (size_t)(qc0 + gq - ns); o[0] = i1; o[1] = i2; }
        }
      }
    } else if (!fast || ntot > bv.qtile) {
      // iter % ICP_FREQ != 0: reuse pointSearch*Ind (StateEstimator.hpp:844, :970) — original indices from global
      for (int i = threadIdx.x; i < nq; i += kThreads) {
        const int gq = q0 + i;
        if (i < nsT) { const int* o = bv.ind_s + 3 * (size_t)(qs0 + gq); pb.pos[3 * i] = o[0]; pb.pos[3 * i + 1] = o[1]; pb.pos[3 * i + 2] = o[2]; }
        else { const int* o = bv.ind_c + 2 * (size_t)(qc0 + gq - ns); pb.pos[3 * i] = o[0]; pb.pos[3 * i + 1] = o[1]; pb.pos[3 * i + 2] = -1; }
      }
    }
    __syncthreads();
    LINS_TICK_F(5, 29, first_pass_of_scan);
    // the searches of this pass are over: reset the work lists for the next pass (whose P1 is fused with the de-skew)
    if (threadIdx.x == kThreads - 1) { sm.wl_n[0] = 0; sm.wl_n[1] = 0; sm.wl_head[0] = 0; sm.wl_head[1] = 0; sm.dbg[0] = 0; sm.dbg[1] = 0; }
    // ---- A5/A6 residuals + A7-A9 fold ------------------------------------------------------------------------------
    // tripod points: fast path -> slots of the sorted copies; otherwise original indices into the walk clouds.
    // (fast path with several tiles on a non-search iteration falls back to original indices, see above)
    const bool by_slot = fast && (search || ntot <= bv.qtile);
    // the 28 f64 accumulators live only here, so they do not take registers away from the search phases
    double acc[kNAcc];
#pragma unroll
    for (int k = 0; k < kNAcc; ++k) acc[k] = 0.0;
    int cntS = 0, cntC = 0;
    for (int i = threadIdx.x; i < nq; i += kThreads) {
      const bool surf = i < nsT;
      const int i1 = pb.pos[3 * i], i2 = pb.pos[3 * i + 1], i3 = pb.pos[3 * i + 2];
      const float4 s = pb.sel[i];
      float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
      bool ok = false;
      if (surf) {
        if (i2 >= 0 && i3 >= 0 && i1 >= 0 && i1 < Ts && i2 < Ts && i3 < Ts) {
          if (by_slot) ok = plane_residual(s, pb.azS[i1], pb.azS[i2], pb.azS[i3], weighted, coeff);
          else ok = plane_residual(s, __ldg(&tgtS[i1]), __ldg(&tgtS[i2]), __ldg(&tgtS[i3]), weighted, coeff);
        }
      } else {
        if (i2 >= 0 && i1 >= 0 && i1 < Tc && i2 < Tc) {
          if (by_slot) ok = line_residual(s, pb.azC[i1], pb.azC[i2], weighted, coeff);
          else ok = line_residual(s, __ldg(&tgtC[i1]), __ldg(&tgtC[i2]), weighted, coeff);
        }
      }
      if (ok) {
        if (MODE == MODE_ICP_REDUCE) accumulate_row_icp(pb.qpt[i], coeff, sm, kp.scan_period, acc);
        else accumulate_row(pb.qpt[i], coeff, sm.R, kp.lidar_scale, acc);
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
    LINS_TICK(6);
    // the fused update finishes the last tile's sum inside warp 0's serial tail (one barrier less)
    block_reduce_acc(sm, acc, cntS, cntC, !(MODE == MODE_IESKF && q0 + bv.qtile >= ntot));
  }
  if (ntot == 0) __syncthreads();  // no tile ran: still order the zeroed sums before the caller reads them
  LINS_TICK(7);
}

}  // namespace lins_dev
