// One association + reduction pass over the queries of EVERY resident unit of the CTA at each unit's current
// linearisation point (rows A2-A9 of SURVEY.md §8; reference findCorrespondingSurfFeatures /
// findCorrespondingCornerFeatures, lins/include/StateEstimator.hpp:829-1063, + the measurement assembly :499-532
// folded into 28 sums per unit).
//
// Queries live in a virtual index space v = slot * qtile + i (qtile a multiple of 32, so a warp never straddles two
// units).  Thread-per-query phases sweep v, warp-per-search phases pull v from one work list shared by all slots.
//
// Fast path (sm.az_ok): targets are ring-sorted, the (ring, azimuth) index of lins_assoc_az.cuh is valid.
// Legacy path: any ring order / ring values / a 1-NN cloud that differs from the walk cloud (the stale-index
// quirk of :1156-1160): brute-force exact 1-NN + the literal sequential walks over global memory.
#pragma once
#include "lins_assoc_az.cuh"

namespace lins_dev {

// Probe-first search (lins_assoc_az.cuh: az_probe_window) is used for the closest-point search of a unit's FIRST pass
// only: there every query is unseeded and the gate-wide window is much larger than the neighbourhood that holds the
// answer.
#ifndef LINS_SEARCH_DIAG
#define LINS_SEARCH_DIAG 0
#endif
constexpr bool kSearchDiag = LINS_SEARCH_DIAG != 0;  // per-search cycle counters for tools/phase_profile.py
// widest windows (azimuth bins per ring) a group of kGroupLanes lanes scans; wider ones go to a whole warp
#ifndef LINS_GROUP_LANES
#define LINS_GROUP_LANES 4
#endif
constexpr int kGroupLanes = LINS_GROUP_LANES;
#ifndef LINS_THREAD_SCAN_BINS
#define LINS_THREAD_SCAN_BINS 16
#endif
#ifndef LINS_THREAD_WALK_BINS
#define LINS_THREAD_WALK_BINS 48
#endif
constexpr int kThreadScanBins = LINS_THREAD_SCAN_BINS, kThreadWalkBins = LINS_THREAD_WALK_BINS;

struct PassBuffers {           // per-query arrays, indexed by v = slot * qtile + i (shared memory, or the CTA's global scratch)
  float4* qpt;               // staged queries (x, y, z, intensity)
  float4* sel;               // de-skewed queries (pointSel)
  unsigned long long* key;   // legacy path: 1-NN keys; fast path: (B2, B3) of the walk windows
  int* pos;                  // 3 per query.  fast path: slots in the sorted copies (closest, Ind2, Ind3);
                             //               legacy path / reloaded correspondences: original indices
  float4* qa;                // fast path: (azimuth, rho, -, bound of everything outside the window) of the de-skewed query
  int4* qw;                  // fast path: level / search windows w1 / w2 / w3 and (ring << 24 | index) of the closest point
  float4* qref;              // fast path: pointSel at the last closest-point search + the certificate bound of its answer
  float4* qref2;             // fast path: pointSel at the last walk search + the bound of Ind2
  int* qccr;                 // fast path: (ring << 24 | original index) of the closest point, -1 = none
  float4* qext;              // fast path: (bound of Ind3, runner-up slots of closest / Ind2 / Ind3 as int bits)
  int* wl;                   // work list of the queries that need a closest-point search / ring walks this pass
  double* wacc;              // [slot][warp of the slot][28] partial folds of the pass
  int* wcnt;                 // [slot][warp of the slot][2] accepted surf / corner measurements
  int nvw;                   // warps per slot = qtile / 32
};

__device__ __forceinline__ AzIndex az_index_of(const Smem& sm, const BatchView& bv, bool surf) {
  AzIndex ix;
  if (surf) { ix.pts = bv.az_s + sm.ts0; ix.bstart = sm.azTabS; ix.elev = sm.elevS; ix.nb = sm.nbS; ix.nrings = sm.nringsS; ix.T = sm.Ts; }
  else { ix.pts = bv.az_c + sm.tc0; ix.bstart = sm.azTabC; ix.elev = sm.elevC; ix.nb = sm.nbC; ix.nrings = sm.nringsC; ix.T = sm.Tc; }
  return ix;
}

// slot of a virtual query index (nslots <= kMaxSlots = 4)
__device__ __forceinline__ int slot_of(int v, int Q) { return (v >= Q) + (v >= 2 * Q) + (v >= 3 * Q); }

template <int MODE>
__device__ void association_pass(CtaMem& cta, Smem* slots, const BatchView& bv, const KParams& kp, const PassBuffers& pb) {
  const int Q = bv.qtile, NQ = bv.nslots * Q;
  const float nearf = (float)kp.nearest_sq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int toff = cta.pass_first ? 32 : 0;  // diagnostics: passes that contain a unit's first pass are clocked separately

  // ---- A2: de-skew, fused with phase P1 of the association (same thread-per-query mapping; P1 touches only its own
  // query's state and the read-only index): per-query level -2 everything certified, -3 closest point certified (walks
  // only), >= 0 full search (= window), -1 cannot match.  The work-list counters were reset after the previous pass.
  for (int v = threadIdx.x; v < NQ; v += kThreads) {
    const int sl = slot_of(v, Q), i = v - sl * Q;
    Smem& sm = slots[sl];
    if (!sm.run || i >= sm.ns + sm.nc) continue;
    const float4 s = transform_to_start(pb.qpt[v], sm, kp.scan_period);
    pb.sel[v] = s;
    pb.key[v] = kKeyMax;
    const bool search = (sm.iter % kp.icp_freq) == 0;
    if (!search) {
      // iter % ICP_FREQ != 0: reuse pointSearch*Ind (StateEstimator.hpp:844, :970); a unit that has not searched in this
      // launch takes them from global memory (they persist between calls like the reference's arrays)
      if (!sm.pos_valid) {
        if (i < sm.ns) { const int* o = bv.ind_s + 3 * (size_t)(sm.qs0 + i); pb.pos[3 * v] = o[0]; pb.pos[3 * v + 1] = o[1]; pb.pos[3 * v + 2] = o[2]; }
        else { const int* o = bv.ind_c + 2 * (size_t)(sm.qc0 + i - sm.ns); pb.pos[3 * v] = o[0]; pb.pos[3 * v + 1] = o[1]; pb.pos[3 * v + 2] = -1; }
      }
      continue;
    }
    if (!sm.az_ok) continue;
    const bool surf = i < sm.ns;
    const bool seeded = !sm.first_pass;
    const AzIndex ixq = az_index_of(sm, bv, surf);
    float4 qa = make_float4(0.f, 0.f, -1.f, 0.f);
    if (seeded) {  // certificates (lins_assoc_az.cuh: cert_accepted / cert_rejected): the stored answers still hold
      const float4 r1 = pb.qref[v], r2 = pb.qref2[v], ex = pb.qext[v];
      const unsigned nearbits = __float_as_uint(nearf);
      const int w1s = pb.pos[3 * v], w2s = pb.pos[3 * v + 1], w3s = pb.pos[3 * v + 2];
      const int r1s = __float_as_int(ex.y), r2s = __float_as_int(ex.z), r3s = __float_as_int(ex.w);
      // every front-runner of the query in one batch of independent (predicated) loads; a slot is read only where the
      // sequential evaluation below could use it (the runner-up fields of a rejected search and the walk fields of a query
      // without a closest point are not maintained)
      const int ccr0 = pb.qccr[v];
      const bool walks = ccr0 >= 0;
      auto entry = [&](int slot, bool use) -> float4 {
        return use && slot >= 0 && slot < ixq.T ? ixq.pts[slot] : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      const float4 tw1 = entry(w1s, true), tr1 = entry(r1s, w1s >= 0);
      const float4 tw2 = entry(w2s, walks), tr2 = entry(r2s, walks && w2s >= 0);
      const float4 tw3 = entry(w3s, walks && surf), tr3 = entry(r3s, walks && surf && w3s >= 0);
      const float moved1 = sqrtf(sqdist_f32(s.x, s.y, s.z, r1.x, r1.y, r1.z));
      const bool ok1 = w1s >= 0 ? cert_accepted<false>(tw1, tr1, s, r1s, r1.w, moved1, nearbits, 0) : cert_rejected(r1.w, moved1);
      bool ok2 = false;
      if (ok1 && walks) {  // (the walks' candidate sets are defined by the closest point: only meaningful while it stands)
        const int c0 = ccr0 & 0x00ffffff;
        const float moved2 = sqrtf(sqdist_f32(s.x, s.y, s.z, r2.x, r2.y, r2.z));
        ok2 = w2s >= 0 ? cert_accepted<true>(tw2, tr2, s, r2s, r2.w, moved2, nearbits, c0) : cert_rejected(r2.w, moved2);
        if (ok2 && surf) ok2 = w3s >= 0 ? cert_accepted<true>(tw3, tr3, s, r3s, ex.x, moved2, nearbits, c0) : cert_rejected(ex.x, moved2);
      }
      if (ok1 && (ok2 || ccr0 < 0)) { pb.qw[v] = make_int4(-2, 0, 0, 0); continue; }
      if (ok1) { az_polar(s, qa); pb.qa[v] = qa; pb.qw[v] = make_int4(-3, 0, 0, 0); continue; }
    }
    const int w1 = az_prepare_nn(ixq, s, nearf, seeded ? pb.pos[3 * v] : -1, (int)pb.qpt[v].w, qa);
    pb.qa[v] = qa;
    pb.qw[v] = make_int4(w1, 0, 0, 0);
    // (list order does not matter: every query's result goes to its own slot.)  Two lists share pb.wl: searches run by a
    // warp each from the front, searches run by a group of kGroupLanes lanes each from the back.  Groups serve the pass in
    // which (nearly) every query of a unit searches — its first — where throughput counts (measured: 1.5x faster there);
    // the handful of searches of a later pass are latency bound and finish sooner with a warp each.
    if (sm.first_pass && w1 >= 0 && (w1 & 0xffff) <= kThreadScanBins) pb.wl[NQ - 1 - atomicAdd(&cta.wl_tn[0], 1)] = v;
    else pb.wl[atomicAdd(&cta.wl_n[0], 1)] = v;
    if (bv.timers && w1 >= 0) atomicAdd(&cta.dbg[0], w1 & 0xffff);
  }
  __syncthreads();
  LINS_TICK(3 + toff);

  if (cta.any_indexed) {
    // ---- A3/A4 fast path.  Scalar preparation (certificates, atan2f, asinf, bounds: P1 above, P3 below) runs one
    // THREAD per query so that all queries proceed in parallel; the memory scans (P2, P4) run one WARP per query.
    if (bv.timers && threadIdx.x == 0) {
      atomicAdd((unsigned long long*)&bv.timers[10], (unsigned long long)cta.wl_n[0]);
      atomicAdd((unsigned long long*)&bv.timers[12], (unsigned long long)cta.dbg[0]);
    }
    const float gate = sqrtf(nearf);
    // result of a closest-point search -> the query's state (one thread)
    auto nn_finish = [&](int v, const AzIndex& ix, const float4 s, const Top3& top, int w1, float Bout) {
      const unsigned long long k1 = top.k1;
      const int p1 = top.p1;
      const float d1 = __uint_as_float((unsigned)(k1 >> 32));
      const bool acc1 = k1 != kKeyMax && p1 >= 0 && (double)d1 < kp.nearest_sq;
      // accepted: what everything but the two front-runners exceeded; else the slack of "nothing within the gate"
      const float bound1 = w1 < 0 ? -1.f : acc1 ? cert_bound(top.d3, Bout) : rejected_slack((unsigned)(k1 >> 32), Bout, gate);
      pb.qref[v] = make_float4(s.x, s.y, s.z, bound1);
      pb.qext[v].y = __int_as_float(acc1 ? top.p2 : -1);
      pb.pos[3 * v] = acc1 ? p1 : -1;
      pb.qccr[v] = acc1 ? ((slot_ring(ix.pts[p1].w) << 24) | (int)(unsigned)(k1 & 0xffffffffu)) : -1;
    };
    // P2, group-level list first: groups of kGroupLanes lanes pull queries (a warp whose groups find the list empty goes
    // straight to the warp-level list)
    constexpr int G = kGroupLanes;
    const int sub = lane & (G - 1), glead = lane & ~(G - 1);
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << glead);
    for (;;) {
      int k = 0;
      if (sub == 0) k = atomicAdd(&cta.wl_thead[0], 1);
      k = __shfl_sync(gmask, k, glead);
      if (k >= cta.wl_tn[0]) break;
      const int v = pb.wl[NQ - 1 - k];
      const int sl = slot_of(v, Q), i = v - sl * Q;
      const Smem& sm = slots[sl];
      const AzIndex ix = az_index_of(sm, bv, i < sm.ns);
      const float4 s = pb.sel[v];
      const int w1 = pb.qw[v].x;
      const float4 qag = pb.qa[v];
      const Top3 top = az_scan_nn_group<G>(ix, s, w1, qag.y, qag.w, sub, gmask);
      if (sub == 0) nn_finish(v, ix, s, top, w1, qag.w);
    }
    // P2: warps pull queries from the work list (the per-query cost is heavy-tailed; a static split leaves warps idle)
    for (;;) {
      int k = 0;
      if (lane == 0) k = atomicAdd(&cta.wl_head[0], 1);
      k = __shfl_sync(0xffffffffu, k, 0);
      if (k >= cta.wl_n[0]) break;
      const int v = pb.wl[k];
      const int sl = slot_of(v, Q), i = v - sl * Q;
      const Smem& sm = slots[sl];
      int w1 = pb.qw[v].x;
      Top3 top;
      top.init();
      const bool surf = i < sm.ns;
      const AzIndex ix = az_index_of(sm, bv, surf);
      const float4 s = pb.sel[v];
      float4 qa = pb.qa[v];
      if (w1 >= 0) {
        const int wp = sm.first_pass ? az_probe_window(ix, qa, w1) : -1;
        if (wp >= 0) {  // wide window (no usable previous answer): probe first, then search inside the implied window
          const Top3 pr = az_scan_nn(ix, s, wp, qa.y, sqrtf(widen(kProbeSq)));  // (a probe needs no exactness: anything it finds is an upper bound)
          if (pr.p1 >= 0) {
            w1 = az_nn_window(ix, az_seed_bound(ix, s, pr.p1, nearf), qa);
            if (lane == 0) pb.qa[v] = qa;
          }
        }
        top = az_scan_nn(ix, s, w1, qa.y, qa.w);
      }
      if (lane == 0) nn_finish(v, ix, s, top, w1, qa.w);
    }
    __syncthreads();
    LINS_TICK(4 + toff);
    for (int v = threadIdx.x; v < NQ; v += kThreads) {  // P3
      const int sl = slot_of(v, Q), i = v - sl * Q;
      const Smem& sm = slots[sl];
      if (!sm.run || !sm.az_ok || i >= sm.ns + sm.nc || (sm.iter % kp.icp_freq) != 0) continue;
      const int lvl = pb.qw[v].x;
      if (lvl == -2) continue;
      const bool surf = i < sm.ns;
      const int ccr = pb.qccr[v];
      if (ccr < 0) {  // no closest point within the gate: nothing to walk
        pb.pos[3 * v + 1] = -1; pb.pos[3 * v + 2] = -1;
        pb.qw[v] = make_int4(-1, 0, 0, 0);
        if (surf) { int* o = bv.ind_s + 3 * (size_t)(sm.qs0 + i); o[0] = -1; o[1] = -1; o[2] = -1; }
        else { int* o = bv.ind_c + 2 * (size_t)(sm.qc0 + i - sm.ns); o[0] = -1; o[1] = -1; }
        continue;
      }
      int w2 = 0, w3 = 0;
      float B2 = 0.f, B3 = 0.f;
      const bool seeded = !sm.first_pass;
      const int sd2 = seeded ? pb.pos[3 * v + 1] : -1, sd3 = seeded ? pb.pos[3 * v + 2] : -1;
      const int c = ccr & 0x00ffffff, cr = (int)((unsigned)ccr >> 24);
      const AzIndex ix = az_index_of(sm, bv, surf);
      const int p1 = pb.pos[3 * v];
      if (surf) az_prepare_walk<true>(ix, pb.sel[v], pb.qa[v], p1, c, cr, sd2, sd3, min(sm.ns, sm.Ts), nearf, w2, w3, B2, B3);
      else az_prepare_walk<false>(ix, pb.sel[v], pb.qa[v], p1, c, cr, sd2, sd3, min(sm.nc, sm.Tc), nearf, w2, w3, B2, B3);
      pb.qw[v] = make_int4(1, w2, w3, ccr);
      reinterpret_cast<float2*>(pb.key)[v] = make_float2(B2, B3);
      if (sm.first_pass && (w2 & 0xffff) <= kThreadWalkBins && (w3 & 0xffff) <= kThreadWalkBins) pb.wl[NQ - 1 - atomicAdd(&cta.wl_tn[1], 1)] = v;
      else pb.wl[atomicAdd(&cta.wl_n[1], 1)] = v;
      if (bv.timers) atomicAdd(&cta.dbg[1], (w2 & 0xffff) + 4 * (w3 & 0xffff));
    }
    __syncthreads();
    if (bv.timers && threadIdx.x == 0) {
      atomicAdd((unsigned long long*)&bv.timers[11], (unsigned long long)cta.wl_n[1]);
      atomicAdd((unsigned long long*)&bv.timers[13], (unsigned long long)cta.dbg[1]);
    }
    // result of a query's walks -> its state + the correspondence IDs (one thread)
    auto walk_finish = [&](int v, const Smem& sm, int i, bool surf, const float4 s, int ccr, const WalkOut& wo) {
      pb.pos[3 * v + 1] = wo.pos2; pb.pos[3 * v + 2] = wo.pos3;
      pb.qref2[v] = make_float4(s.x, s.y, s.z, wo.bound2);
      float4 ex = pb.qext[v];  // (.y = the closest point's runner-up, written by P2 or kept from an earlier pass)
      ex.x = wo.bound3; ex.z = __int_as_float(wo.run2); ex.w = __int_as_float(wo.run3);
      pb.qext[v] = ex;
      const int i1 = ccr & 0x00ffffff;
      if (surf) { int* o = bv.ind_s + 3 * (size_t)(sm.qs0 + i); o[0] = i1; o[1] = wo.i2; o[2] = wo.i3; }
      else { int* o = bv.ind_c + 2 * (size_t)(sm.qc0 + i - sm.ns); o[0] = i1; o[1] = wo.i2; }
    };
    // P4, group-level list first.  :859 / :983 loop-bound quirk (+ OOB clamp): forward candidates count only below the
    // QUERY count
    for (;;) {
      int k = 0;
      if (sub == 0) k = atomicAdd(&cta.wl_thead[1], 1);
      k = __shfl_sync(gmask, k, glead);
      if (k >= cta.wl_tn[1]) break;
      const int v = pb.wl[NQ - 1 - k];
      const int sl = slot_of(v, Q), i = v - sl * Q;
      const Smem& sm = slots[sl];
      const int4 w = pb.qw[v];
      const bool surf = i < sm.ns;
      const float2 B = reinterpret_cast<const float2*>(pb.key)[v];
      const float4 s = pb.sel[v];
      const AzIndex ix = az_index_of(sm, bv, surf);
      const WalkOut wo = surf ? az_scan_walk_group<true, G>(ix, s, w.w, w.y, w.z, min(sm.ns, sm.Ts), nearf, B.x, B.y, sub, gmask)
                              : az_scan_walk_group<false, G>(ix, s, w.w, w.y, w.z, min(sm.nc, sm.Tc), nearf, B.x, B.y, sub, gmask);
      if (sub == 0) walk_finish(v, sm, i, surf, s, w.w, wo);
    }
    for (;;) {  // P4: same work-list scheme
      int k = 0;
      if (lane == 0) k = atomicAdd(&cta.wl_head[1], 1);
      k = __shfl_sync(0xffffffffu, k, 0);
      if (k >= cta.wl_n[1]) break;
      const int v = pb.wl[k];
      const int sl = slot_of(v, Q), i = v - sl * Q;
      const Smem& sm = slots[sl];
      const int4 w = pb.qw[v];
      const bool surf = i < sm.ns;
      const float2 B = reinterpret_cast<const float2*>(pb.key)[v];
      const int w2 = w.y, w3 = w.z;
      const float4 s = pb.sel[v];
      const AzIndex ix = az_index_of(sm, bv, surf);
      const WalkOut wo = surf ? az_scan_walk<true>(ix, s, w.w, w2, w3, min(sm.ns, sm.Ts), nearf, B.x, B.y)
                              : az_scan_walk<false>(ix, s, w.w, w2, w3, min(sm.nc, sm.Tc), nearf, B.x, B.y);
      if (lane == 0) walk_finish(v, sm, i, surf, s, w.w, wo);
    }
  }
  if (cta.any_legacy) {
    // ---- legacy: brute-force exact 1-NN + literal sequential walks, unit by unit ----------------------------------
    for (int sl = 0; sl < bv.nslots; ++sl) {
      const Smem& sm = slots[sl];
      if (!sm.run || sm.az_ok || (sm.iter % kp.icp_freq) != 0) continue;
      const float4* __restrict__ nnS = bv.nn_s ? bv.nn_s + bv.nn_s_off[sm.scan] : bv.ts + sm.ts0;
      const float4* __restrict__ nnC = bv.nn_c ? bv.nn_c + bv.nn_c_off[sm.scan] : bv.tc + sm.tc0;
      const int TnS = bv.nn_s ? bv.nn_s_off[sm.scan + 1] - bv.nn_s_off[sm.scan] : sm.Ts;
      const int TnC = bv.nn_c ? bv.nn_c_off[sm.scan + 1] - bv.nn_c_off[sm.scan] : sm.Tc;
      if (sm.ns > 0 && TnS > 0) nn_brute(pb.sel + sl * Q, pb.key + sl * Q, sm.ns, nnS, TnS);
      if (sm.nc > 0 && TnC > 0) nn_brute(pb.sel + sl * Q + sm.ns, pb.key + sl * Q + sm.ns, sm.nc, nnC, TnC);
    }
    __syncthreads();
    for (int sl = 0; sl < bv.nslots; ++sl) {
      const Smem& sm = slots[sl];
      if (!sm.run || sm.az_ok || (sm.iter % kp.icp_freq) != 0) continue;
      const float4* __restrict__ tgtS = bv.ts + sm.ts0;
      const float4* __restrict__ tgtC = bv.tc + sm.tc0;
      const int fwdS = min(sm.ns, sm.Ts), fwdC = min(sm.nc, sm.Tc);
      for (int i = threadIdx.x; i < sm.ns + sm.nc; i += kThreads) {
        const int v = sl * Q + i;
        const bool surf = i < sm.ns;
        const unsigned long long k1 = pb.key[v];
        const float d1 = __uint_as_float((unsigned)(k1 >> 32));
        const int c = (int)(unsigned)(k1 & 0xffffffffu);
        const bool found = (k1 != kKeyMax) && ((double)d1 < kp.nearest_sq) && c < (surf ? sm.Ts : sm.Tc);
        int i1 = -1, i2 = -1, i3 = -1;
        if (found) {
          i1 = c;
          const float4 s = pb.sel[v];
          if (surf) walk_seq<true>(s, c, tgtS, sm.Ts, fwdS, nearf, i2, i3);
          else walk_seq<false>(s, c, tgtC, sm.Tc, fwdC, nearf, i2, i3);
        }
        pb.pos[3 * v] = i1; pb.pos[3 * v + 1] = i2; pb.pos[3 * v + 2] = i3;
        if (surf) { int* o = bv.ind_s + 3 * (size_t)(sm.qs0 + i); o[0] = i1; o[1] = i2; o[2] = i3; }
        else { int* o = bv.ind_c + 2 * (size_t)(sm.qc0 + i - sm.ns); o[0] = i1; o[1] = i2; }
      }
    }
  }
  __syncthreads();
  LINS_TICK(5 + toff);
  // the searches of this pass are over: reset the work lists for the next pass
  if (threadIdx.x == kThreads - 1) { cta.wl_n[0] = 0; cta.wl_n[1] = 0; cta.wl_tn[0] = 0; cta.wl_tn[1] = 0; cta.wl_head[0] = 0; cta.wl_head[1] = 0; cta.wl_thead[0] = 0; cta.wl_thead[1] = 0; cta.dbg[0] = 0; cta.dbg[1] = 0; }

  // ---- A5/A6 residuals + A7-A9 fold ------------------------------------------------------------------------------
  // tripod points: slots of the sorted copies after a fast-path search, otherwise original indices into the walk clouds.
  for (int v0 = warp * 32; v0 < NQ; v0 += kThreads) {
    const int sl = slot_of(v0, Q), i0 = v0 - sl * Q;
    const Smem& sm = slots[sl];
    const int ntot = sm.ns + sm.nc;
    if (!sm.run || i0 >= ntot) continue;  // (warp-uniform)
    const int v = v0 + lane, i = i0 + lane;
    const bool valid = i < ntot;
    const bool surf = i < sm.ns;
    const bool search = (sm.iter % kp.icp_freq) == 0;
    const bool weighted = sm.iter >= kp.icp_freq;
    const bool by_slot = search ? (sm.az_ok != 0) : (sm.pos_valid && sm.pos_is_slot);
    float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = false;
    if (valid) {
      const int i1 = pb.pos[3 * v], i2 = pb.pos[3 * v + 1], i3 = pb.pos[3 * v + 2];
      s = pb.sel[v];
      if (surf) {
        if (i2 >= 0 && i3 >= 0 && i1 >= 0 && i1 < sm.Ts && i2 < sm.Ts && i3 < sm.Ts) {
          const float4* t = by_slot ? bv.az_s + sm.ts0 : bv.ts + sm.ts0;
          ok = plane_residual(s, t[i1], t[i2], t[i3], weighted, coeff);
        }
      } else {
        if (i2 >= 0 && i1 >= 0 && i1 < sm.Tc && i2 < sm.Tc) {
          const float4* t = by_slot ? bv.az_c + sm.tc0 : bv.tc + sm.tc0;
          ok = line_residual(s, t[i1], t[i2], weighted, coeff);
        }
      }
    }
    double g[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, r = 0.0;
    if (ok) {
      if (MODE == MODE_ICP_REDUCE) jacobian_row_icp(pb.qpt[v], coeff, sm.phi, kp.scan_period, g, r);
      else jacobian_row(pb.qpt[v], coeff, sm.R, kp.lidar_scale, g, r);
    }
    const double tot = warp_fold_row(g, r);
    const unsigned mS = __ballot_sync(0xffffffffu, ok && surf), mC = __ballot_sync(0xffffffffu, ok && !surf);
    const int vw = sl * pb.nvw + (i0 >> 5);
    if (lane < kNAcc) pb.wacc[vw * kNAcc + lane] = tot;
    else if (lane == kNAcc) pb.wcnt[vw * 2] = __popc(mS);
    else if (lane == kNAcc + 1) pb.wcnt[vw * 2 + 1] = __popc(mC);
    if (MODE == MODE_ASSOC && valid) {
      if (surf) {
        const size_t o = (size_t)(sm.qs0 + i);
        if (bv.sel_s) { bv.sel_s[3 * o] = s.x; bv.sel_s[3 * o + 1] = s.y; bv.sel_s[3 * o + 2] = s.z; }
        if (bv.coeff_s) { bv.coeff_s[4 * o] = coeff.x; bv.coeff_s[4 * o + 1] = coeff.y; bv.coeff_s[4 * o + 2] = coeff.z; bv.coeff_s[4 * o + 3] = coeff.w; }
        if (bv.mask_s) bv.mask_s[o] = ok ? 1 : 0;
      } else {
        const size_t o = (size_t)(sm.qc0 + i - sm.ns);
        if (bv.sel_c) { bv.sel_c[3 * o] = s.x; bv.sel_c[3 * o + 1] = s.y; bv.sel_c[3 * o + 2] = s.z; }
        if (bv.coeff_c) { bv.coeff_c[4 * o] = coeff.x; bv.coeff_c[4 * o + 1] = coeff.y; bv.coeff_c[4 * o + 2] = coeff.z; bv.coeff_c[4 * o + 3] = coeff.w; }
        if (bv.mask_c) bv.mask_c[o] = ok ? 1 : 0;
      }
    }
  }
  __syncthreads();
  LINS_TICK(6 + toff);
}

}  // namespace lins_dev
