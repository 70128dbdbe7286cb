/*
 * lins_gpu.h — C-ABI of the B200-native LINS iterated-ESKF update path.
 *
 * This is the drop-in boundary: every entry point replaces one seam of the reference's
 * header-only class fusion::StateEstimator (reference paths relative to /root/reference/):
 *
 *   lins_gpu_set_map        <-> kdtreeCorner_/kdtreeSurf_->setInputCloud(...)
 *                               lins/include/StateEstimator.hpp:363-364, :1156-1160
 *   lins_gpu_ieskf          <-> StateEstimator::performIESKF()            StateEstimator.hpp:465-600
 *   lins_gpu_associate      <-> findCorrespondingSurfFeatures / findCorrespondingCornerFeatures
 *                               StateEstimator.hpp:829-953, :955-1063 (+ transformToStart :1066-1080)
 *   lins_gpu_estimate_transform <-> estimateTransform / calculateTransformation
 *                               StateEstimator.hpp:1163-1196, :1198-1320 (fallback + scan-2 initialiser)
 *   lins_gpu_update_map     <-> updatePointCloud() / transformToEnd()     StateEstimator.hpp:1083-1101, :1116-1161
 *   lins_gpu_batch_*        <-> the same performIESKF, for many independent (scan pair, prior) units
 *                               resident in HBM (offline / batched odometry)
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no Eigen / PCL / ROS / torch types.
 *   - every function returns 0 on success, <0 (LINS_E_*) on error; nothing throws.
 *     lins_gpu_last_error() returns a human readable message for the last failure on that ctx.
 *   - caller owns every host buffer; the library copies at call time and owns all device memory.
 *   - one ctx = one CUDA device + one CUDA stream; a ctx is not re-entrant, distinct ctxs are independent.
 *   - points are pcl::PointXYZI-layout compatible (32 B, 16-B aligned; lins/include/parameters.h:52):
 *       x@0 y@4 z@8 (pad) intensity@16 (pad..31); intensity = ring + SCAN_PERIOD*relTime
 *       (lins/src/image_projection_node.cpp:234, StateEstimator.hpp:647-650).
 *   - state vectors are 19 doubles in filter::GlobalState member order (KalmanFilter.hpp:110-115):
 *       rn[0..2] vn[3..5] qbn[6..9] (x,y,z,w — Eigen::Quaterniond::coeffs() order) ba[10..12] bw[13..15] gn[16..18]
 *   - covariances are 18x18 doubles, column-major (Eigen default), error-state order
 *       pos0 vel3 att6 acc9 gyr12 gra15 (KalmanFilter.hpp:38-45).
 */
#ifndef LINS_GPU_H_
#define LINS_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LINS_STATE_DIM 19
#define LINS_ERR_DIM 18
#define LINS_COV_SIZE 324
#define LINS_MAX_ITER 64

/* error codes */
#define LINS_OK 0
#define LINS_E_INVALID (-1)   /* bad argument */
#define LINS_E_CUDA (-2)      /* CUDA runtime failure (message in last_error) */
#define LINS_E_NOMAP (-3)     /* ieskf/associate before set_map */
#define LINS_E_TOOBIG (-4)    /* cloud larger than the ctx capacity */
#define LINS_E_NODEVICE (-5)  /* no usable CUDA device / sm_100 kernel image */

/* pcl::PointXYZI layout (parameters.h:52 `typedef pcl::PointXYZI PointType`). */
typedef struct lins_point {
  float x, y, z, pad0;
  float intensity, pad1, pad2, pad3;
} lins_point;

/* The 7 globals the hot path reads (parameters.h:104-153, exp_port.yaml:9-20) + test hooks. */
typedef struct lins_params {
  int32_t num_iter;        /* NUM_ITER, shipped 30 (exp_port.yaml:18); <= LINS_MAX_ITER */
  int32_t icp_freq;        /* ICP_FREQ, shipped 1 */
  double nearest_feature_search_sq_dist; /* NEAREST_FEATURE_SEARCH_SQ_DIST, shipped 25 */
  double lidar_std;        /* LIDAR_STD, shipped 0.01 */
  double lidar_scale;      /* LIDAR_SCALE, shipped 1 */
  double scan_period;      /* SCAN_PERIOD, shipped 0.1 */
  int32_t verbose;         /* VERBOSE (unused by the device path) */
  int32_t force_all_iters; /* test hook: 1 = ignore the ||dx||<=1e-2 exit (StateEstimator.hpp:576) so exactly
                              num_iter iterations run ("10 ESKF iters forced", BASELINE.json configs[0]) */
} lins_params;

/* What performIESKF would have logged (StateEstimator.hpp:485-496, :560, :567, :586). */
typedef struct lins_report {
  int32_t iters;      /* iterations executed (count of A2-A10 passes) */
  int32_t converged;  /* hasConverged, StateEstimator.hpp:576-578 */
  int32_t diverged;   /* hasDiverged, StateEstimator.hpp:559-570 (host must run the ICP fallback) */
  int32_t has_nan;    /* the NaN branch of the divergence test, StateEstimator.hpp:552-563 */
  int32_t m_surf[LINS_MAX_ITER];   /* accepted plane measurements per iteration */
  int32_t m_corner[LINS_MAX_ITER]; /* accepted line measurements per iteration */
  double residual_norm[LINS_MAX_ITER]; /* ||residual_|| per iteration */
  double update_norm[LINS_MAX_ITER];   /* ||updateVec_|| per iteration */
} lins_report;

/* Fixed 64-byte per-scan record of the batched mode (the unit the multi-GPU pose gather moves). */
typedef struct lins_scan_result {
  int32_t scan_id;
  uint16_t iters;
  uint16_t flags;      /* bit0 converged, bit1 diverged, bit2 has_nan */
  double pose[7];      /* rn (3) + qbn (x,y,z,w) of the updated filter state */
} lins_scan_result;

/* A batch of independent units: scan i = (new scan's query features, last scan's target features, prior).
   Clouds are concatenated; *_off has n_scans+1 entries (CSR style). */
typedef struct lins_batch_desc {
  int32_t n_scans;
  const lins_point* surf_flat;         const int32_t* surf_flat_off;         /* queries: surfPointsFlat_ */
  const lins_point* corner_sharp;      const int32_t* corner_sharp_off;      /* queries: cornerPointsSharp_ */
  const lins_point* surf_less_flat;    const int32_t* surf_less_flat_off;    /* targets: last surfPointsLessFlat_ */
  const lins_point* corner_less_sharp; const int32_t* corner_less_sharp_off; /* targets: last cornerPointsLessSharp_ */
  const double* state_in;              /* n_scans x 19 : filter_->state_ */
  const double* cov_in;                /* n_scans x 324: filter_->covariance_ */
  int32_t point_format;                /* LINS_POINTS_XYZI32 (0, default): the four clouds are pcl::PointXYZI records as typed
                                          above; LINS_POINTS_PACKED16: they are 16-byte (x, y, z, intensity) float records (cast
                                          the pointers) — what the device keeps anyway, so a caller that stores its clouds this
                                          way (and page-locks them) uploads with no host pass and half the PCIe bytes */
} lins_batch_desc;
#define LINS_POINTS_XYZI32 0
#define LINS_POINTS_PACKED16 1

typedef struct lins_ctx lins_ctx;

/* Create a context on CUDA device `device`. `stream` is a cudaStream_t passed as void* (NULL = the library
   creates its own non-blocking stream). Fails with LINS_E_NODEVICE when no CUDA device is usable: there is
   no CPU fallback. */
int lins_gpu_create(const lins_params* params, int device, void* stream, lins_ctx** out);
void lins_gpu_destroy(lins_ctx* ctx);
const char* lins_gpu_last_error(const lins_ctx* ctx);
int lins_gpu_set_params(lins_ctx* ctx, const lins_params* params);

/* ≙ kdtreeSurf_->setInputCloud(surfPointsLessFlat_), kdtreeCorner_->setInputCloud(cornerPointsLessSharp_).
   Uploads both target clouds and builds the device search index. */
int lins_gpu_set_map(lins_ctx* ctx, const lins_point* surf_less_flat, int n_surf,
                     const lins_point* corner_less_sharp, int n_corner);

/* ≙ performIESKF(): all iterations on device, no host round trip inside the loop.
   state_in/cov_in = filter_->state_/covariance_; state_out/cov_out = what filter_->update(...) would store
   when not diverged (linState_, Joseph-form Pk_). When rep->diverged is set, state_out = the prior
   (filterState) and cov_out = cov_in: the caller then runs lins_gpu_estimate_transform (the reference's
   "======Using ICP Method======" branch, StateEstimator.hpp:585-592). */
int lins_gpu_ieskf(lins_ctx* ctx, const lins_point* surf_flat, int n_surf, const lins_point* corner_sharp,
                   int n_corner, const double* state_in, const double* cov_in, double* state_out,
                   double* cov_out, lins_report* rep);

/* ≙ one call each of findCorrespondingSurfFeatures / findCorrespondingCornerFeatures at iteration `iter`
   with linState_ = lin_state (only rn/qbn are read). Dense (uncompacted) per-query outputs; any output
   pointer may be NULL. ind: -1 = none. mask = the accept test (s > 0.1 && res != 0). coeff = (s*jac, s*res).
   sel = pointSel (the de-skewed query, f32). Indices persist inside the ctx between calls like
   pointSearchSurfInd1/2/3 do, so iter % icp_freq != 0 reuses them. */
int lins_gpu_associate(lins_ctx* ctx, const lins_point* surf_flat, int n_surf, const lins_point* corner_sharp,
                       int n_corner, const double* lin_state, int iter, int32_t* surf_ind /*3*n_surf*/,
                       int32_t* corner_ind /*2*n_corner*/, float* surf_coeff /*4*n_surf*/,
                       float* corner_coeff /*4*n_corner*/, uint8_t* surf_mask, uint8_t* corner_mask,
                       float* surf_sel /*3*n_surf*/, float* corner_sel /*3*n_corner*/);

/* ≙ estimateTransform(scan_last_, scan_new_, t, q): 6-DoF Gauss-Newton ICP on the same association.
   pose_io = t (3) + q (x,y,z,w). iters_out/converged_out may be NULL. */
int lins_gpu_estimate_transform(lins_ctx* ctx, const lins_point* surf_flat, int n_surf,
                                const lins_point* corner_sharp, int n_corner, double* pose_io,
                                int* iters_out, int* converged_out);

/* ≙ updatePointCloud(): transformToEnd of the new scan's less-* clouds with linState_ = lin_state, written
   back in place to the host arrays (the reference overwrites scan_new_->*_ in place), and — iff
   n_corner >= 5 && n_surf >= 20 (StateEstimator.hpp:1156-1157) — installed as the new device map.
   Returns 1 in *map_replaced when the index was rebuilt. */
int lins_gpu_update_map(lins_ctx* ctx, lins_point* surf_less_flat, int n_surf, lins_point* corner_less_sharp,
                        int n_corner, const double* lin_state, int* map_replaced);

/* The same with the device-resident options: lin_state == NULL uses the posterior the last lins_gpu_ieskf left on the
   device; surf_out / corner_out receive the transformed clouds (may alias the inputs; NULL = keep them on the device only:
   no D2H, no stream synchronisation — the call returns with the refresh queued).  *map_replaced as above. */
int lins_gpu_update_map_ex(lins_ctx* ctx, const lins_point* surf_less_flat, int n_surf, const lins_point* corner_less_sharp,
                           int n_corner, const double* lin_state, lins_point* surf_out, lins_point* corner_out,
                           int* map_replaced);

/* Batched mode. upload: pack + H2D, resident afterwards. run: launch the fused kernel over the resident
   batch on the ctx stream (asynchronous). download: D2H of results + stream sync; any pointer may be NULL. */
int lins_gpu_batch_upload(lins_ctx* ctx, const lins_batch_desc* batch);
int lins_gpu_batch_run(lins_ctx* ctx);
/* cumulative number of points lins_gpu_batch_upload moved as host-packed 16-B records / as raw 32-B records (caller-pinned
   clouds are split between the pack threads and the copy engine at run time): what the PCIe byte count of a job is made of. */
int lins_gpu_batch_upload_stats(lins_ctx* ctx, int64_t* packed_points, int64_t* raw_points);
int lins_gpu_batch_download(lins_ctx* ctx, double* state_out /*n x 19*/, double* cov_out /*n x 324*/,
                            lins_scan_result* results /*n*/, lins_report* reports /*n, optional*/);
/* The correspondence IDs the resident batch holds after a run — pointSearchSurfInd1/2/3 and pointSearchCornerInd1/2
   (StateEstimator.hpp:1459-1465) of every unit's LAST search iteration, concatenated in batch order (3 per surf query,
   2 per corner query, -1 = none).  Parity hook for the batched mode; either pointer may be NULL. */
int lins_gpu_batch_download_indices(lins_ctx* ctx, int32_t* surf_ind, int32_t* corner_ind);
/* upload + run + download in one call (the end-to-end entry point). */
int lins_gpu_ieskf_batch(lins_ctx* ctx, const lins_batch_desc* batch, double* state_out, double* cov_out,
                         lins_scan_result* results);
/* device pointer of the resident lins_scan_result array (n_scans x 64 B) for a zero-copy pose gather. */
int lins_gpu_batch_results_device(lins_ctx* ctx, void** dev_ptr, int* n_scans);

/* Optional: page-lock (pin) a host buffer the caller owns — a thin wrapper over cudaHostRegister so that a caller without
   a CUDA binding of its own (the reference is plain C++ / ROS / PCL) can pin the storage of its point clouds once.
   lins_gpu_batch_upload detects pinned clouds and then DMAs the raw 32-B records straight from the caller's memory and
   packs them on the device: no host pass over the points (pageable clouds are packed by host threads into internal
   pinned staging first).  Unregister before freeing the buffer. */
int lins_gpu_host_register(void* ptr, size_t bytes);
int lins_gpu_host_unregister(void* ptr);

/* Split "Jacobian kernel" (SURVEY.md §8(d) unit U1): residual + Jacobian row + 29-scalar reduction over the
   resident batch given the correspondence IDs of iteration `iter` of each scan's current linearisation
   point. Used for the HBM-roofline measurement; results land in an internal n x 29 accumulator array. */
int lins_gpu_batch_jacobian_pass(lins_ctx* ctx, double* accum_out /*n x 29 or NULL*/);

/* ---- row F2: the mapping node's scan-to-map refinement -------------------------------------
   lins/src/lidar_mapping_node.cpp: scan2MapOptimization :1635-1652, cornerOptimization :1351-1461,
   surfOptimization :1463-1524, LMOptimization :1526-1633, pointAssociateToMap :594-608.
   All point clouds are in the mapping node's frame convention (the YZX clouds the estimator publishes);
   transform = transformTobeMapped = (rx, ry, rz, tx, ty, tz), f32 like the reference. */
#define LINS_MAP_MAX_ITER 10
typedef struct lins_map_report {
  int32_t iters;       /* LM iterations executed (<= 10) */
  int32_t converged;   /* deltaR < 0.05 deg && deltaT < 0.05 cm reached (:1628) */
  int32_t degenerate;  /* isDegenerate of iteration 0 (:1606-1617) */
  int32_t skipped;     /* map too small: cornerFromMapDSNum <= 10 || surfFromMapDSNum <= 100 (:1636) */
  int32_t n_sel[LINS_MAP_MAX_ITER];   /* laserCloudSelNum per iteration (< 50 => that LM step is skipped, :1535) */
  float delta_r[LINS_MAP_MAX_ITER];   /* deg */
  float delta_t[LINS_MAP_MAX_ITER];   /* cm */
} lins_map_report;

/* ≙ kdtreeCornerFromMap->setInputCloud(laserCloudCornerFromMapDS), kdtreeSurfFromMap->setInputCloud(...) (:1637-1638):
   uploads both map clouds and builds the device search structure. */
int lins_gpu_map_set(lins_ctx* ctx, const lins_point* corner_from_map, int n_corner, const lins_point* surf_from_map,
                     int n_surf);
/* ≙ the iteration loop of scan2MapOptimization (:1640-1648): up to 10 x (cornerOptimization, surfOptimization,
   LMOptimization) against the map set by lins_gpu_map_set; 5-NN search, line / plane fits, coefficients and the
   A^T A / A^T B reduction on the device, the 6x6 step on the host.  transform_io: transformTobeMapped in / out.
   (transformUpdate, :538-577, blends IMU roll / pitch afterwards and stays with the caller.) */
int lins_gpu_scan2map(lins_ctx* ctx, const lins_point* corner_last, int n_corner, const lins_point* surf_last, int n_surf,
                      float* transform_io /*6*/, lins_map_report* rep);
/* ≙ one cornerOptimization + surfOptimization pass at `transform`, dense per-point outputs (any pointer may be NULL):
   knn = pointSearchInd (5 per point, ascending distance, -1 = fewer than 5 map points), coeff = (s*la, s*lb, s*lc,
   s*ld2) resp. (s*pa, s*pb, s*pc, s*pd2), mask = the point was pushed to laserCloudOri (s > 0.1). */
int lins_gpu_map_associate(lins_ctx* ctx, const lins_point* corner_last, int n_corner, const lins_point* surf_last,
                           int n_surf, const float* transform /*6*/, int32_t* corner_knn /*5*n_corner*/,
                           int32_t* surf_knn /*5*n_surf*/, float* corner_coeff /*4*n_corner*/, float* surf_coeff /*4*n_surf*/,
                           uint8_t* corner_mask, uint8_t* surf_mask);

/* block until everything queued on the ctx stream has finished */
int lins_gpu_sync(lins_ctx* ctx);

/* diagnostics: per-phase SM-cycle counters of the fused kernel (enable, then read after a batch_run) */
int lins_gpu_debug_phase_cycles(lins_ctx* ctx, int enable, long long* out /*64 or NULL*/);

/* kernels launched by this ctx since creation (for bench.py's gpu_launches). */
int64_t lins_gpu_launch_count(const lins_ctx* ctx);
/* library/ABI version, = 1 */
int lins_gpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LINS_GPU_H_ */
