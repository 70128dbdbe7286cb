// tools/synth/lins_sequence.cpp — offline sequence driver: what LinsFusion does between the
// ROS callbacks and the estimator (lins/src/lib/Estimator.cpp:204-252): for every lidar scan, feed the buffered
// IMU samples through processImu, then the segmented cloud through processPCL.  Inputs are synthetic (a smooth
// trajectory through the seeded world of lins_synth.cpp; IMU at 400 Hz; raw scans pushed through the restated
// image projection).  The estimator is the C++ shim fusion::StateEstimator whose hot seams run on the GPU through
// the C-ABI, so this file links against liblins_gpu.so.  Every performIESKF call's inputs / outputs are recorded
// so that tests can replay them through the CPU oracle.  Never includes anything from oracle/.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lins_synth.cpp"

#include "../../lins---lidar-inertial-slam_b200/csrc/host/state_estimator.hpp"
#include "../../lins---lidar-inertial-slam_b200/csrc/host/rosbag_reader.hpp"
#include <algorithm>

using lins::fusion::StateEstimator;

namespace {

struct SeqRecord {
  // one recorded performIESKF unit per RUNNING scan
  std::vector<lins_point> surfFlat, cornerSharp, surfLessFlat, cornerLessSharp;
  std::vector<int32_t> offSF{0}, offCS{0}, offSL{0}, offCL{0};
  std::vector<double> state_in, cov_in, state_out, cov_out;
  std::vector<int32_t> iters, flags, scan_index;
  std::vector<double> rel_true;  // true relative pose of that scan: t (3) + q (x,y,z,w)
  std::vector<int32_t> status;   // estimator status after each scan
  std::vector<double> global_est, global_true;  // per scan: position (3) + quaternion (x,y,z,w)
};

void append_cloud(std::vector<lins_point>& dst, std::vector<int32_t>& off, const Cloud& c) {
  dst.insert(dst.end(), c.points.begin(), c.points.end());
  off.push_back((int32_t)dst.size());
}

struct Traj {  // smooth body-frame velocity / yaw-rate profile
  double v0, va, vw, w0, wa, ww;
  V3D vbody(double t) const { return V3D(v0 + va * std::sin(vw * t), 0.1 * std::sin(0.7 * vw * t), 0.0); }
  V3D vdot(double t) const { return V3D(va * vw * std::cos(vw * t), 0.07 * vw * std::cos(0.7 * vw * t), 0.0); }
  V3D wbody(double t) const { return V3D(0.0, 0.0, w0 + wa * std::sin(ww * t)); }
};

// one simulated sweep: the raw cloud (sensor frame, PointXYZI with intensity 0 like a driver that has none), the IMU samples
// taken during it (index 1..nimu; index 0 is the sample at the sweep's start) and the true poses at its two ends
struct Sweep {
  Cloud raw;
  std::vector<V3D> accs, gyrs;
  Pose start, end;
  double t_end = 0;
};

struct SimDrive {
  const lins_synth_cfg* cfg;
  Rng rng;
  LidarModel lm;
  World w;
  Traj tr;
  Pose P;
  double t = 0.0;
  static constexpr int nimu = 40;
  SimDrive(const lins_synth_cfg* c, uint64_t seed) : cfg(c), rng(seed) {
    lm = cfg->lidar == 1 ? LidarModel::dense64() : LidarModel::vlp16();
    w = make_world(rng, cfg->world);
    tr = Traj{rng.uni(1.0, 0.6 * cfg->v_max), rng.uni(0.2, 1.0), rng.uni(0.5, 1.5), rng.uni(-0.5, 0.5) * cfg->w_max, 0.5 * cfg->w_max, rng.uni(0.3, 1.0)};
    P.R = math_utils::rpy2Quat(V3D(0.0, 0.0, rng.uni(-M_PI, M_PI))).toRotationMatrix();
    P.p = V3D(rng.uni(-3, 3), rng.uni(-3, 3), 1.5);
  }
  double dt() const { return lm.scan_period / nimu; }
  void next(Sweep& sw) {
    const double T = lm.scan_period, d = dt();
    const int N = lm.scan_num;
    const V3D g_w(0, 0, -filter::G0);
    const V3D ba(0.0, 0.0, 0.0), bw(0.0, 0.0, 0.0);  // the shim initialises its biases from FilterParams
    std::vector<Pose> poses(nimu + 1);
    sw.accs.assign(nimu + 1, V3D()); sw.gyrs.assign(nimu + 1, V3D());
    poses[0] = P;
    sw.start = P;
    for (int i = 0; i <= nimu; ++i) {
      const double ti = t + i * d;
      const V3D wb = tr.wbody(ti), vb = tr.vbody(ti);
      sw.accs[i] = cross(wb, vb) + tr.vdot(ti) - poses[i].R.transpose() * g_w + ba + V3D(0.02 * rng.gauss(), 0.02 * rng.gauss(), 0.02 * rng.gauss());
      sw.gyrs[i] = wb + bw + V3D(5e-4 * rng.gauss(), 5e-4 * rng.gauss(), 5e-4 * rng.gauss());
      if (i < nimu) {  // midpoint step of the true motion
        const double tm = ti + 0.5 * d;
        poses[i + 1].p = poses[i].p + d * (poses[i].R * math_utils::axis2Quat(0.5 * d * tr.wbody(tm)).toRotationMatrix() * tr.vbody(tm));
        poses[i + 1].R = poses[i].R * math_utils::axis2Quat(d * tr.wbody(tm)).toRotationMatrix();
      }
    }
    sw.raw.clear();
    for (int f = 0; f < N; ++f) {
      const double s = (double)f / N * nimu;
      const int i0 = std::min((int)s, nimu - 1);
      const double a = s - i0;
      Pose Pf;
      Pf.p = poses[i0].p + a * (poses[i0 + 1].p - poses[i0].p);
      Pf.R = poses[i0].R * math_utils::axis2Quat(a * d * tr.wbody(t + (i0 + 0.5 * a) * d)).toRotationMatrix();
      const double ori = -M_PI + (f + 0.25) * (2.0 * M_PI / N);
      for (int r = 0; r < lm.line_num; ++r) {
        const double el = (-(double)(lm.ang_bottom - 0.1f) + r * (double)lm.ang_res_y) * M_PI / 180.0;
        const V3D ds(std::cos(el) * std::cos(ori), -std::cos(el) * std::sin(ori), std::sin(el));
        double rg = raycast(w, Pf.p, Pf.R * ds);
        if (!(rg < 100.0)) continue;
        rg += cfg->range_noise * rng.gauss();
        sw.raw.push_back(makePoint((float)(rg * ds.x()), (float)(rg * ds.y()), (float)(rg * ds.z()), 0.f));
      }
    }
    P = poses[nimu];
    t += T;
    sw.end = P;
    sw.t_end = t;
  }
};

// what LinsFusion::processPointClouds does with one scan (Estimator.cpp:204-252) + the bookkeeping of the recorder:
// `imu_dt / acc / gyr` = the processImu calls between the previous scan and this one, in order
void feed_scan(StateEstimator& est, ImageProjection& ip, SeqRecord* rec, int k, double scan_time, const Cloud& raw, const std::vector<double>& imu_dt,
               const std::vector<V3D>& acc, const std::vector<V3D>& gyr, const Pose* start_true, const Pose* end_true, bool verbose) {
  ip.process(raw);
  for (size_t i = 0; i < imu_dt.size(); ++i) est.processImu(imu_dt[i], acc[i], gyr[i]);
  const bool will_run = est.status_ == StateEstimator::STATUS_RUNNING;
  double s_in[19];
  filter::Cov18 P_in;
  std::vector<lins_point> mapS, mapC;
  if (will_run) {
    est.filter_->state_.toArray(s_in);
    P_in = est.filter_->covariance_;
    mapS = est.scan_last_->surfPointsLessFlat_.points;
    mapC = est.scan_last_->cornerPointsLessSharp_.points;
  }
  est.last_report_.iters = 0;
  const V3D a_last = acc.empty() ? V3D(0, 0, filter::G0) : acc.back(), g_last = gyr.empty() ? V3D() : gyr.back();
  est.processPCL(scan_time, lins::sensor_utils::Imu(scan_time, a_last, g_last), ip.segmentedCloud, ip.segMsg, ip.outlierCloud);
  rec->status.push_back((int)est.status_);
  if (verbose) std::fprintf(stderr, "[lins_seq] scan %d: status %d, IESKF iterations %d, %zu segmented points\n", k, (int)est.status_, (int)est.last_report_.iters, ip.segmentedCloud.size());
  if (will_run && est.last_report_.iters > 0) {
    // processScan swapped the scans: scan_last_ now IS the scan whose features were the queries
    append_cloud(rec->surfFlat, rec->offSF, est.scan_last_->surfPointsFlat_);
    append_cloud(rec->cornerSharp, rec->offCS, est.scan_last_->cornerPointsSharp_);
    Cloud ms, mc; ms.points = mapS; mc.points = mapC;
    append_cloud(rec->surfLessFlat, rec->offSL, ms);
    append_cloud(rec->cornerLessSharp, rec->offCL, mc);
    rec->state_in.insert(rec->state_in.end(), s_in, s_in + 19);
    rec->cov_in.insert(rec->cov_in.end(), P_in.data(), P_in.data() + 324);
    // filter_->state_ after performIESKF but BEFORE reset(1) is linState_ (not diverged) — reset(1) zeroes rn/qbn,
    // so take the relative pose from linState_ and the remaining blocks from the filter
    double s_out[19];
    lins::filter::GlobalState so = est.linState_;
    so.toArray(s_out);
    rec->state_out.insert(rec->state_out.end(), s_out, s_out + 19);
    rec->iters.push_back(est.last_report_.iters);
    rec->flags.push_back((est.last_report_.converged ? 1 : 0) | (est.last_report_.diverged ? 2 : 0) | (est.last_report_.has_nan ? 4 : 0));
    rec->scan_index.push_back(k);
    double tv[7] = {0, 0, 0, 0, 0, 0, 1};
    if (start_true && end_true) {  // true relative pose of this sweep
      M3D Rrel = start_true->R.transpose() * end_true->R;
      V3D trel = start_true->R.transpose() * (end_true->p - start_true->p);
      Q4D qrel = math_utils::R2Quat(Rrel);
      const double v[7] = {trel.x(), trel.y(), trel.z(), qrel.x(), qrel.y(), qrel.z(), qrel.w()};
      std::memcpy(tv, v, sizeof(v));
    }
    rec->rel_true.insert(rec->rel_true.end(), tv, tv + 7);
  }
  const double ge[7] = {est.globalState_.rn_.x(), est.globalState_.rn_.y(), est.globalState_.rn_.z(), est.globalState_.qbn_.x(),
                        est.globalState_.qbn_.y(), est.globalState_.qbn_.z(), est.globalState_.qbn_.w()};
  rec->global_est.insert(rec->global_est.end(), ge, ge + 7);
  double gt[7] = {0, 0, 0, 0, 0, 0, 1};
  if (end_true) {
    Q4D qt = math_utils::R2Quat(end_true->R);
    const double v[7] = {end_true->p.x(), end_true->p.y(), end_true->p.z(), qt.x(), qt.y(), qt.z(), qt.w()};
    std::memcpy(gt, v, sizeof(v));
  }
  rec->global_true.insert(rec->global_true.end(), gt, gt + 7);
}

lins::fusion::EstimatorParams seq_params(const LidarModel& lm) {
  lins::fusion::EstimatorParams ep;
  ep.lidar = lm;
  ep.filter.init_ba = V3D(0, 0, 0);
  ep.filter.init_bw = V3D(0, 0, 0);
  return ep;
}

}  // namespace

extern "C" {

void* lins_seq_run(const lins_synth_cfg* cfg, uint64_t seed, int n_scans, int device) {
  SeqRecord* rec = new SeqRecord();
  SimDrive sim(cfg, seed);
  StateEstimator est(seq_params(sim.lm), device);
  ImageProjection ip(sim.lm);
  const bool verbose = std::getenv("LINS_SEQ_VERBOSE") != nullptr;
  Sweep sw;
  for (int k = 0; k < n_scans; ++k) {
    if (verbose) std::fprintf(stderr, "[lins_seq] simulating scan %d\n", k);
    sim.next(sw);
    std::vector<double> dts(SimDrive::nimu, sim.dt());
    std::vector<V3D> acc(sw.accs.begin() + 1, sw.accs.end()), gyr(sw.gyrs.begin() + 1, sw.gyrs.end());
    feed_scan(est, ip, rec, k, sw.t_end, sw.raw, dts, acc, gyr, &sw.start, &sw.end, verbose);
  }
  return rec;
}

// The same simulated drive written as a ROS1 bag (csrc/host/rosbag_reader.hpp): the raw sweeps on `lidar_topic`
// (sensor_msgs/PointCloud2, stamped at the sweep's end like LinsFusion uses them) and the IMU samples on `imu_topic`.
int lins_seq_write_bag(const lins_synth_cfg* cfg, uint64_t seed, int n_scans, const char* path, const char* lidar_topic, const char* imu_topic) {
  using namespace lins::rosbag;
  SimDrive sim(cfg, seed);
  Writer w;
  if (w.open(path) != LINS_BAG_OK) return LINS_BAG_E_IO;
  const uint32_t cl = w.add_connection(lidar_topic, "sensor_msgs/PointCloud2", "1158d486dd51d683ce2f1be655c3c181", "(sensor_msgs/PointCloud2)");
  const uint32_t ci = w.add_connection(imu_topic, "sensor_msgs/Imu", "6a62c6daae103f4ff57a132d6f95cec2", "(sensor_msgs/Imu)");
  Sweep sw;
  const double t0 = 1000.0;  // (bag times are seconds since the epoch; keep them small enough that 400 Hz stamps stay exact to the ns)
  uint32_t seq = 0;
  for (int k = 0; k < n_scans; ++k) {
    const double ts = t0 + sim.t;
    sim.next(sw);
    for (int i = 1; i <= SimDrive::nimu; ++i) {
      const double ti = ts + i * sim.dt();
      const double a[3] = {sw.accs[i].x(), sw.accs[i].y(), sw.accs[i].z()}, g[3] = {sw.gyrs[i].x(), sw.gyrs[i].y(), sw.gyrs[i].z()};
      w.write(ci, ti, Writer::encode_imu(seq++, ti, a, g));
    }
    w.write(cl, t0 + sw.t_end, Writer::encode_cloud_xyzi((uint32_t)k, t0 + sw.t_end, "velodyne", sw.raw));
  }
  return w.close();
}

// BASELINE.json configs[1] runner: replay a bag the way LinsFusion does (Estimator.cpp:123-284): IMU messages are buffered,
// every lidar message is a scan at its header stamp; between two scans the buffered IMU samples are propagated with
// dt = min(imu stamp, scan stamp) - estimator time (:230-236); the raw cloud goes through the restated image projection.
// Returns a record handle like lins_seq_run (status / pose per scan, one recorded unit per performIESKF call) or null.
void* lins_seq_run_bag(const char* path, const char* lidar_topic, const char* imu_topic, int max_scans, int lidar_model, int device, int* error) {
  using namespace lins::rosbag;
  if (error) *error = 0;
  Reader rd;
  int rc = rd.open(path);
  if (rc != LINS_BAG_OK) { if (error) *error = rc; return nullptr; }
  struct ImuS { double t; V3D a, g; };
  std::vector<ImuS> imus;
  std::vector<std::pair<double, Cloud>> scans;
  bool bad = false;
  rc = rd.for_each([&](const MessageView& m) {
    if (m.conn->topic == imu_topic) {
      ImuMsg im;
      if (!decode_imu(m.data, m.size, im)) { bad = true; return; }
      imus.push_back(ImuS{im.header.stamp, V3D(im.linear_acceleration[0], im.linear_acceleration[1], im.linear_acceleration[2]),
                          V3D(im.angular_velocity[0], im.angular_velocity[1], im.angular_velocity[2])});
    } else if (m.conn->topic == lidar_topic && (max_scans <= 0 || (int)scans.size() < max_scans)) {
      Header h;
      Cloud c;
      if (!decode_pointcloud2(m.data, m.size, h, c)) { bad = true; return; }
      scans.emplace_back(h.stamp, std::move(c));
    }
  });
  if (rc != LINS_BAG_OK || bad) { if (error) *error = rc != LINS_BAG_OK ? rc : LINS_BAG_E_FORMAT; return nullptr; }
  std::stable_sort(imus.begin(), imus.end(), [](const ImuS& x, const ImuS& y) { return x.t < y.t; });
  std::stable_sort(scans.begin(), scans.end(), [](const std::pair<double, Cloud>& x, const std::pair<double, Cloud>& y) { return x.first < y.first; });
  SeqRecord* rec = new SeqRecord();
  const LidarModel lm = lidar_model == 1 ? LidarModel::dense64() : LidarModel::vlp16();
  StateEstimator est(seq_params(lm), device);
  ImageProjection ip(lm);
  const bool verbose = std::getenv("LINS_SEQ_VERBOSE") != nullptr;
  size_t next_imu = 0;
  double est_time = scans.empty() ? 0.0 : scans.front().first - lm.scan_period;  // (the estimator's clock starts one sweep before the first scan)
  for (size_t k = 0; k < scans.size(); ++k) {
    const double ts = scans[k].first;
    std::vector<double> dts;
    std::vector<V3D> acc, gyr;
    while (est_time < ts && next_imu < imus.size()) {  // Estimator.cpp:228-236
      const ImuS& im = imus[next_imu];
      if (im.t <= est_time) { ++next_imu; continue; }   // upper_bound(estimator time)
      const double dt = std::min(im.t, ts) - est_time;
      dts.push_back(dt); acc.push_back(im.a); gyr.push_back(im.g);
      est_time += dt;
      if (im.t <= ts) ++next_imu;
    }
    est_time = ts;
    feed_scan(est, ip, rec, (int)k, ts, scans[k].second, dts, acc, gyr, nullptr, nullptr, verbose);
  }
  return rec;
}

void lins_seq_destroy(void* h) { delete static_cast<SeqRecord*>(h); }
int lins_seq_num_units(void* h) { return (int)static_cast<SeqRecord*>(h)->iters.size(); }
int lins_seq_num_scans(void* h) { return (int)static_cast<SeqRecord*>(h)->status.size(); }
void lins_seq_desc(void* h, lins_batch_desc* d) {
  SeqRecord* r = static_cast<SeqRecord*>(h);
  d->n_scans = (int32_t)r->iters.size();
  d->surf_flat = r->surfFlat.data(); d->surf_flat_off = r->offSF.data();
  d->corner_sharp = r->cornerSharp.data(); d->corner_sharp_off = r->offCS.data();
  d->surf_less_flat = r->surfLessFlat.data(); d->surf_less_flat_off = r->offSL.data();
  d->corner_less_sharp = r->cornerLessSharp.data(); d->corner_less_sharp_off = r->offCL.data();
  d->state_in = r->state_in.data(); d->cov_in = r->cov_in.data();
  d->point_format = LINS_POINTS_XYZI32;
}
const double* lins_seq_array(void* h, int which) {
  SeqRecord* r = static_cast<SeqRecord*>(h);
  switch (which) {
    case 0: return r->state_out.data();
    case 1: return r->rel_true.data();
    case 2: return r->global_est.data();
    case 3: return r->global_true.data();
  }
  return nullptr;
}
const int32_t* lins_seq_ints(void* h, int which) {
  SeqRecord* r = static_cast<SeqRecord*>(h);
  switch (which) {
    case 0: return r->iters.data();
    case 1: return r->flags.data();
    case 2: return r->scan_index.data();
    case 3: return r->status.data();
  }
  return nullptr;
}

}  // extern "C"
