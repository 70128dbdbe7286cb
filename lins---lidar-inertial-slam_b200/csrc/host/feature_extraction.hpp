// Host-side restatement of the four-stage LOAM feature extraction that runs in front of the IESKF update
// (reference: lins/include/StateEstimator.hpp  undistortPcl :619-654, calculateSmoothness :656-678,
// markOccludedPoints :680-713, extractFeatures :719-827) plus the pcl::VoxelGrid<PointXYZI> down-sampling
// it calls (:189, :822-825; leaf 0.2 m).  Sequential per ring-sextant (std::sort + greedy pick with neighbour
// suppression) so it stays on the CPU.  PRODUCT code.
//
// pcl::VoxelGrid is third-party (not in /root/reference): restated from PCL 1.7/1.8
// filters/impl/voxel_grid.hpp: voxel index from floor(p * inverse_leaf) - min_b, sort by voxel index, one
// centroid (all fields, f32 accumulation) per voxel, output in ascending voxel-index order.  std::sort's
// order inside a voxel is unspecified there; here it is the stable (input) order.
#ifndef LINS_HOST_FEATURE_EXTRACTION_HPP_
#define LINS_HOST_FEATURE_EXTRACTION_HPP_

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "cloud.hpp"
#include "math_utils.hpp"

namespace lins {

struct Smooth {  // StateEstimator.hpp:57-64
  double value = 0.0;
  size_t ind = 0;
};
struct byValue {  // StateEstimator.hpp:66-70
  bool operator()(Smooth const& left, Smooth const& right) const { return left.value < right.value; }
};

class VoxelGrid {
 public:
  void setLeafSize(float lx, float ly, float lz) {
    leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz;
    for (int i = 0; i < 3; ++i) inv_[i] = 1.0f / leaf_[i];
  }
  void filter(const Cloud& in, Cloud& out) const {
    out.clear();
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool any = false;
    for (const auto& p : in.points) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      any = true;
      mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
      mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    if (!any) return;
    int min_b[3], max_b[3], div_b[3];
    for (int i = 0; i < 3; ++i) {
      min_b[i] = static_cast<int>(std::floor(mn[i] * inv_[i]));
      max_b[i] = static_cast<int>(std::floor(mx[i] * inv_[i]));
      div_b[i] = max_b[i] - min_b[i] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    struct IdxPt { unsigned int idx; unsigned int pt; };
    std::vector<IdxPt> iv;
    iv.reserve(in.points.size());
    for (unsigned int k = 0; k < in.points.size(); ++k) {
      const auto& p = in.points[k];
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      int i0 = static_cast<int>(std::floor(p.x * inv_[0]) - static_cast<float>(min_b[0]));
      int i1 = static_cast<int>(std::floor(p.y * inv_[1]) - static_cast<float>(min_b[1]));
      int i2 = static_cast<int>(std::floor(p.z * inv_[2]) - static_cast<float>(min_b[2]));
      iv.push_back(IdxPt{static_cast<unsigned int>(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), k});
    }
    std::stable_sort(iv.begin(), iv.end(), [](const IdxPt& a, const IdxPt& b) { return a.idx < b.idx; });
    size_t first = 0;
    while (first < iv.size()) {
      size_t last = first + 1;
      while (last < iv.size() && iv[last].idx == iv[first].idx) ++last;
      float cx = 0, cy = 0, cz = 0, ci = 0;
      for (size_t k = first; k < last; ++k) {
        const auto& p = in.points[iv[k].pt];
        cx += p.x; cy += p.y; cz += p.z; ci += p.intensity;
      }
      float n = static_cast<float>(last - first);
      out.push_back(makePoint(cx / n, cy / n, cz / n, ci / n));
      first = last;
    }
  }

 private:
  float leaf_[3] = {0.2f, 0.2f, 0.2f}, inv_[3] = {5.f, 5.f, 5.f};
};

// Everything processPCL computes before the status switch (StateEstimator.hpp:283-289), per scan.
struct ScanFeatures {
  Cloud undistPointCloud;
  Cloud cornerPointsSharp, cornerPointsLessSharp, surfPointsFlat, surfPointsLessFlat;
};

struct FeatureParams {
  double edge_threshold = 0.5, surf_threshold = 0.5;  // exp_port.yaml:12-13
  double imu_lidar_extrinsic_angle = 0.0;             // exp_port.yaml:7
};

class FeatureExtractor {
 public:
  FeatureExtractor(const LidarModel& m = LidarModel(), const FeatureParams& fp = FeatureParams()) : lm(m), prm(fp) {
    downSizeFilter_.setLeafSize(0.2f, 0.2f, 0.2f);  // StateEstimator.hpp:189
  }

  void run(const Cloud& distPointCloud, const CloudInfo& segInfo, ScanFeatures& out) {
    const size_t cap = std::max<size_t>((size_t)lm.line_num * lm.scan_num, distPointCloud.size() + 16);
    cloudCurvature_.assign(cap, 0.0);
    cloudSmoothness_.assign(cap, Smooth());
    cloudNeighborPicked_.assign(cap, 0);
    cloudLabel_.assign(cap, 0);
    undistortPcl(distPointCloud, segInfo, out.undistPointCloud);
    calculateSmoothness(out.undistPointCloud, segInfo);
    markOccludedPoints(out.undistPointCloud, segInfo);
    extractFeatures(out.undistPointCloud, segInfo, out);
  }

 private:
  LidarModel lm;
  FeatureParams prm;
  VoxelGrid downSizeFilter_;
  std::vector<double> cloudCurvature_;
  std::vector<Smooth> cloudSmoothness_;
  std::vector<int> cloudNeighborPicked_, cloudLabel_;

  // StateEstimator.hpp:1104-1114 rotatePoint
  void rotatePoint(const PointType& pi, PointType& po) const {
    double y = math_utils::deg2rad(prm.imu_lidar_extrinsic_angle);
    double c = std::cos(y), s = std::sin(y);
    double px = pi.x, py = pi.y, pz = pi.z;
    po = pi;
    po.x = (float)(c * px - s * py); po.y = (float)(s * px + c * py); po.z = (float)pz;
  }

  // StateEstimator.hpp:619-654.  Re-stamps intensity = ring + SCAN_PERIOD * relTime.
  void undistortPcl(const Cloud& dist, const CloudInfo& segInfo, Cloud& undist) const {
    bool halfPassed = false;
    undist.clear();
    for (size_t i = 0; i < dist.points.size(); i++) {
      PointType point;
      rotatePoint(dist.points[i], point);
      double ori = -std::atan2(point.y, point.x);
      if (!halfPassed) {
        if (ori < segInfo.startOrientation - M_PI / 2) ori += 2 * M_PI;
        else if (ori > segInfo.startOrientation + M_PI * 3 / 2) ori -= 2 * M_PI;
        if (ori - segInfo.startOrientation > M_PI) halfPassed = true;
      } else {
        ori += 2 * M_PI;
        if (ori < segInfo.endOrientation - M_PI * 3 / 2) ori += 2 * M_PI;
        else if (ori > segInfo.endOrientation + M_PI / 2) ori -= 2 * M_PI;
      }
      double relTime = (ori - segInfo.startOrientation) / segInfo.orientationDiff;
      point.intensity = int(dist.points[i].intensity) + lm.scan_period * relTime;
      undist.push_back(point);
    }
  }

  // StateEstimator.hpp:656-678
  void calculateSmoothness(const Cloud& undist, const CloudInfo& segInfo) {
    int cloudSize = (int)undist.points.size();
    const auto& R = segInfo.segmentedCloudRange;
    for (int i = 5; i < cloudSize - 5; i++) {
      double diffRange = R[i - 5] + R[i - 4] + R[i - 3] + R[i - 2] + R[i - 1] - R[i] * 10 + R[i + 1] + R[i + 2] +
                         R[i + 3] + R[i + 4] + R[i + 5];
      cloudCurvature_[i] = diffRange * diffRange;
      cloudNeighborPicked_[i] = 0;
      cloudLabel_[i] = 0;
      cloudSmoothness_[i].value = cloudCurvature_[i];
      cloudSmoothness_[i].ind = i;
    }
  }

  // StateEstimator.hpp:680-713
  void markOccludedPoints(const Cloud& undist, const CloudInfo& segInfo) {
    int cloudSize = (int)undist.points.size();
    const auto& R = segInfo.segmentedCloudRange;
    const auto& C = segInfo.segmentedCloudColInd;
    for (int i = 5; i < cloudSize - 6; ++i) {
      float depth1 = R[i], depth2 = R[i + 1];
      int columnDiff = std::abs(int(C[i + 1] - C[i]));
      if (columnDiff < 10) {
        if (depth1 - depth2 > 0.3) {
          for (int k = -5; k <= 0; ++k) cloudNeighborPicked_[i + k] = 1;
        } else if (depth2 - depth1 > 0.3) {
          for (int k = 1; k <= 6; ++k) cloudNeighborPicked_[i + k] = 1;
        }
      }
      float diff1 = std::abs(R[i - 1] - R[i]);
      float diff2 = std::abs(R[i + 1] - R[i]);
      if (diff1 > 0.02 * R[i] && diff2 > 0.02 * R[i]) cloudNeighborPicked_[i] = 1;
    }
  }

  // The +-5 neighbour suppression of StateEstimator.hpp:764-777 / :796-811.  The reference indexes
  // ind+l unchecked; with the `ind = 0` default-entry quirk (cloudSmoothness_[4] of ring 0 is never written by
  // calculateSmoothness, so it sorts first with ind 0) that reads/writes index -1: undefined behaviour there,
  // bounds-checked here.
  void suppress(const CloudInfo& segInfo, int ind) {
    const auto& C = segInfo.segmentedCloudColInd;
    const int n = (int)cloudNeighborPicked_.size();
    for (int l = 1; l <= 5; l++) {
      if (ind + l >= n) break;
      int columnDiff = std::abs(int(C[ind + l] - C[ind + l - 1]));
      if (columnDiff > 10) break;
      cloudNeighborPicked_[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
      if (ind + l < 0) break;
      int columnDiff = std::abs(int(C[ind + l] - C[ind + l + 1]));
      if (columnDiff > 10) break;
      cloudNeighborPicked_[ind + l] = 1;
    }
  }

  // StateEstimator.hpp:719-827
  void extractFeatures(const Cloud& undist, const CloudInfo& segInfo, ScanFeatures& out) {
    out.cornerPointsSharp.clear(); out.cornerPointsLessSharp.clear();
    out.surfPointsFlat.clear(); out.surfPointsLessFlat.clear();
    Cloud surfPointsLessFlatScan, surfPointsLessFlatScanDS;
    for (int i = 0; i < lm.line_num; i++) {
      surfPointsLessFlatScan.clear();
      for (int j = 0; j < 6; j++) {
        int sp = (segInfo.startRingIndex[i] * (6 - j) + segInfo.endRingIndex[i] * j) / 6;
        int ep = (segInfo.startRingIndex[i] * (5 - j) + segInfo.endRingIndex[i] * (j + 1)) / 6 - 1;
        if (sp >= ep) continue;
        std::sort(cloudSmoothness_.begin() + sp, cloudSmoothness_.begin() + ep, byValue());

        int largestPickedNum = 0;
        for (int k = ep; k >= sp; k--) {
          int ind = (int)cloudSmoothness_[k].ind;
          if (cloudNeighborPicked_[ind] == 0 && cloudCurvature_[ind] > prm.edge_threshold &&
              segInfo.segmentedCloudGroundFlag[ind] == false) {
            largestPickedNum++;
            if (largestPickedNum <= 2) {
              cloudLabel_[ind] = 2;
              out.cornerPointsSharp.push_back(undist.points[ind]);
              out.cornerPointsLessSharp.push_back(undist.points[ind]);
            } else if (largestPickedNum <= 20) {
              cloudLabel_[ind] = 1;
              out.cornerPointsLessSharp.push_back(undist.points[ind]);
            } else {
              break;
            }
            cloudNeighborPicked_[ind] = 1;
            suppress(segInfo, ind);
          }
        }

        int smallestPickedNum = 0;
        for (int k = sp; k <= ep; k++) {
          int ind = (int)cloudSmoothness_[k].ind;
          if (cloudNeighborPicked_[ind] == 0 && cloudCurvature_[ind] < prm.surf_threshold &&
              segInfo.segmentedCloudGroundFlag[ind] == true) {
            cloudLabel_[ind] = -1;
            out.surfPointsFlat.push_back(undist.points[ind]);
            smallestPickedNum++;
            if (smallestPickedNum >= 4) break;
            cloudNeighborPicked_[ind] = 1;
            suppress(segInfo, ind);
          }
        }

        for (int k = sp; k <= ep; k++)
          if (cloudLabel_[k] <= 0) surfPointsLessFlatScan.push_back(undist.points[k]);
      }
      surfPointsLessFlatScanDS.clear();
      downSizeFilter_.filter(surfPointsLessFlatScan, surfPointsLessFlatScanDS);
      out.surfPointsLessFlat += surfPointsLessFlatScanDS;
    }
  }
};

}  // namespace lins
#endif
