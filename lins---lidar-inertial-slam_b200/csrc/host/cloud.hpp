// Host-side point-cloud containers at the boundary.  PCL is not available in this image; these are layout
// compatible stand-ins: lins_point == pcl::PointXYZI (32 B) and CloudInfo == cloud_msgs::cloud_info
// (reference: cloud_msgs/msg/cloud_info.msg:1-12).  PRODUCT code.
#ifndef LINS_HOST_CLOUD_HPP_
#define LINS_HOST_CLOUD_HPP_

#include <cstdint>
#include <memory>
#include <vector>

#include "../../../include/lins_gpu.h"

namespace lins {

typedef lins_point PointType;  // parameters.h:52
static_assert(sizeof(PointType) == 32, "pcl::PointXYZI layout");

inline PointType makePoint(float x, float y, float z, float intensity) {
  PointType p;
  p.x = x; p.y = y; p.z = z; p.pad0 = 1.0f;  // PCL_ADD_POINT4D sets data[3] = 1
  p.intensity = intensity; p.pad1 = p.pad2 = p.pad3 = 0.f;
  return p;
}

struct Cloud {
  std::vector<PointType> points;
  void clear() { points.clear(); }
  size_t size() const { return points.size(); }
  void push_back(const PointType& p) { points.push_back(p); }
  Cloud& operator+=(const Cloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); return *this; }
  typedef std::shared_ptr<Cloud> Ptr;
};

// cloud_msgs::cloud_info
struct CloudInfo {
  std::vector<int32_t> startRingIndex, endRingIndex;
  float startOrientation = 0, endOrientation = 0, orientationDiff = 0;
  std::vector<uint8_t> segmentedCloudGroundFlag;
  std::vector<uint32_t> segmentedCloudColInd;
  std::vector<float> segmentedCloudRange;
  void resize(int line_num, int n) {
    startRingIndex.assign(line_num, 0); endRingIndex.assign(line_num, 0);
    segmentedCloudGroundFlag.assign(n, 0); segmentedCloudColInd.assign(n, 0); segmentedCloudRange.assign(n, 0);
  }
};

// Lidar geometry: the reference hard-wires VLP-16 (parameters.h:82-84, StateEstimator.hpp:54-55); it is a
// runtime struct here so the 64x1024 stress shape (BASELINE.json configs[3]) can be generated.
struct LidarModel {
  int line_num = 16;         // LINE_NUM
  int scan_num = 1800;       // SCAN_NUM
  float ang_res_x = 0.2f;    // deg / column
  float ang_res_y = 2.0f;    // deg / ring
  float ang_bottom = 15.0f + 0.1f;
  int ground_scan_ind = 5;   // groundScanInd
  double scan_period = 0.1;  // SCAN_PERIOD
  static LidarModel vlp16() { return LidarModel(); }
  static LidarModel dense64() {
    LidarModel m;
    m.line_num = 64; m.scan_num = 1024; m.ang_res_x = 360.0f / 1024.0f; m.ang_res_y = 45.0f / 63.0f;
    m.ang_bottom = 22.5f + 0.1f; m.ground_scan_ind = 24;
    return m;
  }
};

}  // namespace lins
#endif
