// Host-side restatement of the upstream range-image projection / ground removal / BFS segmentation node
// (reference: lins/src/image_projection_node.cpp:191-415, LeGO-LOAM derived).  It stays on the CPU
// (BASELINE.json north_star: "image projection ... stay on CPU"); it exists here so synthetic raw scans can be
// turned into the /segmented_cloud + cloud_info + /outlier_cloud wire format the odometry path consumes
// (SURVEY.md §8 row F4).  No ROS, no OpenCV: cv::Mat -> flat std::vector.  PRODUCT code.
#ifndef LINS_HOST_IMAGE_PROJECTION_HPP_
#define LINS_HOST_IMAGE_PROJECTION_HPP_

#include <cfloat>
#include <cmath>
#include <limits>

#include "cloud.hpp"

namespace lins {

class ImageProjection {
 public:
  explicit ImageProjection(const LidarModel& m = LidarModel()) : lm(m) {
    const int n = lm.line_num * lm.scan_num;
    fullCloud.points.resize(n);
    allPushedIndX.resize(n); allPushedIndY.resize(n); queueIndX.resize(n); queueIndY.resize(n);
    // image_projection_node.cpp:128-140.  The reference stores the (-1,0) and (0,-1) offsets in a
    // std::pair<uint8_t,uint8_t> (image_projection_node.cpp:71), so they become +255: the "up" neighbour is
    // always out of range and the "left" neighbour is the pixel 255 columns to the right.  Kept as is.
    neighborIterator = {{255, 0}, {0, 1}, {0, 255}, {1, 0}};
    resetParameters();
  }

  // outputs (≙ the three published messages)
  Cloud segmentedCloud, outlierCloud;
  CloudInfo segMsg;

  // ≙ cloudHandler (image_projection_node.cpp:177-189) minus ROS I/O.  laserCloudIn: raw points in firing order.
  void process(const Cloud& laserCloudIn) {
    resetParameters();
    findStartEndAngle(laserCloudIn);
    projectPointCloud(laserCloudIn);
    groundRemoval();
    cloudSegmentation();
  }

 private:
  LidarModel lm;
  Cloud fullCloud;
  std::vector<float> rangeMat;
  std::vector<int8_t> groundMat;
  std::vector<int> labelMat;
  int labelCount = 1;
  std::vector<std::pair<int, int>> neighborIterator;
  std::vector<uint16_t> allPushedIndX, allPushedIndY, queueIndX, queueIndY;

  float& range(int r, int c) { return rangeMat[(size_t)r * lm.scan_num + c]; }
  int8_t& ground(int r, int c) { return groundMat[(size_t)r * lm.scan_num + c]; }
  int& label(int r, int c) { return labelMat[(size_t)r * lm.scan_num + c]; }

  void resetParameters() {  // image_projection_node.cpp:150-166
    const int n = lm.line_num * lm.scan_num;
    segmentedCloud.clear(); outlierCloud.clear();
    rangeMat.assign(n, FLT_MAX); groundMat.assign(n, 0); labelMat.assign(n, 0);
    labelCount = 1;
    PointType nanPoint = makePoint(std::numeric_limits<float>::quiet_NaN(), std::numeric_limits<float>::quiet_NaN(),
                                   std::numeric_limits<float>::quiet_NaN(), -1.f);
    std::fill(fullCloud.points.begin(), fullCloud.points.end(), nanPoint);
    segMsg.resize(lm.line_num, n);
  }

  void findStartEndAngle(const Cloud& in) {  // image_projection_node.cpp:191-203 (incl. the [size-2].x slip)
    const size_t n = in.points.size();
    if (n < 2) return;
    segMsg.startOrientation = -std::atan2(in.points[0].y, in.points[0].x);
    segMsg.endOrientation = -std::atan2(in.points[n - 1].y, in.points[n - 2].x) + 2 * M_PI;
    if (segMsg.endOrientation - segMsg.startOrientation > 3 * M_PI) {
      segMsg.endOrientation -= 2 * M_PI;
    } else if (segMsg.endOrientation - segMsg.startOrientation < M_PI) {
      segMsg.endOrientation += 2 * M_PI;
    }
    segMsg.orientationDiff = segMsg.endOrientation - segMsg.startOrientation;
  }

  void projectPointCloud(const Cloud& in) {  // image_projection_node.cpp:205-243
    for (size_t i = 0; i < in.points.size(); ++i) {
      PointType thisPoint = makePoint(in.points[i].x, in.points[i].y, in.points[i].z, 0.f);
      float verticalAngle = std::atan2(thisPoint.z, std::sqrt(thisPoint.x * thisPoint.x + thisPoint.y * thisPoint.y)) * 180 / M_PI;
      float rowF = (verticalAngle + lm.ang_bottom) / lm.ang_res_y;
      if (!(rowF >= 0) || rowF >= (float)lm.line_num) continue;
      int rowIdn = (int)rowF;
      float horizonAngle = std::atan2(thisPoint.x, thisPoint.y) * 180 / M_PI;
      double colD = -std::round((horizonAngle - 90.0) / lm.ang_res_x) + lm.scan_num / 2;
      if (colD < 0) continue;
      long columnIdn = (long)colD;
      if (columnIdn >= lm.scan_num) columnIdn -= lm.scan_num;
      if (columnIdn < 0 || columnIdn >= lm.scan_num) continue;
      float rng = std::sqrt(thisPoint.x * thisPoint.x + thisPoint.y * thisPoint.y + thisPoint.z * thisPoint.z);
      range(rowIdn, (int)columnIdn) = rng;
      thisPoint.intensity = (float)rowIdn + (float)columnIdn / 10000.0;
      fullCloud.points[columnIdn + (size_t)rowIdn * lm.scan_num] = thisPoint;
    }
  }

  void groundRemoval() {  // image_projection_node.cpp:245-291
    const float sensorMountAngle = 0.0f;
    for (int j = 0; j < lm.scan_num; ++j) {
      for (int i = 0; i < lm.ground_scan_ind; ++i) {
        size_t lowerInd = j + (size_t)i * lm.scan_num, upperInd = j + (size_t)(i + 1) * lm.scan_num;
        if (fullCloud.points[lowerInd].intensity == -1 || fullCloud.points[upperInd].intensity == -1) {
          ground(i, j) = -1;
          continue;
        }
        float diffX = fullCloud.points[upperInd].x - fullCloud.points[lowerInd].x;
        float diffY = fullCloud.points[upperInd].y - fullCloud.points[lowerInd].y;
        float diffZ = fullCloud.points[upperInd].z - fullCloud.points[lowerInd].z;
        float angle = std::atan2(diffZ, std::sqrt(diffX * diffX + diffY * diffY)) * 180 / M_PI;
        if (std::abs(angle - sensorMountAngle) <= 10) {
          ground(i, j) = 1;
          ground(i + 1, j) = 1;
        }
      }
    }
    for (int i = 0; i < lm.line_num; ++i)
      for (int j = 0; j < lm.scan_num; ++j)
        if (ground(i, j) == 1 || range(i, j) == FLT_MAX) label(i, j) = -1;
  }

  void cloudSegmentation() {  // image_projection_node.cpp:293-339
    for (int i = 0; i < lm.line_num; ++i)
      for (int j = 0; j < lm.scan_num; ++j)
        if (label(i, j) == 0) labelComponents(i, j);

    int sizeOfSegCloud = 0;
    for (int i = 0; i < lm.line_num; ++i) {
      segMsg.startRingIndex[i] = sizeOfSegCloud - 1 + 5;
      for (int j = 0; j < lm.scan_num; ++j) {
        if (label(i, j) > 0 || ground(i, j) == 1) {
          if (label(i, j) == 999999) {
            if (i > lm.ground_scan_ind && j % 5 == 0) {
              outlierCloud.push_back(fullCloud.points[j + (size_t)i * lm.scan_num]);
              continue;
            } else {
              continue;
            }
          }
          if (ground(i, j) == 1) {
            if (j % 5 != 0 && j > 5 && j < lm.scan_num - 5) continue;
          }
          segMsg.segmentedCloudGroundFlag[sizeOfSegCloud] = (ground(i, j) == 1);
          segMsg.segmentedCloudColInd[sizeOfSegCloud] = j;
          segMsg.segmentedCloudRange[sizeOfSegCloud] = range(i, j);
          segmentedCloud.push_back(fullCloud.points[j + (size_t)i * lm.scan_num]);
          ++sizeOfSegCloud;
        }
      }
      segMsg.endRingIndex[i] = sizeOfSegCloud - 1 - 5;
    }
  }

  void labelComponents(int row, int col) {  // image_projection_node.cpp:341-413
    const float segmentTheta = 1.0472f;
    const float segmentAlphaX = lm.ang_res_x / 180.0 * M_PI, segmentAlphaY = lm.ang_res_y / 180.0 * M_PI;
    const int segmentValidPointNum = 5, segmentValidLineNum = 3;
    std::vector<char> lineCountFlag(lm.line_num, 0);
    queueIndX[0] = row; queueIndY[0] = col;
    int queueSize = 1, queueStartInd = 0, queueEndInd = 1;
    allPushedIndX[0] = row; allPushedIndY[0] = col;
    int allPushedIndSize = 1;
    while (queueSize > 0) {
      int fromIndX = queueIndX[queueStartInd], fromIndY = queueIndY[queueStartInd];
      --queueSize; ++queueStartInd;
      label(fromIndX, fromIndY) = labelCount;
      for (const auto& it : neighborIterator) {
        int thisIndX = fromIndX + it.first, thisIndY = fromIndY + it.second;
        if (thisIndX < 0 || thisIndX >= lm.line_num) continue;
        if (thisIndY < 0) thisIndY = lm.scan_num - 1;
        if (thisIndY >= lm.scan_num) thisIndY = 0;
        if (label(thisIndX, thisIndY) != 0) continue;
        float d1 = std::max(range(fromIndX, fromIndY), range(thisIndX, thisIndY));
        float d2 = std::min(range(fromIndX, fromIndY), range(thisIndX, thisIndY));
        float alpha = it.first == 0 ? segmentAlphaX : segmentAlphaY;
        float angle = std::atan2(d2 * std::sin(alpha), (d1 - d2 * std::cos(alpha)));
        if (angle > segmentTheta) {
          queueIndX[queueEndInd] = thisIndX; queueIndY[queueEndInd] = thisIndY;
          ++queueSize; ++queueEndInd;
          label(thisIndX, thisIndY) = labelCount;
          lineCountFlag[thisIndX] = 1;
          allPushedIndX[allPushedIndSize] = thisIndX; allPushedIndY[allPushedIndSize] = thisIndY;
          ++allPushedIndSize;
        }
      }
    }
    bool feasibleSegment = false;
    if (allPushedIndSize >= 30) {
      feasibleSegment = true;
    } else if (allPushedIndSize >= segmentValidPointNum) {
      int lineCount = 0;
      for (int i = 0; i < lm.line_num; ++i) if (lineCountFlag[i]) ++lineCount;
      if (lineCount >= segmentValidLineNum) feasibleSegment = true;
    }
    if (feasibleSegment) {
      ++labelCount;
    } else {
      for (int i = 0; i < allPushedIndSize; ++i) label(allPushedIndX[i], allPushedIndY[i]) = 999999;
    }
  }
};

}  // namespace lins
#endif
