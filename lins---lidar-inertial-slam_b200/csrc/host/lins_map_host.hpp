// lins_map_host.hpp — host side of row F2: the 6x6 step of LMOptimization after matAtA / matAtB exist
// (lins/src/lidar_mapping_node.cpp:1598-1632) and the per-iteration sin / cos of transformTobeMapped (:579-592,
// :1527-1532).  PRODUCT code; all f32 like the reference (cv::Mat CV_32F), small kernels from lins_cv_small.hpp.
#ifndef LINS_HOST_MAP_HOST_HPP_
#define LINS_HOST_MAP_HOST_HPP_

#include <cmath>
#include <cstring>

#include "lins_cv_small.hpp"

namespace lins {
namespace mapping {

struct LmState {
  bool isDegenerate = false;  // survives the iterations of one scan2MapOptimization call (:1606-1620)
  float matP[36];
};

// matX = solve(matAtA, matAtB); degeneracy handling of iteration 0; transformTobeMapped += matX; returns converged.
inline bool lm_step(const float* AtA, const float* AtB, int iterCount, float* T, LmState& st, float& deltaR, float& deltaT) {
  float Aw[36], X[6];
  std::memcpy(Aw, AtA, sizeof(Aw));
  std::memcpy(X, AtB, sizeof(X));
  if (!lins_cv::qr_solve<6, 6>(Aw, X)) std::memset(X, 0, sizeof(X));  // a failed cv::solve zeroes matX
  if (iterCount == 0) {
    float Ae[36], E[6], V[36], V2[36], Vc[36], Vinv[36];
    std::memcpy(Ae, AtA, sizeof(Ae));
    lins_cv::jacobi_eigen<6>(Ae, E, V);
    std::memcpy(V2, V, sizeof(V2));
    st.isDegenerate = false;
    const float eignThre[6] = {100, 100, 100, 100, 100, 100};
    for (int i = 5; i >= 0; --i) {
      if (E[i] < eignThre[i]) { for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0; st.isDegenerate = true; }
      else break;
    }
    std::memcpy(Vc, V, sizeof(Vc));
    if (!lins_cv::lu_invert<6>(Vc, Vinv)) std::memset(Vinv, 0, sizeof(Vinv));
    lins_cv::gemm<6, 6, 6>(Vinv, V2, st.matP);
  }
  if (st.isDegenerate) {
    float X2[6];
    std::memcpy(X2, X, sizeof(X2));
    lins_cv::gemm<6, 6, 1>(st.matP, X2, X);
  }
  for (int i = 0; i < 6; ++i) T[i] += X[i];
  auto rad2deg = [](float a) { return a * 57.29578f; };  // pcl::rad2deg(float)
  deltaR = (float)std::sqrt(std::pow((double)rad2deg(X[0]), 2) + std::pow((double)rad2deg(X[1]), 2) + std::pow((double)rad2deg(X[2]), 2));
  deltaT = (float)std::sqrt(std::pow((double)(X[3] * 100), 2) + std::pow((double)(X[4] * 100), 2) + std::pow((double)(X[5] * 100), 2));
  return deltaR < 0.05 && deltaT < 0.05;
}

}  // namespace mapping
}  // namespace lins
#endif
