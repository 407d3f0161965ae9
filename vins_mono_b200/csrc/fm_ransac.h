#pragma once
#include <cstdint>
namespace vb {
// Inlier mask of cv::findFundamentalMat(m1, m2, FM_RANSAC, threshold, confidence, status) for `count`
// correspondences (x,y float pairs).  Returns false (status all zero) when no model was found.
bool fundamental_ransac_mask(const float* m1, const float* m2, int count, double threshold, double confidence,
                             uint8_t* status);
}  // namespace vb
