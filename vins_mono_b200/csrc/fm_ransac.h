#pragma once
#include <cstdint>
namespace vb {
// Inlier mask of cv::findFundamentalMat(m1, m2, FM_RANSAC, threshold, confidence, status) for `count`
// correspondences (x,y float pairs).  Returns false (status all zero) when no model was found.
bool fundamental_ransac_mask(const float* m1, const float* m2, int count, double threshold, double confidence,
                             uint8_t* status);
// The same call returning the model as well (row-major 3x3, the best RANSAC / LMedS hypothesis: OpenCV does not refit it).
bool fundamental_ransac(const float* m1, const float* m2, int count, double threshold, double confidence, uint8_t* status,
                        double* F);
}  // namespace vb
