// Internal launch interface of the front-end kernels (fe_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace vb {

constexpr int MAX_PYR_LEVELS = 5;
constexpr int SORT_SMEM_KEYS = 16384;  // 128 KB of 64-bit candidate keys sorted inside one CTA's shared memory

struct PyramidView {
    const uint8_t* img[MAX_PYR_LEVELS];
    int rows[MAX_PYR_LEVELS], cols[MAX_PYR_LEVELS], pitch[MAX_PYR_LEVELS];
    int nlev;  // index of the coarsest level (levels 0..nlev)
};

void launch_clahe(const uint8_t* src, int rows, int cols, int spitch, uint8_t* lut, uint8_t* dst, int dpitch,
                  cudaStream_t s);
void launch_pyrdown(const uint8_t* src, int rows, int cols, int spitch, uint8_t* dst, int dpitch, cudaStream_t s);
void launch_lk(const PyramidView& prev, const PyramidView& next, const float* prev_pts, int n, float* next_pts,
               uint8_t* status, cudaStream_t s);
void launch_mask_discs(uint8_t* mask, int rows, int cols, int pitch, const int* centres, int n, int radius,
                       const int* halfw, cudaStream_t s);
void launch_min_eig(const uint8_t* img, int rows, int cols, int pitch, const uint8_t* mask, int mpitch, float* eig,
                    int epitch, unsigned* max_sortable, cudaStream_t s);
void launch_gftt_tail(const float* eig, int rows, int cols, int epitch, const uint8_t* mask, int mpitch,
                      const unsigned* max_sortable, double quality, unsigned long long* keys, int capacity, int* count,
                      int max_corners, float min_dist, int* cell_cnt, short2* cell_pts, float* out_pts, int* out_n,
                      cudaStream_t s);

}  // namespace vb
