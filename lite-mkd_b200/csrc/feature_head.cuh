// Student feature heads feeding the matching path, see feature_head.cu.
// Reference: model/backbone/resnet18_2fc.py:41-67 (fc1 / fc2), resnet18_student.py:38-58 (res18_2048).
#pragma once
#include "common.cuh"

namespace lmkd {

// pooled[r][c] = mean over the out_hw x out_hw adaptive-max-pool windows of fmap[r][c][H][W]
int frame_pool_fwd(const float* fmap, float* pooled, int64_t rows, int C, int H, int W, int out_hw, cudaStream_t st);
// grad_fmap[r][c][h][w] = grad_pooled[r][c] / out_hw^2 * (number of windows whose first maximum is (h, w))
int frame_pool_bwd(const float* fmap, const float* grad_pooled, float* grad_fmap, int64_t rows, int C, int H, int W,
                   int out_hw, cudaStream_t st);

// yb = bf16(y) and colsum[n] += sum_rows y[row][n]   (the bias gradient rides on the cast the GEMMs need)
int cast_colsum(const float* y, __nv_bfloat16* yb, float* colsum, int64_t rows, int cols, cudaStream_t st);

}  // namespace lmkd
