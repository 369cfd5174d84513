// Device-resident teacher-feature store, see feature_store.cu.
// Reference format: one [1, L, 2048] fp32 `feature.npy` per video (writer teacher/code/extract_multi_feature.py:
// 113-121), read back video by video and concatenated per episode (video_reader.py:388-395, 470-471).
#pragma once
#include "common.cuh"

namespace lmkd {

// out[i][:] = float(store[index[i]][:]) for i < count; rows of row_elems elements (multiple of 8);
// an index outside [0, store_rows) ORs 4 into *status and yields zeros
int episode_gather(const void* store, int store_bf16, int64_t store_rows, const int64_t* index, int64_t count,
                   int64_t row_elems, float* out, int* status, cudaStream_t st);

// fused feature MSE whose teacher operand is read in place from the store:
//   partial sums of (s - t)^2 per block, ds = gscale * (s - t), t = float(store[index[row]])
int feat_mse_store_fwdbwd(const float* s, const void* store, int store_bf16, int64_t store_rows, const int64_t* index,
                          int64_t count, int64_t row_elems, float* ds, float gscale, float* partials, int max_partials,
                          int* npartials, int* status, cudaStream_t st);

}  // namespace lmkd
