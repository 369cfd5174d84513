// Feature preparation kernels: fp32 -> bf16 casts fused with what the consumer needs
// (row L2 norms for the cosine similarity; positional encoding + dropout for TRX).
#pragma once
#include "common.cuh"

namespace lmkd {

// counter-based keep mask shared by forward, backward and the test hook:
// returns the dropout scale (0 or 1/(1-p)) of element `idx`
__host__ __device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<uint32_t>(z >> 32);
}
__host__ __device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  const float u = (mix32(seed, idx) >> 8) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}

// x[rows, D] fp32 -> xb[rows, D] bf16 and norms[rows] = |x|_2 (fp32); nanflag[row / rows_per_flag] |= isnan
int feat_cast_norm(const float* x, __nv_bfloat16* xb, float* norms, int* nanflag, int64_t rows, int D,
                   int64_t rows_per_flag, cudaStream_t stream);

// TRX input: out[(b, n, l), :] = bf16( dropout( x[(b, n, l), :] + pe[l, :] ) ), videos of an
// episode laid out supports first then queries
// effective seed = seed + (seed_dev ? *seed_dev : 0); it is written to *seed_used (device) for the backward
int trx_pe_cast(const float* support, const float* query, const float* pe, __nv_bfloat16* out, int B,
                int Ns, int Nq, int L, int D, float p, uint64_t seed, const uint64_t* seed_dev,
                uint64_t* seed_used, cudaStream_t stream);

// grad of the above: gs/gq (+)= dropout_scale * dx
int trx_dx_scatter(const float* dx, float* gsupport, float* gquery, int B, int Ns, int Nq, int L, int D,
                   float p, const uint64_t* seed_used, int accumulate, cudaStream_t stream);

// plain fp32 -> bf16 (weights)
int cast_bf16(const float* x, __nv_bfloat16* y, int64_t n, cudaStream_t stream);

int dropout_mask(float* out, int64_t n, float p, uint64_t seed, cudaStream_t stream);

}  // namespace lmkd
