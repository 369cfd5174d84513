// Feature preparation kernels: fp32 -> bf16 casts fused with what the consumer needs
// (row L2 norms for the cosine similarity; positional encoding + dropout for TRX).
#pragma once
#include "common.cuh"

namespace lmkd {

// counter-based keep mask shared by forward, backward and the test hook.  One 64-bit hash covers the four
// consecutive elements of a 16-byte vector (a 16-bit uniform each, so p is honoured to 1.5e-5); element
// idx keeps its value, scaled by 1/(1-p), iff its uniform is >= p.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t group) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (group + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  return static_cast<uint32_t>(p * 65536.0f + 0.999f);
}
// scales of elements 4*group .. 4*group+3
__host__ __device__ __forceinline__ void dropout_scale4(uint64_t seed, uint64_t group, uint32_t thr, float inv_keep,
                                                        float (&sc)[4]) {
  const uint64_t z = mix64(seed, group);
  sc[0] = (static_cast<uint32_t>(z) & 0xFFFFu) >= thr ? inv_keep : 0.f;
  sc[1] = (static_cast<uint32_t>(z >> 16) & 0xFFFFu) >= thr ? inv_keep : 0.f;
  sc[2] = (static_cast<uint32_t>(z >> 32) & 0xFFFFu) >= thr ? inv_keep : 0.f;
  sc[3] = (static_cast<uint32_t>(z >> 48) & 0xFFFFu) >= thr ? inv_keep : 0.f;
}
__host__ __device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  const uint64_t z = mix64(seed, idx >> 2);
  return (static_cast<uint32_t>(z >> (16 * (idx & 3))) & 0xFFFFu) >= dropout_threshold(p) ? inv_keep : 0.f;
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
// bf16 -> fp32 (features staged from the host in bf16); n a multiple of 8
int upcast_bf16(const __nv_bfloat16* x, float* y, int64_t n, cudaStream_t stream);

int dropout_mask(float* out, int64_t n, float p, uint64_t seed, cudaStream_t stream);

}  // namespace lmkd
