// OTAM kernels (see otam.cu)
#pragma once
#include "common.cuh"

namespace lmkd {

// dist [B, Nq*L, ld] fp32 (frame distances) -> pair [B, Nq, Ns] (both directions summed)
// single_dir = 1 runs only the query->support direction (raw OTAM_cum_dist)
int otam_dp_fwd(const float* dist, float* pair, int B, int Nq, int Ns, int L, int M, int64_t ld, float lbda,
                int single_dir, cudaStream_t stream);
// gpair [B, Nq, Ns] -> dnum [B, Nq*L, ld] bf16 (= dL/d<x,y>), gnq [B*Nq*L] += dL/d|x|, gns += dL/d|y|
// if ddist_raw != NULL only d loss / d dist [.., ld] is written (no cosine chain)
int otam_dp_bwd(const float* dist, const float* gpair, const float* nq, const float* ns, __nv_bfloat16* dnum,
                float* gnq, float* gns, float* ddist_raw, int B, int Nq, int Ns, int L, int M, int64_t ld, float lbda,
                float eps, int single_dir, cudaStream_t stream);
int otam_class_fwd(const float* pair, const float* labels, const int* nanflag, float* probs, int B, int Nq, int Ns,
                   int way, int* status, cudaStream_t stream);
int otam_class_bwd(const float* gprobs, const float* probs, const float* labels, const int* nanflag, float* gpair,
                   int B, int Nq, int Ns, int way, cudaStream_t stream);
// zero gq[b] (nq floats) and gs[b] (ns floats) of every episode whose nanflag is set (backward of the NaN guard)
int otam_zero_flagged(const int* nanflag, float* gq, float* gs, int B, int64_t nq, int64_t ns, cudaStream_t stream);
// out = den > 0 ? num / den : 0
int div_safe(const float* num, const float* den, float* out, int64_t n, cudaStream_t stream);

}  // namespace lmkd
