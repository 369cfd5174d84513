// STRM DistanceLoss head (model/classifiers/strm_res18_sup.py:162-255; identical copies in strmclassifiers.py and
// strmclassifiers_res18.py): non-GEMM kernels.  The tuple MLP  relu(clsW . concat(frames of a tuple))  is factored
// into per-frame partial projections exactly like the TRX projection (trx.cu), so the [N, T, c*D] tuple tensor of
// :197-201 is never built.
#pragma once
#include "trx.cuh"

namespace lmkd {

// W fp32 [dm, card*D] -> Wcat bf16 [card, dm, D]
int strm_pack_weight(const float* W, __nv_bfloat16* Wcat, const TrxDims& s, cudaStream_t st);
// dWcat fp32 [card, dm, D] -> gW fp32 [dm, card*D]
int strm_unpack_wgrad(const float* dWcat, float* gW, const TrxDims& s, cudaStream_t st);

// P fp32 [M, card*dm] -> E = relu(sum_j P[(n, tau_j)][j] + bias) as bf16 rows, plus |E|^2 of the rounded rows.
// Queries: Eq [B, NqT, dm], nq2 [B, NqT]; supports class-sorted: Es [B, way, KTp, dm], ns2 [B, way, KTp]; rows of
// a class block past cnt*T are zero with ns2 = +huge (never the minimum).
int strm_tuple_relu_fwd(const float* P, const float* bias, const int* tuples, const int* slot, const int* cnt,
                        __nv_bfloat16* Eq, __nv_bfloat16* Es, float* nq2, float* ns2, const TrxDims& s, cudaStream_t st);

// best [B, way, NqT] (uint64: float_bits(min d^2) << 32 | column) -> logits [B, Nq, way] = -(1/T) sum_tau sqrt(d^2)
int strm_logits_fwd(const unsigned long long* best, const int* cnt, float* logits, const TrxDims& s, cudaStream_t st);

// gradient of the min-distance logits w.r.t. the embeddings: dEq [B, NqT, dm] (overwritten), dEs [B, way, KTp, dm]
// (atomic adds into a zeroed buffer: several query tuples may pick the same support tuple)
int strm_dist_bwd(const float* glogits, const unsigned long long* best, const int* cnt, const __nv_bfloat16* Eq,
                  const __nv_bfloat16* Es, float* dEq, float* dEs, const TrxDims& s, cudaStream_t st);

// ReLU backward + gather to frames: dPcat bf16 [M, card*dm]; gbias [dm] += column sums of the masked tuple gradients
int strm_relu_gather_bwd(const float* dEq, const float* dEs, const __nv_bfloat16* Eq, const __nv_bfloat16* Es,
                         const int* slot, const int* inv_off, const int* inv_idx, __nv_bfloat16* dPcat, float* gbias,
                         const TrxDims& s, cudaStream_t st);

}  // namespace lmkd
