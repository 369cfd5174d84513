// Batched bf16 x bf16 -> fp32 contraction on tcgen05 tensor cores (see gemm_tcgen05.cu).
//
//   C[b2][b1][m][n] = epilogue( sum_k A[b2][b1][m][k] * B[b2][b1][n][k] )
//
// Either operand may be stored K-major (k contiguous) or MN-major (m / n contiguous), so
// every product the matching path needs (X·Wᵀ, P·V, dSᵀ·K, dYᵀ·X ...) runs without a
// transposed copy in HBM.
#pragma once
#include "common.cuh"

namespace lmkd {

struct GemmOperand {
  const __nv_bfloat16* ptr = nullptr;
  int mn_major = 0;      // 0: [rows][K] (K contiguous)   1: [K][rows] (rows contiguous)
  int64_t ld = 0;        // pitch (elements) of the non-contiguous dimension
  int64_t stride_b1 = 0; // batch strides (elements)
  int64_t stride_b2 = 0;
};

enum EpiKind : int {
  EPI_STORE_F32 = 0,   // C(f32)  = alpha * acc * (rowv ? rowv[m] : 1)
  EPI_STORE_BF16 = 1,  // C(bf16) = same
  EPI_ACCUM_F32 = 2,   // C(f32) += same
  EPI_COSDIST = 3,     // C(f32)  = 1 - acc / (rowv[m] * colv[n] + eps)
  EPI_DIFF_SQ = 4,     // D = aux(bf16)[m][n] - acc ; C(bf16) = D ; rowred[m] += sum_n D^2
  EPI_AXPY_F32 = 5,    // C(f32)  = alpha * acc + rowv[m] * aux(f32)[m][n]
  EPI_LNRED_F32 = 6,   // C(f32)  = alpha * acc ; rowred[2m] += sum_n C*colv[n] ; rowred[2m+1] += sum_n C*(aux(bf16)[m][n]-colv2[n])
  EPI_BIAS_F32 = 7,    // C(f32)  = alpha * acc + colv[n]                      (Linear layer with bias)
  // softmax backward fused into the dP product (TRX attention): with p = aux(bf16)[m][n] * rowv[m],
  //   C2(bf16) = p   and   C(bf16) = p * (acc - rowv2[m])
  EPI_SMBWD_BF16 = 8,
  // squared Euclidean distance with a per-row arg-min (STRM DistanceLoss: cdist + min over the support tuples):
  //   d2 = max(rowv[m] + colv[n] - 2 acc, 0);  rowred (as uint64 [m]) = atomicMin( float_bits(d2) << 32 | n )
  // nothing is stored per element; columns to be ignored carry colv = +huge
  EPI_MINDIST = 9,
  EPI_BIAS_BF16 = 10,  // C(bf16) = act(alpha * acc + colv[n]), act = ReLU when `relu` is set (encoder Linear layers)
  EPI_LNRED_BF16 = 11, // EPI_LNRED_F32 with C stored as bf16 (the row reductions still see the fp32 accumulator)
  // Input-gradient scatter (backward of dropout(x + pe) over [episode][supports | queries][frame] rows): row m belongs
  // to group m / group_rows; its first split_rows rows go to C, the others to C2 (both dense [rows, N] fp32);
  // value = alpha * acc * dropout_scale(*seed, m * N + n) when drop_p > 0; `accumulate` adds to the buffers
  EPI_DXSCATTER = 12,
  EPI_AXPY_B16 = 13,   // C(f32)  = alpha * acc + rowv[m] * aux(bf16)[m][n]   (AXPY with a half-width aux operand)
};

struct GemmEpilogue {
  int kind = EPI_STORE_F32;
  float alpha = 1.f;
  float eps = 0.f;
  int relu = 0;                   // EPI_BIAS_BF16: clamp at zero
  int c_transposed = 0;           // EPI_STORE_F32 / EPI_STORE_BF16 only: element (m, n) goes to C[n * ldc + m] (lanes = consecutive m
                                  // write full lines), e.g. dV = (dO^T . P)^T with the long dimension d as tile rows
  int group_rows = 0, split_rows = 0;            // EPI_DXSCATTER
  const unsigned long long* seed = nullptr;      // EPI_DXSCATTER: device-resident dropout seed (read when drop_p > 0)
  float drop_p = 0.f;
  int accumulate = 0;
  void* C = nullptr;
  int64_t ldc = 0, c_b1 = 0, c_b2 = 0;
  const float* rowv = nullptr;
  int64_t rv_b1 = 0, rv_b2 = 0;
  const float* rowv2 = nullptr;   // second per-row vector (same batch strides as rowv)
  void* C2 = nullptr;             // second output (same layout and batch strides as C)
  const float* colv = nullptr;
  int64_t cv_b1 = 0, cv_b2 = 0;
  const float* colv2 = nullptr;   // second per-column vector (same batch strides as colv)
  const void* aux = nullptr;
  int64_t ldaux = 0, aux_b1 = 0, aux_b2 = 0;
  float* rowred = nullptr;
  int64_t rr_b1 = 0, rr_b2 = 0;
};

struct GemmDesc {
  int M = 0, N = 0, K = 0;
  int nb1 = 1, nb2 = 1;
  GemmOperand A, B;
  int block_n = 0;  // 0 = pick
  GemmEpilogue epi;
};

// returns 0 on success (error text via get_error())
int gemm_bf16(const GemmDesc& g, cudaStream_t stream);

// optional per-launch CUDA-event timing of the GEMM kernel (roofline measurement in bench.py)
void gemm_timing_enable(int on);
// synchronises on the recorded events; returns total kernel ms, true-shape FLOPs and launch count, then resets
int gemm_timing_read(double* ms, double* flops, int* launches);

}  // namespace lmkd
