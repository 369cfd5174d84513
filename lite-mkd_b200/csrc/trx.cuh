// TRX (temporal-relational cross-transformer) non-GEMM kernels; see trx.cu.
// Reference semantics: model/classifiers/TRX.py:75-152, teacher/code/model.py:253-346.
#pragma once
#include "common.cuh"

namespace lmkd {

struct TrxDims {
  int B, Ns, Nq, L, D, d, card, way, shot;
  int T;        // C(L, card)
  int N;        // Ns + Nq
  int KT;       // shot * T   (columns of one class group)
  int KTp;      // KT rounded up to 16 (pitch of one class group)
  int NqT;      // Nq * T
  int NqT_full; // Nq*T of the whole episode when this describes a chunk of its queries (== NqT otherwise)
  int m_off;    // first query-tuple row of the chunk inside the episode (0 otherwise)
  int64_t M;    // B * N * L  (frame rows)
  int64_t R;    // B * N * T  (tuple rows)
};

// slot[b][n] = class * shot + rank-within-class (or -1), cnt[b][c] = supports of class c
int trx_class_slots(const float* labels, int* slot, int* cnt, int* status, const TrxDims& s, cudaStream_t st);

// zero rows [cnt*T, KTp) of every (b, class) block of Ks / Vs (padding and missing shots)
int trx_zero_pad_rows(const int* cnt, __nv_bfloat16* Ks, __nv_bfloat16* Vs, const TrxDims& s, cudaStream_t st);

// P fp32 [M, 2*card*d] (per-frame partial projections) -> normalised keys / raw values, bf16.
// Queries: Kq/Vq [B, NqT, d]; supports (class-sorted, padded): Ks/Vs [B, way, KTp, d].
// stats [R, 2] = (mean, rstd) per tuple row (row id = (b*N + n)*T + tau).
int trx_tuple_ln_fwd(const float* P, const float* bk, const float* bv, const float* gamma, const float* beta,
                     const int* tuples, const int* slot, __nv_bfloat16* Kq, __nv_bfloat16* Vq, __nv_bfloat16* Ks,
                     __nv_bfloat16* Vs, float* stats, float ln_eps, int* scratch_flag /* one device int, may be null */,
                     const TrxDims& s, cudaStream_t st);

// S fp32 [B, NqT, way*KTp] (already scaled by 1/sqrt(d)) -> Patt bf16, softmax within each class group
int trx_softmax_fwd(const float* S, const int* cnt, __nv_bfloat16* Patt, const TrxDims& s, cudaStream_t st);

// rowred [B, way, NqT] (sum_n diff^2 per tuple row) -> logits [B, Nq, way] = -(1/T) sum_tau
int trx_logits_fwd(const float* rowred, const int* cnt, float* logits, const TrxDims& s, cudaStream_t st);

// srow[b][c][m] = 2 g[b][q(m)][c] / T;  with linv / rs non-null also rs = srow * linv (row scale of the
// un-normalised probabilities the fused attention kernel keeps)
int trx_attn_bwd_prep(const float* glogits, const int* cnt, float* srow, const float* linv, float* rs,
                      const TrxDims& s, cudaStream_t st);

// dS = Patt * (dP - sum_group(Patt * dP)) and Ps = Patt * srow   (both bf16)
int trx_softmax_bwd(const __nv_bfloat16* Patt, const float* dP, const int* cnt, const float* srow, __nv_bfloat16* dS,
                    __nv_bfloat16* Ps, const TrxDims& s, cudaStream_t st);

// LayerNorm backward per tuple row; writes dxk/dxv [R, d] fp32 and per-block partials of
// (dgamma, dbeta, dbk, dbv) into partials [nblocks, 4, d]; returns nblocks through *nblocks_out
int trx_ln_bwd(const float* P, const float* bk, const float* gamma, const float* stats, const int* tuples,
               const int* slot, const float* dKq, const float* dKs, const float* dVs, const float* srow,
               const __nv_bfloat16* Dq, float* dxk, float* dxv, float* partials, int max_blocks, int* nblocks_out,
               const TrxDims& s, cudaStream_t st);
// accumulate != 0: the four parameter gradients are += (gradient accumulation straight into .grad)
int trx_reduce_partials(const float* partials, int nblocks, float* ggamma, float* gbeta, float* gbk, float* gbv,
                        int d, int accumulate, cudaStream_t st);

// fused LayerNorm-backward + gather (no dxk/dxv round trip); usable when trx_bwd_fused_fits()
bool trx_bwd_fused_fits(const TrxDims& s);
// lnred_q [B*NqT, 2], lnred_s [B*way*KTp, 2]: per tuple row (sum_i dK*gamma, sum_i dK*(K^ - beta)),
// produced by the EPI_LNRED epilogue of the dK GEMMs; grad_rows_bf16 != 0: dKq / dKs / dVs point at bf16 rows
int trx_ln_gather_bwd_fused(const float* P, const float* bk, const float* gamma, const float* stats,
                            const int* tuples, const int* slot, const float* dKq, const float* dKs, const float* dVs,
                            int grad_rows_bf16, const float* lnred_q, const float* lnred_s, const float* srow,
                            const __nv_bfloat16* Dq,
                            __nv_bfloat16* dPcat, float* partials, int max_blocks, int* nblocks_out,
                            const TrxDims& s, cudaStream_t st);

// dPcat bf16 [M, 2*card*d]: column block (which, j) of frame row (b, n, l) = sum of dx{k,v} over
// tuples whose j-th frame is l (inverse lists inv_off [card*L + 1], inv_idx [card*T])
int trx_tuple_gather_bwd(const float* dxk, const float* dxv, const int* inv_off, const int* inv_idx,
                         __nv_bfloat16* dPcat, const TrxDims& s, cudaStream_t st);

// ---- TRX_sup: cosine similarity between the per-class query prototypes (TRX_sup.py:114-164) ----
// O_c[q] = v_q[q] - diff_c[q]  ([T, d] flattened); sim[b][q][i][j] = <O_i, O_j> / (|O_i| |O_j|)
// gram [B, Nq, way, way] keeps <O_i, O_j> for the backward.
int trx_proto_sim_fwd(const __nv_bfloat16* Vq, const __nv_bfloat16* Dq, const int* cnt, float* gram, float* sim,
                      const TrxDims& s, cudaStream_t st);
// E_c = srow_c * diff_c + sum_j a_cj * O_j  (total gradient w.r.t. prototype O_c), bf16 [B, way, NqT, d];
// a is derived from d loss / d sim (gsim) and the stored Gram matrix
int trx_proto_sim_bwd(const __nv_bfloat16* Vq, const __nv_bfloat16* Dq, const int* cnt, const float* gram,
                      const float* gsim, const float* srow, __nv_bfloat16* E, const TrxDims& s, cudaStream_t st);

// Wk/Wv fp32 [d, card*D] -> Wcat bf16 [2, card, d, D]
int trx_pack_weights(const float* Wk, const float* Wv, __nv_bfloat16* Wcat, const TrxDims& s, cudaStream_t st);
// dWcat fp32 [2, card, d, D] -> gWk, gWv fp32 [d, card*D]
int trx_unpack_wgrad(const float* dWcat, float* gWk, float* gWv, const TrxDims& s, int accumulate, cudaStream_t st);

}  // namespace lmkd
