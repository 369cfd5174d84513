/* lmkd — C ABI of the B200-native episodic matching + D2M distillation path.
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json `north_star`.  The
 * reference (HuiGuanLab/Lite-MKD) is pure Python/PyTorch, so "what its FFI for this path would
 * bind" is the set of tensor-level operations its classifier / Distiller modules perform; each
 * entry point below cites the reference interface it replaces (paths relative to the reference
 * repository root).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless stated otherwise;
 *   - tensors are dense row-major with the shapes given; fp32 unless stated;
 *   - the caller owns ALL memory (inputs, outputs, workspaces); the library never allocates or
 *     frees device memory and keeps no pointer after a call returns;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value 0 = success, non-zero = failure with a message from lmkd_last_error();
 *   - `status` (optional, may be NULL) is a device int the kernels OR error bits into
 *     (1 = label outside [0, way), 2 = more than `shot` supports in one class); the caller reads
 *     it at its next natural synchronisation point;
 *   - workspaces are opaque byte buffers sized by the matching *_workspace_bytes(); the forward
 *     call fills it, the backward call of the same shape reads it.
 */
#ifndef LMKD_H_
#define LMKD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* lmkd_last_error(void);
int lmkd_version(void);

/* ---- frame similarity (teacher/code/model.py:3260-3269 cos_sim) ---------------------------
 * dist[b][i][j] = 1 - <x_i, y_j> / (|x_i| |y_j| + eps); x = query frames [B, nx, D],
 * y = support frames [B, ny, D]; bf16 tcgen05 contraction, fp32 norms and epilogue.
 * ld = lmkd_sim_pitch(ny) is the row pitch of dist. */
int64_t lmkd_sim_pitch(int64_t ny);
size_t lmkd_sim_workspace_bytes(int B, int nx, int ny, int D);
int lmkd_sim_fwd(const float* x, const float* y, int B, int nx, int ny, int D, float eps, float* dist,
                 void* workspace, void* stream);

/* ---- OTAM head (teacher/code/model.py:3271-3343 OTAM_cum_dist + CNN_OTAM.forward) ----------
 * support [B, Ns, L, D], labels [B, Ns] (float class ids, as the reference's data loader
 * produces), query [B, Nq, L, D] -> probs [B, Nq, way] = softmax over classes of the negated
 * class-mean bidirectional soft-DTW distance.  lambda = 0.1 and eps = 0.01 in the reference. */
size_t lmkd_otam_workspace_bytes(int B, int Ns, int Nq, int L, int D, int way);
int lmkd_otam_fwd(const float* support, const float* labels, const float* query, int B, int Ns, int Nq, int L,
                  int D, int way, float lambda, float eps, float* probs, float* pair_dists /* [B,Nq,Ns] or NULL */,
                  void* workspace, int* status, void* stream);
int lmkd_otam_bwd(const float* grad_probs, const float* probs, const float* support, const float* labels,
                  const float* query, int B, int Ns, int Nq, int L, int D, int way, float lambda, float eps,
                  float* grad_support, float* grad_query, void* workspace, void* stream);
/* OTAM_cum_dist alone (teacher/code/model.py:3271-3299), one direction, on a given distance
 * tensor dists [P, L, M] -> out [P]; if grad_out [P] and grad_dists [P, L, M] are non-NULL the
 * backward is run too.  Used by the parity tests against the reference's own outputs. */
int lmkd_otam_cum_dist(const float* dists, int64_t P, int L, int M, float lambda, float* out,
                       const float* grad_out, float* grad_dists, void* stream);

/* ---- TRX head, one cardinality (model/classifiers/TRX.py:51-164 TemporalCrossTransformer;
 *      generic D / cardinality: teacher/code/model.py:226-361) -------------------------------- */
typedef struct {
  int B, Ns, Nq, L, D;     /* episodes, supports, queries, frames, feature dim               */
  int d;                   /* trans_linear_out_dim (options.py:22, 1152)                     */
  int card;                /* temporal_set_size (2 or 3)                                     */
  int way, shot;           /* classes; max supports per class                                */
  float dropout_p;         /* PositionalEncoding dropout (TRX.py:28,49); 0 in eval()         */
  uint64_t seed;           /* dropout stream (host part)                                     */
  const uint64_t* seed_dev;/* optional DEVICE counter added to `seed` when the forward runs  */
                           /* (lets a captured CUDA graph draw a new mask on every replay);  */
                           /* the forward records the effective seed in the workspace and    */
                           /* the backward reads it from there                               */
  float ln_eps;            /* LayerNorm eps (1e-5)                                           */
} lmkd_trx_shape;

size_t lmkd_trx_workspace_bytes(const lmkd_trx_shape* s, int need_grad);
/* tuples [T, card] int32 (lexicographic combinations, TRX.py:70-73); pe [L, D] fp32;
 * Wk, Wv [d, card*D]; bk, bv, gamma, beta [d]; logits out [B, Nq, way] (fp32, ON DEVICE — the
 * reference allocates them on the CPU, TRX.py:118). */
int lmkd_trx_fwd(const lmkd_trx_shape* s, const float* support, const float* labels, const float* query,
                 const float* pe, const int32_t* tuples, const float* Wk, const float* bk, const float* Wv,
                 const float* bv, const float* gamma, const float* beta, float* logits,
                 float* proto_sim /* [B, Nq, way, way] or NULL: TRX_sup prototype cosine matrix
                                     (model/classifiers/TRX_sup.py:114-164) */,
                 void* workspace, int need_grad /* 0 none, 1 logits only, 2 logits + proto_sim */, int* status,
                 void* stream);
/* inv_off [card*L + 1], inv_idx [card*T]: for (j, l) the tuples whose j-th frame is l.
 * Outputs are OVERWRITTEN: grad_support [B,Ns,L,D], grad_query [B,Nq,L,D], gWk/gWv [d, card*D],
 * gbk/gbv/ggamma/gbeta [d] -- unless `accumulate` says otherwise: bit 0 set = the six PARAMETER gradients are
 * added to what the buffers hold (gradient accumulation over micro-batches straight into .grad / an all-reduce
 * bucket); bit 1 set = grad_support / grad_query are added to (the cardinalities of a TrxBranch,
 * teacher/code/model.py:1094-1128, sum their feature gradients this way instead of in a separate pass). */
int lmkd_trx_bwd(const lmkd_trx_shape* s, const float* grad_logits,
                 const float* grad_proto_sim /* NULL unless the forward ran with need_grad = 2 */,
                 const int32_t* tuples, const int32_t* inv_off,
                 const int32_t* inv_idx, const float* bk, const float* gamma, const float* beta, float* grad_support,
                 float* grad_query, float* gWk, float* gbk, float* gWv, float* gbv, float* ggamma, float* gbeta,
                 void* workspace, int need_grad /* the value the forward ran with: 1 or 2 */,
                 int accumulate /* bit 0: parameter gradients, bit 1: feature gradients */, void* stream);
/* The attention block of lmkd_trx_fwd alone (TRX.py:120-148), exposed for tests and roofline benches: scores,
 * per-class softmax over the first cnt[b][c]*T of KTp = round_up(shot*T, 16) support tuples, prototype and
 * distance in ONE kernel; scores and probabilities stay in tensor memory.  bf16 inputs kq, vq [B, Nq*T, d],
 * ks, vs [B, way, KTp, d] (rows past cnt*T must be zero); outputs rowred / rowdot [B, way, Nq*T] are ACCUMULATED
 * into (sum diff^2, sum diff*prototype), linv = 1/rowsum; dq [B, way, Nq*T, d] bf16 = v_q - prototype and
 * patt [B, Nq*T, way*KTp] bf16 = exp(score - rowmax) are optional (NULL = not written; patt needs linv).
 * lmkd_trx_attn_fused_fits: 1 when the shape runs through this kernel inside lmkd_trx_fwd (KTp <= 384, d % 64 == 0). */
int lmkd_trx_attn_fused_fits(const lmkd_trx_shape* s);
/* Class groups wider than tensor memory (KTp > 384, e.g. 32-frame clips) and TRX_sup use materialised scores /
 * probabilities; those buffers are held for as many queries at a time as fit this byte budget (default 24e9, or
 * the environment variable LMKD_TRX_ATTN_BYTES; 0 restores the default).  With more than one pass the backward
 * recomputes the probabilities of each pass.  Must not change between lmkd_trx_workspace_bytes, lmkd_trx_fwd and
 * lmkd_trx_bwd of the same call. */
void lmkd_trx_set_attn_budget(double bytes);
int lmkd_trx_attn_fwd(const lmkd_trx_shape* s, const void* kq, const void* vq, const void* ks, const void* vs,
                      const int32_t* cnt, void* dq, void* patt, float* rowred, float* rowdot, float* linv,
                      void* stream);
/* ---- STRM DistanceLoss head (model/classifiers/strm_res18_sup.py:162-255; the same class in strmclassifiers.py and
 *      strmclassifiers_res18.py) -- SURVEY.md §8f rank 3 --------------------------------------------------------
 * logits[b][q][c] = -(1/T) sum_tau min_{(s, sigma): label[s] == c} | e_q[q, tau] - e_s[s, sigma] |_2 with
 * e[n, tau] = relu(W . concat(dropout(x)[n, tau_1..tau_card]) + bias) (no positional encoding).  Shape struct as for
 * TRX with d = output width of clsW (trans_linear_in_dim / 2); W [d, card*D], bias [d].  The backward routes the
 * gradient through the arg-min support tuple of every (query tuple, class), like torch.cdist + min.
 * Outputs of the backward are OVERWRITTEN: grad_support, grad_query, gW [d, card*D], gbias [d]. */
size_t lmkd_strm_dist_workspace_bytes(const lmkd_trx_shape* s, int need_grad);
int lmkd_strm_dist_fwd(const lmkd_trx_shape* s, const float* support, const float* labels, const float* query,
                       const int32_t* tuples, const float* W, const float* bias, float* logits, void* workspace,
                       int need_grad, int* status, void* stream);
int lmkd_strm_dist_bwd(const lmkd_trx_shape* s, const float* grad_logits, const int32_t* inv_off, const int32_t* inv_idx,
                       float* grad_support, float* grad_query, float* gW, float* gbias, void* workspace, void* stream);
/* dropout keep/scale mask exactly as the kernels generate it (test hook): out[i] in {0, 1/(1-p)} */
int lmkd_dropout_mask(float* out, int64_t n, float p, uint64_t seed, void* stream);

/* ---- SupportDK (model/classifiers/TRX_2fcsup.py:162-189) -----------------------------------
 * support [B, way*shot, L, D] class-sorted (labels ignored, as in the reference) ->
 * out [B, way, way-1]; protos [B, way, L, D] is caller-provided scratch kept for the backward. */
int lmkd_support_dk_fwd(const float* support, int B, int way, int shot, int L, int D, float* protos, float* out,
                        void* stream);
int lmkd_support_dk_bwd(const float* grad_out, const float* protos, int B, int way, int shot, int L, int D,
                        float* grad_support, void* stream);

/* ---- frame-mean Euclidean heads (model/classifiers/e_dist.py:22-61, COS.py:29-62) -------------
 * logits[b][q][c] = -mean_{s in class c} | mean_l query[b][q][l] - mean_l support[b][s][l] |_2 */
size_t lmkd_edist_workspace_bytes(int B, int Ns, int Nq, int D);
int lmkd_edist_fwd(const float* support, const float* labels, const float* query, int B, int Ns, int Nq, int L, int D,
                   int way, float* logits, void* workspace, int* status, void* stream);
int lmkd_edist_bwd(const float* grad_logits, const float* labels, int B, int Ns, int Nq, int L, int D, int way,
                   float* grad_support, float* grad_query, void* workspace, void* stream);

/* ---- Student feature heads feeding the path ----------------------------
 * Reference: model/backbone/resnet18_2fc.py:41-67 (adap_max -> reshape/permute/mean -> fc1, fc2 -> reshape) and
 * model/backbone/resnet18_student.py:38-58 (same with the single res18_2048 layer).
 *
 * lmkd_frame_pool_*: AdaptiveMaxPool2d((out_hw, out_hw)) followed by the mean over the out_hw^2 patches
 *   (resnet18_2fc.py:41-53): fmap [rows, C, H, W] fp32 (NCHW trunk output) -> pooled [rows, C].
 *   The backward routes grad_pooled / out_hw^2 to the first maximum of every window.
 * lmkd_feature_head_*: y[h] = x . weight[h]^T + bias[h] for `heads` Linear layers sharing the input
 *   (resnet18_2fc.py:55-64): x [rows, in_dim], weight [heads, out_dim, in_dim], bias [heads, out_dim] ->
 *   y [heads, rows, out_dim], i.e. y[h] viewed as [rows / L, L, out_dim] is context_features_{h+1}.
 *   bf16 operands, fp32 accumulation.  The backward writes grad_x = sum_h grad_y[h] . weight[h],
 *   grad_weight[h] = grad_y[h]^T . x and grad_bias[h] = column sums of grad_y[h]; any of the three may be NULL.
 *   in_dim and out_dim must be multiples of 8.  The forward keeps bf16 copies of x and weight in the workspace. */
int lmkd_frame_pool_fwd(const float* fmap, int64_t rows, int C, int H, int W, int out_hw, float* pooled,
                        void* stream);
int lmkd_frame_pool_bwd(const float* fmap, const float* grad_pooled, int64_t rows, int C, int H, int W, int out_hw,
                        float* grad_fmap, void* stream);
size_t lmkd_feature_head_workspace_bytes(int64_t rows, int in_dim, int out_dim, int heads);
int lmkd_feature_head_fwd(const float* x, const float* weight, const float* bias, int64_t rows, int in_dim,
                          int out_dim, int heads, float* y, void* workspace, void* stream);
int lmkd_feature_head_bwd(const float* grad_y, int64_t rows, int in_dim, int out_dim, int heads, float* grad_x,
                          float* grad_weight, float* grad_bias, void* workspace, void* stream);

/* ---- D2M losses (distillers.py) --------------------------------------------------------------
 * One additive term of a recipe on per-episode logits [B, rows, cols] (cols <= 64):
 *   kind 0 CE   : F.cross_entropy(s, y)                        (e.g. distillers.py:70)
 *   kind 1 KD   : kd_loss(s, t, T)                             (distillers.py:7-15)
 *   kind 2 ICR  : inter_class_relation(s, t)                   (distillers.py:26-30)
 * loss[b] = sum_i w_i (fa_i + fb_i * focal[b]) term_i[b]; focal = 1 - exp(-max(CE(fnum)/(CE(fden)+1e-8),0))
 * (the WSL weight, distillers.py:86-93), 0 if fnum is NULL.  grad (may be NULL) receives
 * d loss[b] / d s for an upstream gradient of 1. */
typedef struct {
  int kind, rows, cols;
  const float* s;
  const float* t;
  const int64_t* y;
  float* grad;
  int grad_accumulate;
  float w, fa, fb;
} lmkd_loss_term;

int lmkd_d2m_logit_loss(const lmkd_loss_term* terms /* HOST array */, int nterms, float temperature,
                        const float* fnum, const float* fden, const int64_t* fy, int frows, int fcols, int B,
                        float* loss /* [B] */, float* values /* [B, nterms] or NULL */,
                        float* focal /* [B] or NULL */, void* stream);

/* Fused feature-MSE forward + backward, ONE pass over HBM (KL_feature, distillers.py:126-150):
 *   *loss (+)= lscale * sum (s - t)^2 ;  ds = gscale * (s - t)
 * For F.mse_loss over n_e elements per episode with weight w: lscale = w / n_e, gscale = 2 w / n_e.
 * partials: caller scratch of lmkd_mse_partials() floats.  dtype 0 = fp32, 1 = bf16 storage. */
int lmkd_mse_partials(void);
int lmkd_d2m_feature_mse_fwdbwd(const void* s, const void* t, void* ds, int64_t n, int dtype, float lscale,
                                float gscale, float* partials, float* loss, int accumulate, void* stream);
/* ---- Teacher-feature store ----------------------------------------------
 * Reference: the teacher writes one [1, L, 2048] fp32 `feature.npy` per video
 * (teacher/code/extract_multi_feature.py:113-121); the student's loader reads them back one np.load per video and
 * concatenates them per episode (video_reader.py:388-395, 470-471).  Here the packed store
 * [store_rows, row_elems] (row = one video's L*D features, fp32 or bf16) is resident in HBM and an episode is a
 * list of row indices.  store_dtype 0 = fp32, 1 = bf16; row_elems a multiple of 8; an index outside
 * [0, store_rows) ORs 4 into *status (device int, may be NULL) and contributes zeros.
 *   lmkd_episode_gather: out[i] = float(store[index[i]]), i < count            (out [count, row_elems] fp32)
 *   lmkd_d2m_feature_mse_store_fwdbwd: the fused feature-MSE pass above with the teacher operand read in
 *     place through the indices: *loss (+)= lscale * sum (s - t)^2, ds = gscale * (s - t), t[i] = store[index[i]]. */
int lmkd_episode_gather(const void* store, int store_dtype, int64_t store_rows, const int64_t* index, int64_t count,
                        int64_t row_elems, float* out, int* status, void* stream);
int lmkd_d2m_feature_mse_store_fwdbwd(const float* s, const void* store, int store_dtype, int64_t store_rows,
                                      const int64_t* index, int64_t count, int64_t row_elems, float* ds, float lscale,
                                      float gscale, float* partials, float* loss, int accumulate, int* status,
                                      void* stream);
/* ---- Teacher multi-modal fusion forward ------------------------------------
 * Reference: ThreeTransforTemproal / TwoTransforFusion `.extract_feature` (teacher/code/model.py:1385-1392,
 * 1325-1331): per modality TrainablePositionalEncoding = LayerNorm(x + position embedding) (:1135-1151), feature-axis
 * concatenation, `nlayers` torch TransformerEncoderLayers (post-norm, ReLU, nhead = number of modalities,
 * dim_feedforward 2048, :1313-1316 / :1372-1375), then the `f1` Linear down to `dout`; eval mode, so every dropout
 * is the identity (extract_multi_feature.py:114).  ThreeTRXShiftLoopTime.extract_feature (:1648-1664) is three such
 * calls summed: (rgb, depth, flow), (rgb, depth rolled by shirt_num) and (rgb, flow rolled by shirt_num) --
 * `shift[m]` rolls modality m along the frame axis (x'[l] = x[(l + shift) % L]) and `accumulate` adds into `out`.
 * Matrices are bf16 [out_features, in_features] (cast once with lmkd_cast_bf16), biases / LayerNorm / embeddings fp32.
 * x: `nmod` device pointers to [nvideos, L, dmod] fp32.  out: [nvideos, L, dout] fp32. */
typedef struct {
  const void* w_qkv; const float* b_qkv;   /* self_attn.in_proj_weight [3d, d], in_proj_bias [3d] */
  const void* w_o;   const float* b_o;     /* self_attn.out_proj [d, d], [d] */
  const void* w_ff1; const float* b_ff1;   /* linear1 [dff, d], [dff] */
  const void* w_ff2; const float* b_ff2;   /* linear2 [d, dff], [d] */
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
} lmkd_fusion_layer;
typedef struct {
  int32_t nmod, dmod, nhead, dff, nlayers, dout;
  float ln_eps;
  const float* pe_emb[4];
  const float* pe_g[4];
  const float* pe_b[4];
  const lmkd_fusion_layer* layers;         /* HOST array [nlayers] */
  const void* w_out; const float* b_out;   /* f1 [dout, nmod * dmod], [dout] */
} lmkd_fusion_encoder;
size_t lmkd_fusion_workspace_bytes(const lmkd_fusion_encoder* enc, int64_t nvideos, int L);
int lmkd_fusion_fwd(const lmkd_fusion_encoder* enc, const float* const* x, const int32_t* shift /* host, or NULL */,
                    int64_t nvideos, int L, float* out, int accumulate, void* workspace, void* stream);

/* x *= *g unless *g == 1 — applies a device-resident upstream gradient without a host sync */
int lmkd_scale_by_device_scalar(float* x, int64_t n, const float* g, void* stream);

/* aggregate_accuracy (utils.py:116-121): *correct += #rows with argmax(logits) == label */
int lmkd_accuracy_count(const float* logits, const int64_t* labels, int64_t rows, int cols, int* correct,
                        void* stream);

/* ---- generic batched bf16 contraction on tcgen05 (exposed for tests and roofline benches) ----
 * C[b][m][n] (+)= alpha * sum_k A[b][m][k] B[b][n][k]; a_mn / b_mn = 1 when the operand is stored
 * [K][rows] (rows contiguous) instead of [rows][K].  A, B are bf16, C fp32. */
int lmkd_gemm_bf16(int M, int N, int K, int batch, const void* A, int a_mn, int64_t lda, int64_t a_bs,
                   const void* B, int b_mn, int64_t ldb, int64_t b_bs, float* C, int64_t ldc, int64_t c_bs,
                   float alpha, int accumulate, int block_n, void* stream);
int lmkd_cast_bf16(const float* x, void* y, int64_t n, void* stream);
/* y (fp32) = x (bf16): student features staged from the host in bf16 (half the H2D bytes; the heads round their
 * inputs to bf16 anyway) are widened into the fp32 tensors the classifier API takes.  n a multiple of 8. */
int lmkd_upcast_bf16(const void* x, float* y, int64_t n, void* stream);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------
 * lmkd_launch_count: kernels this library has launched in this process (reset != 0 zeroes it).
 * lmkd_gemm_timing_*: when enabled every tcgen05 GEMM launch is bracketed by CUDA events on its
 * stream; _read synchronises on them and returns total kernel ms, true-shape FLOPs
 * (2*M*N*K*batch) and the launch count, then clears the record.  Host pointers. */
/* lmkd_kernel_timing_read: the same record by kernel category (enabled by lmkd_gemm_timing_enable):
 *   0 tcgen05 contractions incl. the fused attention kernel (work = FLOPs)   1 tuple assembly / LayerNorm fwd+bwd
 *   (work = algorithmic bytes)   2 OTAM recurrence fwd / bwd (work = DP cells)   3 fused feature-MSE (work = bytes). */
int lmkd_kernel_timing_read(int category, double* ms, double* work, int* launches);
long long lmkd_launch_count(int reset);
void lmkd_gemm_timing_enable(int on);
int lmkd_gemm_timing_read(double* ms, double* flops, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* LMKD_H_ */
