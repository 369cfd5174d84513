// D2M distillation losses (reference: distillers.py) and the small support-level head.
#pragma once
#include "common.cuh"

namespace lmkd {

enum TermKind : int { TERM_CE = 0, TERM_KD = 1, TERM_ICR = 2 };

// One additive term of a Distiller recipe, evaluated per episode on [rows, cols] logits.
//   CE : mean_rows -log softmax(s)[y]                      (F.cross_entropy, distillers.py)
//   KD : T^2 mean_rows KL(softmax(t/T) || softmax(s/T))    (kd_loss, distillers.py:7-15)
//   ICR: 1 - mean_rows pearson(softmax(s), softmax(t))     (inter_class_relation, :26-30)
// coefficient = w * (fa + fb * focal) where focal = 1 - exp(-max(CE(fnum)/(CE(fden)+1e-8), 0))
struct LossTerm {
  int kind;
  int rows, cols;
  const float* s;        // student logits [B, rows, cols]
  const float* t;        // teacher logits [B, rows, cols]       (KD, ICR)
  const int64_t* y;      // labels [B, rows]                      (CE)
  float* grad;           // d loss / d s, [B, rows, cols]; may be null; several terms may share it
  int grad_accumulate;   // 0: overwrite, 1: += (second and later terms on the same tensor)
  float w, fa, fb;
};

constexpr int kMaxTerms = 8;

struct LossSpec {
  int nterms;
  LossTerm terms[kMaxTerms];
  float temperature;
  // focal weight inputs (null fnum => focal = 0)
  const float* fnum;     // logits whose CE is the numerator   [B, frows, fcols]
  const float* fden;     // logits whose CE is the denominator
  const int64_t* fy;     // labels [B, frows]
  int frows, fcols;
};

// loss[b] = sum_i coef_i * term_i ; values[b][i] = term_i (unweighted); focal[b] (if non-null)
int d2m_logit_loss(const LossSpec& spec, int B, float* loss, float* values, float* focal, cudaStream_t st);

// Fused feature MSE forward+backward in ONE pass over HBM (KL_feature, distillers.py:141):
//   partial sums of (s-t)^2 per block -> `partials`; ds = gscale * (s - t) written streaming.
// loss contribution = lscale * sum (s-t)^2, finished by mse_finish into *loss_out (+=).
int feat_mse_fwdbwd(const float* s, const float* t, float* ds, int64_t n, float gscale, float* partials,
                    int max_partials, int* npartials, cudaStream_t st);
int feat_mse_fwdbwd_bf16(const __nv_bfloat16* s, const __nv_bfloat16* t, __nv_bfloat16* ds, int64_t n, float gscale,
                         float* partials, int max_partials, int* npartials, cudaStream_t st);
int mse_finish(const float* partials, int npartials, float lscale, float* loss_out, int accumulate, cudaStream_t st);

// x *= *g unless *g == 1 (device scalar); used by autograd backward without a host sync
int scale_by_device_scalar(float* x, int64_t n, const float* g, cudaStream_t st);

// SupportDK (TRX_2fcsup.py:162-189): protos = mean over shots; out[b][i][m] = -|p_i - p_n|^2 / L
int support_dk_fwd(const float* support, float* protos, float* out, int B, int way, int shot, int L, int D,
                   cudaStream_t st);
int support_dk_bwd(const float* gout, const float* protos, float* gsupport, int B, int way, int shot, int L, int D,
                   cudaStream_t st);

// aggregate_accuracy (utils.py:116-121): correct[0] += #(argmax == label) over rows; first max wins
int accuracy_count(const float* logits, const int64_t* labels, int64_t rows, int cols, int* correct, cudaStream_t st);

}  // namespace lmkd
