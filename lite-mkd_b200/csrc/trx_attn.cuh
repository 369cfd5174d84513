// Fused class-grouped tuple attention of the TRX head (forward), see trx_attn.cu.
// Reference semantics: model/classifiers/TRX.py:120-148 (scores, per-class softmax over the K*T support tuples,
// prototype, squared distance); generic cardinality: teacher/code/model.py:296-335.
#pragma once
#include "trx.cuh"

namespace lmkd {

struct TrxAttnFwd {
  const __nv_bfloat16* kq;   // [B, NqT, d]        LayerNorm'd query keys
  const __nv_bfloat16* vq;   // [B, NqT, d]        query values
  const __nv_bfloat16* ks;   // [B, way, KTp, d]   class-sorted support keys (zero rows for padding)
  const __nv_bfloat16* vs;   // [B, way, KTp, d]
  const int* cnt;            // [B, way]           supports per class
  __nv_bfloat16* dq;         // [B, way, NqT, d]   v_q - prototype_c, or null (no-grad passes)
  __nv_bfloat16* patt;       // [B, NqT, way*KTp]  exp(score - rowmax), un-normalised, or null
  float* rowred;             // [B, way, NqT]      sum_i diff^2             (+=: zeroed by the caller)
  float* rowdot;             // [B, way, NqT]      sum_i diff * prototype   (+=), or null
  float* linv;               // [B, way, NqT]      1 / sum_j exp(score - rowmax), or null
};

// true when the class group fits tensor memory next to the output stages (KTp <= 384) and d is a multiple of 64
bool trx_attn_fused_fits(const TrxDims& s);
int trx_attn_fwd(const TrxAttnFwd& a, const TrxDims& s, cudaStream_t st);

}  // namespace lmkd
