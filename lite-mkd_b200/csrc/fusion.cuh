// Teacher multi-modal fusion forward; see fusion.cu.
// Reference semantics: teacher/code/model.py:1135-1151, 1300-1331, 1361-1392, 1648-1664.
#pragma once
#include "common.cuh"

namespace lmkd {

// one torch.nn.TransformerEncoderLayer (post-norm, ReLU); matrices bf16 [out, in], everything else fp32
struct FusionLayer {
  const void* w_qkv; const float* b_qkv;     // self_attn.in_proj_{weight,bias}: [3d, d], [3d]
  const void* w_o;   const float* b_o;       // self_attn.out_proj: [d, d], [d]
  const void* w_ff1; const float* b_ff1;     // linear1: [dff, d], [dff]
  const void* w_ff2; const float* b_ff2;     // linear2: [d, dff], [d]
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

struct FusionEncoder {
  int nmod, dmod, nhead, dff, nlayers, dout;
  float ln_eps;
  const float* pe_emb[4];                    // position_embeddings.weight [>= L, dmod] per modality
  const float* pe_g[4];
  const float* pe_b[4];                      // LayerNorm of the positional encoding
  const FusionLayer* layers;                 // host array [nlayers]
  const void* w_out; const float* b_out;     // f1: bf16 [dout, nmod * dmod], [dout]
};

struct FusionWs {
  float *X, *Y;
  __nv_bfloat16 *Xb, *qkv, *ctx, *H;
  size_t bytes;
};

int fusion_check(const FusionEncoder& e, int64_t nvideos, int L);
FusionWs fusion_layout(void* ws, const FusionEncoder& e, int64_t nvideos, int L);
// x: nmod device pointers [nvideos, L, dmod] fp32; shift: per-modality temporal roll (host, may be null);
// out [nvideos, L, dout] fp32, written (accumulate == 0) or added to
int fusion_forward(const FusionEncoder& e, const float* const* x, const int* shift, int64_t nvideos, int L, float* out,
                   int accumulate, void* workspace, cudaStream_t st);

}  // namespace lmkd
