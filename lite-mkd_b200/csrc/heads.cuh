// Frame-mean Euclidean heads (see heads.cu)
#pragma once
#include "common.cuh"

namespace lmkd {

// sm [B*Ns, D], qm [B*Nq, D] frame means and pd [B, Nq, Ns] pair distances are kept for the backward
int edist_fwd(const float* support, const float* labels, const float* query, float* sm, float* qm, float* pd,
              float* logits, int B, int Ns, int Nq, int L, int D, int way, int* status, cudaStream_t st);
int edist_bwd(const float* glogits, const float* labels, const float* sm, const float* qm, const float* pd,
              float* gsupport, float* gquery, int B, int Ns, int Nq, int L, int D, int way, cudaStream_t st);

}  // namespace lmkd
