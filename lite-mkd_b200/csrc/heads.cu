// Frame-mean prototype heads (reference: model/classifiers/e_dist.py:22-61, COS.py:29-62):
//   logit[q][c] = -mean_{s in class c} | mean_l query[q][l] - mean_l support[s][l] |_2
// fp32 throughout; the only HBM-sized traffic is one read of the features (forward) and one
// write of their gradient (backward).
#include "heads.cuh"

namespace lmkd {

namespace {

// m[row][:] = mean over the L frames of x[row][l][:]   (thread per float4 column)
__global__ void frame_mean_kernel(const float* __restrict__ x, float* __restrict__ m, int64_t rows, int L, int D4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * D4) return;
  const int64_t row = i / D4;
  const int c4 = static_cast<int>(i - row * D4);
  const float4* src = reinterpret_cast<const float4*>(x) + row * L * D4 + c4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < L; ++l) {
    const float4 v = __ldg(src + static_cast<int64_t>(l) * D4);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const float inv = 1.f / L;
  reinterpret_cast<float4*>(m)[i] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}

// block per (b, q): pair distances to every support, then the class means
__global__ void __launch_bounds__(128)
edist_fwd_kernel(const float* __restrict__ qm, const float* __restrict__ sm, const float* __restrict__ labels,
                 float* __restrict__ pd, float* __restrict__ logits, int Nq, int Ns, int D4, int way,
                 int* __restrict__ status) {
  extern __shared__ float spd[];                 // [Ns]
  const int64_t bq = blockIdx.x;
  const int64_t b = bq / Nq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4* q4 = reinterpret_cast<const float4*>(qm) + bq * D4;
  for (int s = warp; s < Ns; s += 4) {
    const float4* s4 = reinterpret_cast<const float4*>(sm) + (b * Ns + s) * D4;
    float acc = 0.f;
    for (int c4 = lane; c4 < D4; c4 += 32) {
      const float4 a = __ldg(q4 + c4), y = __ldg(s4 + c4);
      const float d0 = a.x - y.x, d1 = a.y - y.y, d2 = a.z - y.z, d3 = a.w - y.w;
      acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float dist = sqrtf(acc);
      spd[s] = dist;
      pd[bq * Ns + s] = dist;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < way; c += blockDim.x) {
    float sum = 0.f;
    int cnt = 0;
    for (int s = 0; s < Ns; ++s) {
      const int cls = static_cast<int>(labels[b * Ns + s]);
      if (cls < 0 || cls >= way) { if (status) atomicOr(status, 1); continue; }
      if (cls == c) { sum += spd[s]; ++cnt; }
    }
    logits[bq * way + c] = cnt > 0 ? -sum / cnt : 0.f;     // absent class keeps the zero of e_dist.py:41
  }
}

// grid (B, Nq + Ns): gradient of one frame-mean embedding, broadcast to its L frames
__global__ void __launch_bounds__(256)
edist_bwd_kernel(const float* __restrict__ glogits, const float* __restrict__ qm, const float* __restrict__ sm,
                 const float* __restrict__ labels, const float* __restrict__ pd, float* __restrict__ gquery,
                 float* __restrict__ gsupport, int Nq, int Ns, int L, int D4, int way) {
  extern __shared__ float w[];                   // coefficient of (qm - sm) for every partner
  const int64_t b = blockIdx.x;
  const int idx = blockIdx.y;
  const bool is_q = idx < Nq;
  const int partners = is_q ? Ns : Nq;
  for (int p = threadIdx.x; p < partners; p += blockDim.x) {
    const int q = is_q ? idx : p, s = is_q ? p : idx - Nq;
    const int cls = static_cast<int>(labels[b * Ns + s]);
    float coef = 0.f;
    if (cls >= 0 && cls < way) {
      int cnt = 0;
      for (int j = 0; j < Ns; ++j) cnt += (static_cast<int>(labels[b * Ns + j]) == cls);
      const float dist = pd[(b * Nq + q) * Ns + s];
      // logit = -mean dist ; d dist / d qm = (qm - sm) / dist  (0 at dist = 0, like torch.cdist)
      if (dist > 0.f) coef = -glogits[(b * Nq + q) * way + cls] / (cnt * dist);
    }
    w[p] = coef;
  }
  __syncthreads();
  const float4* me = reinterpret_cast<const float4*>(is_q ? qm : sm) + (b * (is_q ? Nq : Ns) + (is_q ? idx : idx - Nq)) * D4;
  const float4* others = reinterpret_cast<const float4*>(is_q ? sm : qm) + b * partners * D4;
  float4* dst = reinterpret_cast<float4*>(is_q ? gquery : gsupport) +
                (b * (is_q ? Nq : Ns) + (is_q ? idx : idx - Nq)) * static_cast<int64_t>(L) * D4;
  const float sign = is_q ? 1.f : -1.f;          // d/d sm = -(qm - sm) / dist
  const float inv_l = 1.f / L;
  for (int c4 = threadIdx.x; c4 < D4; c4 += blockDim.x) {
    const float4 m = __ldg(me + c4);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < partners; ++p) {
      const float4 o = __ldg(others + static_cast<int64_t>(p) * D4 + c4);
      const float k = w[p];
      // (qm - sm) seen from this side: me - other for a query, other - me for a support
      a.x += k * (is_q ? m.x - o.x : o.x - m.x);
      a.y += k * (is_q ? m.y - o.y : o.y - m.y);
      a.z += k * (is_q ? m.z - o.z : o.z - m.z);
      a.w += k * (is_q ? m.w - o.w : o.w - m.w);
    }
    a.x *= sign * inv_l; a.y *= sign * inv_l; a.z *= sign * inv_l; a.w *= sign * inv_l;
    for (int l = 0; l < L; ++l) dst[static_cast<int64_t>(l) * D4 + c4] = a;
  }
}

}  // namespace

int edist_fwd(const float* support, const float* labels, const float* query, float* sm, float* qm, float* pd,
              float* logits, int B, int Ns, int Nq, int L, int D, int way, int* status, cudaStream_t st) {
  LMKD_CHECK(D % 4 == 0, "e_dist: feature dim must be a multiple of 4");
  const int D4 = D / 4;
  const int64_t rs = static_cast<int64_t>(B) * Ns, rq = static_cast<int64_t>(B) * Nq;
  frame_mean_kernel<<<static_cast<unsigned>(ceil_div(rs * D4, 256)), 256, 0, st>>>(support, sm, rs, L, D4);
  LMKD_LAUNCH_CHECK("frame_mean_kernel");
  frame_mean_kernel<<<static_cast<unsigned>(ceil_div(rq * D4, 256)), 256, 0, st>>>(query, qm, rq, L, D4);
  LMKD_LAUNCH_CHECK("frame_mean_kernel");
  edist_fwd_kernel<<<static_cast<unsigned>(rq), 128, sizeof(float) * Ns, st>>>(qm, sm, labels, pd, logits, Nq, Ns, D4,
                                                                              way, status);
  LMKD_LAUNCH_CHECK("edist_fwd_kernel");
  return 0;
}

int edist_bwd(const float* glogits, const float* labels, const float* sm, const float* qm, const float* pd,
              float* gsupport, float* gquery, int B, int Ns, int Nq, int L, int D, int way, cudaStream_t st) {
  const int D4 = D / 4;
  dim3 grid(B, Nq + Ns);
  const size_t smem = sizeof(float) * (Ns > Nq ? Ns : Nq);
  edist_bwd_kernel<<<grid, 256, smem, st>>>(glogits, qm, sm, labels, pd, gquery, gsupport, Nq, Ns, L, D4, way);
  LMKD_LAUNCH_CHECK("edist_bwd_kernel");
  return 0;
}

}  // namespace lmkd
