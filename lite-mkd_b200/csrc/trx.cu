// TRX non-GEMM kernels.  The dense tuple tensor of the reference (TRX.py:90-94, [N, T, c*D])
// is never built: the Linear over concatenated frames is factored into per-frame partial
// projections P[(b,n,l)][(which, j)][:] = X~[(b,n,l)] . W_which[:, j*D:(j+1)*D]^T (one big tcgen05
// GEMM), and the kernels below assemble tuples as sums of c rows of P.
#include "trx.cuh"

namespace lmkd {

namespace {

constexpr int kWarps = 8;

// ---- class slots ---------------------------------------------------------------------------
__global__ void class_slots_kernel(const float* __restrict__ labels, int* __restrict__ slot,
                                   int* __restrict__ cnt, int* __restrict__ status, int B, int Ns, int way,
                                   int shot) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int* c = cnt + static_cast<int64_t>(b) * way;
  for (int i = 0; i < way; ++i) c[i] = 0;
  for (int n = 0; n < Ns; ++n) {
    const int cls = static_cast<int>(labels[static_cast<int64_t>(b) * Ns + n]);
    int sl = -1;
    if (cls < 0 || cls >= way) {
      if (status) atomicOr(status, 1);      // label outside [0, way)
    } else if (c[cls] >= shot) {
      if (status) atomicOr(status, 2);      // more than `shot` supports in one class
    } else {
      sl = cls * shot + c[cls];
      c[cls]++;
    }
    slot[static_cast<int64_t>(b) * Ns + n] = sl;
  }
}

// ---- tuple assembly + LayerNorm (forward) ----------------------------------------------------
// block = one video (b, n); warps loop over its T tuples.  smem: kWarps x d floats.
__global__ void __launch_bounds__(kWarps * 32)
tuple_ln_fwd_kernel(const float* __restrict__ P, const float* __restrict__ bk, const float* __restrict__ bv,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const int* __restrict__ tuples, const int* __restrict__ slot,
                    __nv_bfloat16* __restrict__ Kq, __nv_bfloat16* __restrict__ Vq,
                    __nv_bfloat16* __restrict__ Ks, __nv_bfloat16* __restrict__ Vs, float* __restrict__ stats,
                    float ln_eps, const TrxDims s) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xrow = smem + warp * s.d;
  const int64_t vid = blockIdx.x;                 // b * N + n
  const int n = static_cast<int>(vid % s.N);
  const int64_t b = vid / s.N;
  const int64_t pcols = 2ll * s.card * s.d;
  const float* Pv = P + vid * s.L * pcols;        // this video's L frame rows
  int64_t out_row;                                // row in the destination K/V buffer, -1 = dropped
  __nv_bfloat16 *Kd, *Vd;
  if (n < s.Ns) {
    const int sl = slot[b * s.Ns + n];
    Kd = Ks; Vd = Vs;
    out_row = sl < 0 ? -1 : (b * s.way + sl / s.shot) * s.KTp + static_cast<int64_t>(sl % s.shot) * s.T;
  } else {
    Kd = Kq; Vd = Vq;
    out_row = b * s.NqT + static_cast<int64_t>(n - s.Ns) * s.T;
  }
  for (int tau = warp; tau < s.T; tau += kWarps) {
    const int* tp = tuples + tau * s.card;
    float sum = 0.f;
    for (int i = lane; i < s.d; i += 32) {
      float x = __ldg(bk + i);
      for (int j = 0; j < s.card; ++j) x += __ldg(Pv + tp[j] * pcols + static_cast<int64_t>(j) * s.d + i);
      xrow[i] = x;
      sum += x;
    }
    sum = warp_sum(sum);
    const float mean = sum / s.d;
    float var = 0.f;
    for (int i = lane; i < s.d; i += 32) {
      const float dlt = xrow[i] - mean;
      var += dlt * dlt;
    }
    var = warp_sum(var) / s.d;
    const float rstd = rsqrtf(var + ln_eps);
    if (lane == 0) {
      stats[(vid * s.T + tau) * 2 + 0] = mean;
      stats[(vid * s.T + tau) * 2 + 1] = rstd;
    }
    if (out_row >= 0) {
      __nv_bfloat16* kd = Kd + (out_row + tau) * s.d;
      __nv_bfloat16* vd = Vd + (out_row + tau) * s.d;
      for (int i = lane; i < s.d; i += 32) {
        kd[i] = __float2bfloat16_rn((xrow[i] - mean) * rstd * __ldg(gamma + i) + __ldg(beta + i));
        float v = __ldg(bv + i);
        for (int j = 0; j < s.card; ++j)
          v += __ldg(Pv + tp[j] * pcols + static_cast<int64_t>(s.card + j) * s.d + i);
        vd[i] = __float2bfloat16_rn(v);
      }
    }
    __syncwarp();
  }
}

// ---- class-grouped softmax ---------------------------------------------------------------------
// one warp per score row (b, m); each class group of KTp columns is an independent softmax over
// its first cnt*T columns (TRX.py:127-134); padding columns get probability 0
__global__ void softmax_fwd_kernel(const float* __restrict__ S, const int* __restrict__ cnt,
                                   __nv_bfloat16* __restrict__ Patt, const TrxDims s) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= static_cast<int64_t>(s.B) * s.NqT) return;
  const int64_t b = row / s.NqT;
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  for (int c = 0; c < s.way; ++c) {
    const int valid = cnt[b * s.way + c] * s.T;
    const float* src = S + row * pitch + static_cast<int64_t>(c) * s.KTp;
    __nv_bfloat16* dst = Patt + row * pitch + static_cast<int64_t>(c) * s.KTp;
    float mx = -INFINITY;
    for (int i = lane; i < valid; i += 32) mx = fmaxf(mx, src[i]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int i = lane; i < valid; i += 32) den += __expf(src[i] - mx);
    den = warp_sum(den);
    const float inv = valid > 0 ? 1.f / den : 0.f;
    for (int i = lane; i < s.KTp; i += 32)
      dst[i] = __float2bfloat16_rn(i < valid ? __expf(src[i] - mx) * inv : 0.f);
  }
}

__global__ void softmax_bwd_kernel(const __nv_bfloat16* __restrict__ Patt, const float* __restrict__ dP,
                                   const int* __restrict__ cnt, __nv_bfloat16* __restrict__ dS, const TrxDims s) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= static_cast<int64_t>(s.B) * s.NqT) return;
  const int64_t b = row / s.NqT;
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  for (int c = 0; c < s.way; ++c) {
    const int valid = cnt[b * s.way + c] * s.T;
    const int64_t off = row * pitch + static_cast<int64_t>(c) * s.KTp;
    float dot = 0.f;
    for (int i = lane; i < valid; i += 32) dot += __bfloat162float(Patt[off + i]) * dP[off + i];
    dot = warp_sum(dot);
    for (int i = lane; i < s.KTp; i += 32)
      dS[off + i] = __float2bfloat16_rn(i < valid ? __bfloat162float(Patt[off + i]) * (dP[off + i] - dot) : 0.f);
  }
}

__global__ void logits_fwd_kernel(const float* __restrict__ rowred, const int* __restrict__ cnt,
                                  float* __restrict__ logits, const TrxDims s) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // (b, q, c)
  if (i >= static_cast<int64_t>(s.B) * s.Nq * s.way) return;
  const int c = static_cast<int>(i % s.way);
  const int q = static_cast<int>((i / s.way) % s.Nq);
  const int64_t b = i / (static_cast<int64_t>(s.way) * s.Nq);
  float acc = 0.f;
  if (cnt[b * s.way + c] > 0) {
    const float* src = rowred + (b * s.way + c) * s.NqT + static_cast<int64_t>(q) * s.T;
    for (int t = 0; t < s.T; ++t) acc += src[t];
    acc = -acc / s.T;
  }
  logits[i] = acc;   // classes without supports keep logit 0 (TRX.py:118 zeros init)
}

__global__ void attn_bwd_prep_kernel(const float* __restrict__ glogits, const int* __restrict__ cnt,
                                     const __nv_bfloat16* __restrict__ Patt, float* __restrict__ srow,
                                     __nv_bfloat16* __restrict__ Ps, const TrxDims s) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;   // (b, m)
  if (row >= static_cast<int64_t>(s.B) * s.NqT) return;
  const int64_t b = row / s.NqT;
  const int m = static_cast<int>(row % s.NqT);
  const int q = m / s.T;
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  for (int c = 0; c < s.way; ++c) {
    const float g = cnt[b * s.way + c] > 0 ? glogits[(b * s.Nq + q) * s.way + c] : 0.f;
    const float sc = 2.f * g / s.T;
    if (lane == 0) srow[(b * s.way + c) * s.NqT + m] = sc;
    const int64_t off = row * pitch + static_cast<int64_t>(c) * s.KTp;
    for (int i = lane; i < s.KTp; i += 32) Ps[off + i] = __float2bfloat16_rn(__bfloat162float(Patt[off + i]) * sc);
  }
}

// ---- LayerNorm backward per tuple row -----------------------------------------------------------
// persistent grid; warp per tuple row; per-block smem accumulators for dgamma/dbeta/dbk/dbv
__global__ void __launch_bounds__(kWarps * 32)
ln_bwd_kernel(const float* __restrict__ P, const float* __restrict__ bk, const float* __restrict__ gamma,
              const float* __restrict__ stats, const int* __restrict__ tuples, const int* __restrict__ slot,
              const float* __restrict__ dKq, const float* __restrict__ dKs, const float* __restrict__ dVs,
              const float* __restrict__ srow, const __nv_bfloat16* __restrict__ Dq, float* __restrict__ dxk,
              float* __restrict__ dxv, float* __restrict__ partials, const TrxDims s) {
  extern __shared__ float smem[];
  float* acc = smem;                                  // [4][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* gxh = smem + 4 * s.d + warp * 2 * s.d;       // per-warp stash: gamma*gy and xhat
  float* xh = gxh + s.d;
  for (int i = threadIdx.x; i < 4 * s.d; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int64_t pcols = 2ll * s.card * s.d;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kWarps + warp; row < s.R;
       row += static_cast<int64_t>(gridDim.x) * kWarps) {
    const int tau = static_cast<int>(row % s.T);
    const int64_t vid = row / s.T;
    const int n = static_cast<int>(vid % s.N);
    const int64_t b = vid / s.N;
    const int* tp = tuples + tau * s.card;
    const float* Pv = P + vid * s.L * pcols;
    const float* dk = nullptr;
    const float* dv = nullptr;
    int64_t m = -1;
    if (n < s.Ns) {
      const int sl = slot[b * s.Ns + n];
      if (sl >= 0) {
        const int64_t r = (b * s.way + sl / s.shot) * s.KTp + static_cast<int64_t>(sl % s.shot) * s.T + tau;
        dk = dKs + r * s.d;
        dv = dVs + r * s.d;
      }
    } else {
      m = static_cast<int64_t>(n - s.Ns) * s.T + tau;
      dk = dKq + (b * s.NqT + m) * s.d;
    }
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < s.d; i += 32) {
      float x = __ldg(bk + i);
      for (int j = 0; j < s.card; ++j) x += __ldg(Pv + tp[j] * pcols + static_cast<int64_t>(j) * s.d + i);
      const float xhat = (x - mean) * rstd;
      const float gy = dk ? dk[i] : 0.f;
      const float g = gy * __ldg(gamma + i);
      gxh[i] = g;
      xh[i] = xhat;
      s1 += g;
      s2 += g * xhat;
      if (gy != 0.f) {
        atomicAdd(acc + i, gy * xhat);         // dgamma
        atomicAdd(acc + s.d + i, gy);          // dbeta
      }
    }
    s1 = warp_sum(s1) / s.d;
    s2 = warp_sum(s2) / s.d;
    for (int i = lane; i < s.d; i += 32) {
      const float dx = rstd * (gxh[i] - s1 - xh[i] * s2);
      dxk[row * s.d + i] = dx;
      float gv;
      if (n < s.Ns) {
        gv = dv ? dv[i] : 0.f;
      } else {
        // d logit / d v_q = -sum_c srow[c][m] * (v_q - O_c)     (srow = 2 g / T)
        gv = 0.f;
        for (int c = 0; c < s.way; ++c) {
          const int64_t rc = (b * s.way + c) * s.NqT + m;
          gv -= srow[rc] * __bfloat162float(Dq[rc * s.d + i]);
        }
      }
      dxv[row * s.d + i] = gv;
      if (dx != 0.f) atomicAdd(acc + 2 * s.d + i, dx);   // dbk
      if (gv != 0.f) atomicAdd(acc + 3 * s.d + i, gv);   // dbv
    }
    __syncwarp();
  }
  __syncthreads();
  float* out = partials + static_cast<int64_t>(blockIdx.x) * 4 * s.d;
  for (int i = threadIdx.x; i < 4 * s.d; i += blockDim.x) out[i] = acc[i];
}

__global__ void reduce_partials_kernel(const float* __restrict__ partials, int nblocks, float* __restrict__ ggamma,
                                       float* __restrict__ gbeta, float* __restrict__ gbk, float* __restrict__ gbv,
                                       int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over 4*d
  if (i >= 4 * d) return;
  float acc = 0.f;
  for (int b = 0; b < nblocks; ++b) acc += partials[static_cast<int64_t>(b) * 4 * d + i];
  const int which = i / d, k = i - which * d;
  float* dst = which == 0 ? ggamma : which == 1 ? gbeta : which == 2 ? gbk : gbv;
  dst[k] = acc;
}

// ---- tuple gather (backward of the assembly) -----------------------------------------------------
// grid (M frame rows, 2 which); threads over d
__global__ void tuple_gather_bwd_kernel(const float* __restrict__ dxk, const float* __restrict__ dxv,
                                        const int* __restrict__ inv_off, const int* __restrict__ inv_idx,
                                        __nv_bfloat16* __restrict__ dPcat, const TrxDims s) {
  const int64_t frow = blockIdx.x;                  // (b, n, l)
  const int which = blockIdx.y;
  const int l = static_cast<int>(frow % s.L);
  const int64_t vid = frow / s.L;
  const float* src = (which == 0 ? dxk : dxv) + vid * s.T * s.d;
  const int64_t pcols = 2ll * s.card * s.d;
  for (int j = 0; j < s.card; ++j) {
    const int beg = inv_off[j * s.L + l], end = inv_off[j * s.L + l + 1];
    __nv_bfloat16* dst = dPcat + frow * pcols + static_cast<int64_t>(which * s.card + j) * s.d;
    for (int i = threadIdx.x; i < s.d; i += blockDim.x) {
      float acc = 0.f;
      for (int e = beg; e < end; ++e) acc += __ldg(src + static_cast<int64_t>(inv_idx[e]) * s.d + i);
      dst[i] = __float2bfloat16_rn(acc);
    }
  }
}

__global__ void pack_weights_kernel(const float* __restrict__ Wk, const float* __restrict__ Wv,
                                    __nv_bfloat16* __restrict__ Wcat, int d, int D, int card) {
  // Wcat[which][j][i][col] = W_which[i][j*D + col]
  const int64_t total = 2ll * card * d * D;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(t % D);
    const int i = static_cast<int>((t / D) % d);
    const int j = static_cast<int>((t / (static_cast<int64_t>(D) * d)) % card);
    const int which = static_cast<int>(t / (static_cast<int64_t>(D) * d * card));
    const float* W = which == 0 ? Wk : Wv;
    Wcat[t] = __float2bfloat16_rn(__ldg(W + static_cast<int64_t>(i) * card * D + static_cast<int64_t>(j) * D + col));
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dWcat, float* __restrict__ gWk,
                                    float* __restrict__ gWv, int d, int D, int card) {
  const int64_t total = 2ll * card * d * D;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(t % D);
    const int i = static_cast<int>((t / D) % d);
    const int j = static_cast<int>((t / (static_cast<int64_t>(D) * d)) % card);
    const int which = static_cast<int>(t / (static_cast<int64_t>(D) * d * card));
    float* G = which == 0 ? gWk : gWv;
    G[static_cast<int64_t>(i) * card * D + static_cast<int64_t>(j) * D + col] = dWcat[t];
  }
}

int grid_for(int64_t items, int threads) {
  int64_t blocks = ceil_div(items, threads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

int trx_class_slots(const float* labels, int* slot, int* cnt, int* status, const TrxDims& s, cudaStream_t st) {
  class_slots_kernel<<<static_cast<unsigned>(ceil_div(s.B, 64)), 64, 0, st>>>(labels, slot, cnt, status, s.B, s.Ns,
                                                                             s.way, s.shot);
  LMKD_LAUNCH_CHECK("class_slots_kernel");
  return 0;
}

int trx_tuple_ln_fwd(const float* P, const float* bk, const float* bv, const float* gamma, const float* beta,
                     const int* tuples, const int* slot, __nv_bfloat16* Kq, __nv_bfloat16* Vq, __nv_bfloat16* Ks,
                     __nv_bfloat16* Vs, float* stats, float ln_eps, const TrxDims& s, cudaStream_t st) {
  const size_t smem = sizeof(float) * kWarps * s.d;
  LMKD_CHECK(smem <= 160 * 1024, "trans_linear_out_dim %d too large", s.d);
  static bool attr_set = false;
  if (!attr_set) {
    LMKD_CUDA(cudaFuncSetAttribute(tuple_ln_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  tuple_ln_fwd_kernel<<<static_cast<unsigned>(static_cast<int64_t>(s.B) * s.N), kWarps * 32, smem, st>>>(
      P, bk, bv, gamma, beta, tuples, slot, Kq, Vq, Ks, Vs, stats, ln_eps, s);
  LMKD_LAUNCH_CHECK("tuple_ln_fwd_kernel");
  return 0;
}

int trx_softmax_fwd(const float* S, const int* cnt, __nv_bfloat16* Patt, const TrxDims& s, cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(s.B) * s.NqT;
  softmax_fwd_kernel<<<static_cast<unsigned>(ceil_div(rows * 32, 256)), 256, 0, st>>>(S, cnt, Patt, s);
  LMKD_LAUNCH_CHECK("softmax_fwd_kernel");
  return 0;
}

int trx_logits_fwd(const float* rowred, const int* cnt, float* logits, const TrxDims& s, cudaStream_t st) {
  const int64_t n = static_cast<int64_t>(s.B) * s.Nq * s.way;
  logits_fwd_kernel<<<static_cast<unsigned>(ceil_div(n, 128)), 128, 0, st>>>(rowred, cnt, logits, s);
  LMKD_LAUNCH_CHECK("logits_fwd_kernel");
  return 0;
}

int trx_attn_bwd_prep(const float* glogits, const int* cnt, const __nv_bfloat16* Patt, float* srow,
                      __nv_bfloat16* Ps, const TrxDims& s, cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(s.B) * s.NqT;
  attn_bwd_prep_kernel<<<static_cast<unsigned>(ceil_div(rows * 32, 256)), 256, 0, st>>>(glogits, cnt, Patt, srow, Ps,
                                                                                       s);
  LMKD_LAUNCH_CHECK("attn_bwd_prep_kernel");
  return 0;
}

int trx_softmax_bwd(const __nv_bfloat16* Patt, const float* dP, const int* cnt, __nv_bfloat16* dS, const TrxDims& s,
                    cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(s.B) * s.NqT;
  softmax_bwd_kernel<<<static_cast<unsigned>(ceil_div(rows * 32, 256)), 256, 0, st>>>(Patt, dP, cnt, dS, s);
  LMKD_LAUNCH_CHECK("softmax_bwd_kernel");
  return 0;
}

int trx_ln_bwd(const float* P, const float* bk, const float* gamma, const float* stats, const int* tuples,
               const int* slot, const float* dKq, const float* dKs, const float* dVs, const float* srow,
               const __nv_bfloat16* Dq, float* dxk, float* dxv, float* partials, int max_blocks, int* nblocks_out,
               const TrxDims& s, cudaStream_t st) {
  const size_t smem = sizeof(float) * (4 * s.d + kWarps * 2 * s.d);
  LMKD_CHECK(smem <= 200 * 1024, "trans_linear_out_dim %d too large", s.d);
  static bool attr_set = false;
  if (!attr_set) {
    LMKD_CUDA(cudaFuncSetAttribute(ln_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  int64_t blocks = ceil_div(s.R, kWarps);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 2;
  if (blocks > cap) blocks = cap;
  if (blocks > max_blocks) blocks = max_blocks;
  *nblocks_out = static_cast<int>(blocks);
  ln_bwd_kernel<<<static_cast<unsigned>(blocks), kWarps * 32, smem, st>>>(P, bk, gamma, stats, tuples, slot, dKq, dKs,
                                                                         dVs, srow, Dq, dxk, dxv, partials, s);
  LMKD_LAUNCH_CHECK("ln_bwd_kernel");
  return 0;
}

int trx_reduce_partials(const float* partials, int nblocks, float* ggamma, float* gbeta, float* gbk, float* gbv,
                        int d, cudaStream_t st) {
  reduce_partials_kernel<<<static_cast<unsigned>(ceil_div(4 * d, 128)), 128, 0, st>>>(partials, nblocks, ggamma, gbeta,
                                                                                     gbk, gbv, d);
  LMKD_LAUNCH_CHECK("reduce_partials_kernel");
  return 0;
}

int trx_tuple_gather_bwd(const float* dxk, const float* dxv, const int* inv_off, const int* inv_idx,
                         __nv_bfloat16* dPcat, const TrxDims& s, cudaStream_t st) {
  LMKD_CHECK(s.M < (1ll << 31), "too many frame rows");
  dim3 grid(static_cast<unsigned>(s.M), 2);
  tuple_gather_bwd_kernel<<<grid, 128, 0, st>>>(dxk, dxv, inv_off, inv_idx, dPcat, s);
  LMKD_LAUNCH_CHECK("tuple_gather_bwd_kernel");
  return 0;
}

int trx_pack_weights(const float* Wk, const float* Wv, __nv_bfloat16* Wcat, const TrxDims& s, cudaStream_t st) {
  const int64_t total = 2ll * s.card * s.d * s.D;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, st>>>(Wk, Wv, Wcat, s.d, s.D, s.card);
  LMKD_LAUNCH_CHECK("pack_weights_kernel");
  return 0;
}

int trx_unpack_wgrad(const float* dWcat, float* gWk, float* gWv, const TrxDims& s, cudaStream_t st) {
  const int64_t total = 2ll * s.card * s.d * s.D;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, st>>>(dWcat, gWk, gWv, s.d, s.D, s.card);
  LMKD_LAUNCH_CHECK("unpack_wgrad_kernel");
  return 0;
}

}  // namespace lmkd
