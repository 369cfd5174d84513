// STRM DistanceLoss head, non-GEMM kernels (see strm.cuh).  These are the straightforward versions (block per video
// or per row, L2-resident gathers): the head is a widening row, its contractions -- the
// factored tuple MLP and the tuple-to-tuple distance matrix with the arg-min in the epilogue -- run on the tcgen05 GEMM.
#include "strm.cuh"

namespace lmkd {

namespace {

constexpr int kWarps = 8;
constexpr float kHuge = 1.0e30f;

__global__ void __launch_bounds__(256)
strm_pack_weight_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wcat, int dm, int D, int card) {
  const int r = blockIdx.x;                          // (j, i)
  const int i = r % dm, j = r / dm;
  const float4* src = reinterpret_cast<const float4*>(W + static_cast<int64_t>(i) * card * D + static_cast<int64_t>(j) * D);
  uint2* dst = reinterpret_cast<uint2*>(Wcat + static_cast<int64_t>(r) * D);
  for (int c4 = threadIdx.x; c4 < D / 4; c4 += blockDim.x) {
    const float4 v = __ldg(src + c4);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    dst[c4] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
}

__global__ void __launch_bounds__(256)
strm_unpack_wgrad_kernel(const float* __restrict__ dWcat, float* __restrict__ gW, int dm, int D, int card) {
  const int r = blockIdx.x;
  const int i = r % dm, j = r / dm;
  const float4* src = reinterpret_cast<const float4*>(dWcat + static_cast<int64_t>(r) * D);
  float4* dst = reinterpret_cast<float4*>(gW + static_cast<int64_t>(i) * card * D + static_cast<int64_t>(j) * D);
  for (int c4 = threadIdx.x; c4 < D / 4; c4 += blockDim.x) dst[c4] = __ldg(src + c4);
}

// block = one video (b, n); warps loop over its T tuples
__global__ void __launch_bounds__(kWarps * 32)
strm_tuple_relu_fwd_kernel(const float* __restrict__ P, const float* __restrict__ bias, const int* __restrict__ tuples,
                           const int* __restrict__ slot, __nv_bfloat16* __restrict__ Eq, __nv_bfloat16* __restrict__ Es,
                           float* __restrict__ nq2, float* __restrict__ ns2, const TrxDims s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t vid = blockIdx.x;
  const int n = static_cast<int>(vid % s.N);
  const int64_t b = vid / s.N;
  const int64_t pcols = static_cast<int64_t>(s.card) * s.d;
  const float* Pv = P + vid * s.L * pcols;
  __nv_bfloat16* Ed;
  float* nd;
  int64_t out_row;
  if (n < s.Ns) {
    const int sl = slot[b * s.Ns + n];
    if (sl < 0) return;                                  // dropped support (label out of range / class overfull)
    out_row = (b * s.way + sl / s.shot) * s.KTp + static_cast<int64_t>(sl % s.shot) * s.T;
    Ed = Es; nd = ns2;
  } else {
    out_row = b * s.NqT + static_cast<int64_t>(n - s.Ns) * s.T;
    Ed = Eq; nd = nq2;
  }
  for (int tau = warp; tau < s.T; tau += kWarps) {
    const int* tp = tuples + tau * s.card;
    float sq = 0.f;
    __nv_bfloat16* ed = Ed + (out_row + tau) * s.d;
    for (int i = lane * 2; i < s.d; i += 64) {           // d is a multiple of 8: pairs never straddle the end
      float x0 = __ldg(bias + i), x1 = __ldg(bias + i + 1);
      for (int j = 0; j < s.card; ++j) {
        const float2 pv = __ldg(reinterpret_cast<const float2*>(Pv + tp[j] * pcols + static_cast<int64_t>(j) * s.d + i));
        x0 += pv.x; x1 += pv.y;
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
      const float2 r = __bfloat1622float2(h);             // the norm of what the distance GEMM will actually see
      sq = fmaf(r.x, r.x, fmaf(r.y, r.y, sq));
      *reinterpret_cast<__nv_bfloat162*>(ed + i) = h;
    }
    sq = warp_sum(sq);
    if (lane == 0) nd[out_row + tau] = sq;
  }
}

// rows [cnt*T, KTp) of every (b, class) block: zero embeddings, +huge norm
__global__ void __launch_bounds__(128)
strm_pad_rows_kernel(const int* __restrict__ cnt, __nv_bfloat16* __restrict__ Es, float* __restrict__ ns2, int T, int KTp,
                     int d) {
  const int64_t bc = blockIdx.x;
  const int first = cnt[bc] * T;
  const int n8 = (KTp - first) * d / 8;
  uint4* e = reinterpret_cast<uint4*>(Es + (bc * KTp + first) * d);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < n8; i += blockDim.x) e[i] = z;
  for (int i = first + threadIdx.x; i < KTp; i += blockDim.x) ns2[bc * KTp + i] = kHuge;
}

__global__ void strm_logits_fwd_kernel(const unsigned long long* __restrict__ best, const int* __restrict__ cnt,
                                       float* __restrict__ logits, const TrxDims s) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // (b, q, c)
  if (i >= static_cast<int64_t>(s.B) * s.Nq * s.way) return;
  const int c = static_cast<int>(i % s.way);
  const int q = static_cast<int>((i / s.way) % s.Nq);
  const int64_t b = i / (static_cast<int64_t>(s.way) * s.Nq);
  float acc = 0.f;
  if (cnt[b * s.way + c] > 0) {
    const unsigned long long* src = best + (b * s.way + c) * s.NqT + static_cast<int64_t>(q) * s.T;
    for (int t = 0; t < s.T; ++t) acc += sqrtf(__uint_as_float(static_cast<unsigned int>(src[t] >> 32)));
    acc = -acc / s.T;
  }
  logits[i] = acc;     // classes without supports keep logit 0 (zero-initialised dist_all, :212)
}

// warp per query tuple row (b, m): d logit[q][c] / d E = -(1/T) (Eq - Es*) / |Eq - Es*| for the arg-min support tuple
constexpr int kDistPairs = 18;     // bf16 pairs per lane kept in registers by strm_dist_bwd: rows up to 1152 wide
__global__ void __launch_bounds__(kWarps * 32)
strm_dist_bwd_kernel(const float* __restrict__ glogits, const unsigned long long* __restrict__ best,
                     const int* __restrict__ cnt, const __nv_bfloat16* __restrict__ Eq, const __nv_bfloat16* __restrict__ Es,
                     float* __restrict__ dEq, float* __restrict__ dEs, const TrxDims s) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= static_cast<int64_t>(s.B) * s.NqT) return;
  const int64_t b = row / s.NqT;
  const int m = static_cast<int>(row % s.NqT);
  const int q = m / s.T;
  const __nv_bfloat162* eq = reinterpret_cast<const __nv_bfloat162*>(Eq + row * s.d);
  float2* out = reinterpret_cast<float2*>(dEq + row * s.d);
  const int np = s.d >> 1;                               // bf16 pairs per row
  if (np <= 32 * kDistPairs) {
    // the query row, its gradient and the difference to the nearest support row stay in registers: one read of each
    // row, one write of dEq, vector reductions into dEs (the first version re-read both rows for the gradient, updated
    // dEq through global memory once per class and issued two scalar atomics per pair: 531 us for 64 episodes)
    float2 a[kDistPairs], acc[kDistPairs];
#pragma unroll
    for (int k = 0; k < kDistPairs; ++k) {
      const int i = lane + 32 * k;
      a[k] = i < np ? __bfloat1622float2(eq[i]) : make_float2(0.f, 0.f);
      acc[k] = make_float2(0.f, 0.f);
    }
    for (int c = 0; c < s.way; ++c) {
      if (cnt[b * s.way + c] <= 0) continue;
      const unsigned long long key = best[(b * s.way + c) * s.NqT + m];
      const int col = static_cast<int>(key & 0xffffffffu);
      const float g = glogits[(b * s.Nq + q) * s.way + c];
      if (g == 0.f || col >= s.KTp) continue;
      const int64_t srow = (b * s.way + c) * s.KTp + col;
      const __nv_bfloat162* es = reinterpret_cast<const __nv_bfloat162*>(Es + srow * s.d);
      float2 df[kDistPairs];
      float d2 = 0.f;
#pragma unroll
      for (int k = 0; k < kDistPairs; ++k) {
        const int i = lane + 32 * k;
        const float2 e = i < np ? __bfloat1622float2(es[i]) : make_float2(0.f, 0.f);
        df[k] = make_float2(a[k].x - e.x, a[k].y - e.y);
        d2 = fmaf(df[k].x, df[k].x, fmaf(df[k].y, df[k].y, d2));
      }
      // the distance itself is recomputed from the two rows: |a|^2 + |b|^2 - 2<a,b> (what selected the arg-min) loses
      // digits to cancellation exactly where the rows are close, and 1/dist scales the whole gradient
      const float dist = sqrtf(warp_sum(d2));
      if (!(dist > 0.f)) continue;                       // torch.cdist backward is 0 at zero distance
      const float coef = -g / (s.T * dist);
      float2* des = reinterpret_cast<float2*>(dEs + srow * s.d);
#pragma unroll
      for (int k = 0; k < kDistPairs; ++k) {
        const int i = lane + 32 * k;
        if (i < np) {
          const float dx = coef * df[k].x, dy = coef * df[k].y;
          acc[k].x += dx;
          acc[k].y += dy;
          atomicAdd(des + i, make_float2(-dx, -dy));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kDistPairs; ++k) {
      const int i = lane + 32 * k;
      if (i < np) out[i] = acc[k];
    }
    return;
  }
  for (int i = lane; i < s.d / 2; i += 32) out[i] = make_float2(0.f, 0.f);
  __syncwarp();
  for (int c = 0; c < s.way; ++c) {
    if (cnt[b * s.way + c] <= 0) continue;
    const unsigned long long key = best[(b * s.way + c) * s.NqT + m];
    const int col = static_cast<int>(key & 0xffffffffu);
    const float g = glogits[(b * s.Nq + q) * s.way + c];
    if (g == 0.f || col >= s.KTp) continue;
    const int64_t srow = (b * s.way + c) * s.KTp + col;
    const __nv_bfloat162* es = reinterpret_cast<const __nv_bfloat162*>(Es + srow * s.d);
    // the distance itself is recomputed from the two rows: |a|^2 + |b|^2 - 2<a,b> (what selected the arg-min) loses
    // digits to cancellation exactly where the rows are close, and 1/dist scales the whole gradient
    float d2 = 0.f;
    for (int i = lane; i < s.d / 2; i += 32) {
      const float2 a = __bfloat1622float2(eq[i]), e = __bfloat1622float2(es[i]);
      d2 = fmaf(a.x - e.x, a.x - e.x, fmaf(a.y - e.y, a.y - e.y, d2));
    }
    const float dist = sqrtf(warp_sum(d2));
    if (!(dist > 0.f)) continue;                         // torch.cdist backward is 0 at zero distance
    const float coef = -g / (s.T * dist);
    float* des = dEs + srow * s.d;
    for (int i = lane; i < s.d / 2; i += 32) {
      const float2 a = __bfloat1622float2(eq[i]), e = __bfloat1622float2(es[i]);
      const float dx = coef * (a.x - e.x), dy = coef * (a.y - e.y);
      float2 o = out[i];
      o.x += dx; o.y += dy;
      out[i] = o;
      atomicAdd(des + 2 * i, -dx);
      atomicAdd(des + 2 * i + 1, -dy);
    }
  }
}

// block per frame row (b, n, l); threads over dm.  Column block j of the row = sum over the tuples whose j-th
// frame is l of relu'(E) * dE; the bias gradient takes each tuple once (through its first frame).
__global__ void __launch_bounds__(128)
strm_relu_gather_bwd_kernel(const float* __restrict__ dEq, const float* __restrict__ dEs,
                            const __nv_bfloat16* __restrict__ Eq, const __nv_bfloat16* __restrict__ Es,
                            const int* __restrict__ slot, const int* __restrict__ inv_off, const int* __restrict__ inv_idx,
                            __nv_bfloat16* __restrict__ dPcat, float* __restrict__ gbias, const TrxDims s) {
  const int64_t frow = blockIdx.x;
  const int l = static_cast<int>(frow % s.L);
  const int64_t vid = frow / s.L;
  const int n = static_cast<int>(vid % s.N);
  const int64_t b = vid / s.N;
  const int64_t pcols = static_cast<int64_t>(s.card) * s.d;
  const float* dE = nullptr;
  const __nv_bfloat16* E = nullptr;
  if (n < s.Ns) {
    const int sl = slot[b * s.Ns + n];
    if (sl >= 0) {
      const int64_t r0 = (b * s.way + sl / s.shot) * s.KTp + static_cast<int64_t>(sl % s.shot) * s.T;
      dE = dEs + r0 * s.d;
      E = Es + r0 * s.d;
    }
  } else {
    const int64_t r0 = b * s.NqT + static_cast<int64_t>(n - s.Ns) * s.T;
    dE = dEq + r0 * s.d;
    E = Eq + r0 * s.d;
  }
  for (int j = 0; j < s.card; ++j) {
    const int beg = inv_off[j * s.L + l], end = inv_off[j * s.L + l + 1];
    __nv_bfloat16* dst = dPcat + frow * pcols + static_cast<int64_t>(j) * s.d;
    for (int i = threadIdx.x; i < s.d; i += blockDim.x) {
      float acc = 0.f;
      if (dE != nullptr)
        for (int e = beg; e < end; ++e) {
          const int64_t o = static_cast<int64_t>(inv_idx[e]) * s.d + i;
          if (__bfloat162float(E[o]) > 0.f) acc += dE[o];
        }
      dst[i] = __float2bfloat16_rn(acc);
      if (j == 0 && acc != 0.f) atomicAdd(gbias + i, acc);
    }
  }
}

}  // namespace

int strm_pack_weight(const float* W, __nv_bfloat16* Wcat, const TrxDims& s, cudaStream_t st) {
  strm_pack_weight_kernel<<<s.card * s.d, 256, 0, st>>>(W, Wcat, s.d, s.D, s.card);
  LMKD_LAUNCH_CHECK("strm_pack_weight_kernel");
  return 0;
}

int strm_unpack_wgrad(const float* dWcat, float* gW, const TrxDims& s, cudaStream_t st) {
  strm_unpack_wgrad_kernel<<<s.card * s.d, 256, 0, st>>>(dWcat, gW, s.d, s.D, s.card);
  LMKD_LAUNCH_CHECK("strm_unpack_wgrad_kernel");
  return 0;
}

int strm_tuple_relu_fwd(const float* P, const float* bias, const int* tuples, const int* slot, const int* cnt,
                        __nv_bfloat16* Eq, __nv_bfloat16* Es, float* nq2, float* ns2, const TrxDims& s, cudaStream_t st) {
  strm_pad_rows_kernel<<<static_cast<unsigned>(static_cast<int64_t>(s.B) * s.way), 128, 0, st>>>(cnt, Es, ns2, s.T, s.KTp, s.d);
  LMKD_LAUNCH_CHECK("strm_pad_rows_kernel");
  strm_tuple_relu_fwd_kernel<<<static_cast<unsigned>(static_cast<int64_t>(s.B) * s.N), kWarps * 32, 0, st>>>(
      P, bias, tuples, slot, Eq, Es, nq2, ns2, s);
  LMKD_LAUNCH_CHECK("strm_tuple_relu_fwd_kernel");
  return 0;
}

int strm_logits_fwd(const unsigned long long* best, const int* cnt, float* logits, const TrxDims& s, cudaStream_t st) {
  const int64_t n = static_cast<int64_t>(s.B) * s.Nq * s.way;
  strm_logits_fwd_kernel<<<static_cast<unsigned>(ceil_div(n, 128)), 128, 0, st>>>(best, cnt, logits, s);
  LMKD_LAUNCH_CHECK("strm_logits_fwd_kernel");
  return 0;
}

int strm_dist_bwd(const float* glogits, const unsigned long long* best, const int* cnt, const __nv_bfloat16* Eq,
                  const __nv_bfloat16* Es, float* dEq, float* dEs, const TrxDims& s, cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(s.B) * s.NqT;
  strm_dist_bwd_kernel<<<static_cast<unsigned>(ceil_div(rows, kWarps)), kWarps * 32, 0, st>>>(glogits, best, cnt, Eq, Es,
                                                                                          dEq, dEs, s);
  LMKD_LAUNCH_CHECK("strm_dist_bwd_kernel");
  return 0;
}

int strm_relu_gather_bwd(const float* dEq, const float* dEs, const __nv_bfloat16* Eq, const __nv_bfloat16* Es,
                         const int* slot, const int* inv_off, const int* inv_idx, __nv_bfloat16* dPcat, float* gbias,
                         const TrxDims& s, cudaStream_t st) {
  LMKD_CHECK(s.M < (1ll << 31), "too many frame rows");
  strm_relu_gather_bwd_kernel<<<static_cast<unsigned>(s.M), 128, 0, st>>>(dEq, dEs, Eq, Es, slot, inv_off, inv_idx, dPcat,
                                                                          gbias, s);
  LMKD_LAUNCH_CHECK("strm_relu_gather_bwd_kernel");
  return 0;
}

}  // namespace lmkd
