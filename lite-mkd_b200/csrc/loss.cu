// D2M loss kernels.
//  * feat_mse_*       : the HBM-bound one — a single streaming pass reads student and teacher
//                       features once and writes the feature gradient, 16-byte vector accesses,
//                       warp-shuffle partial sums, one partial per block (deterministic finish).
//  * d2m_logit_loss   : every logits-level term of a Distiller recipe (CE, temperature KL,
//                       inter-class relation, WSL focal weight) for one episode per block, values
//                       and gradients in one launch.
#include "loss.cuh"

namespace lmkd {

namespace {

constexpr int kMaxCols = 64;
constexpr int kLossThreads = 128;

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_u4(uint4* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

constexpr int kMseUnroll = 4;

__global__ void __launch_bounds__(256)
feat_mse_kernel(const float* __restrict__ s, const float* __restrict__ t, float* __restrict__ ds, int64_t n,
                float gscale, float* __restrict__ partials) {
  __shared__ float scratch[32];
  const int64_t n4 = n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(s);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  float4* d4 = reinterpret_cast<float4*>(ds);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (kMseUnroll - 1) * stride < n4; i += kMseUnroll * stride) {
    float4 a[kMseUnroll], b[kMseUnroll];
#pragma unroll
    for (int u = 0; u < kMseUnroll; ++u) a[u] = ldg_stream(s4 + i + u * stride);
#pragma unroll
    for (int u = 0; u < kMseUnroll; ++u) b[u] = ldg_stream(t4 + i + u * stride);
#pragma unroll
    for (int u = 0; u < kMseUnroll; ++u) {
      float4 d = make_float4(a[u].x - b[u].x, a[u].y - b[u].y, a[u].z - b[u].z, a[u].w - b[u].w);
      acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
      d.x *= gscale; d.y *= gscale; d.z *= gscale; d.w *= gscale;
      stg_stream(d4 + i + u * stride, d);
    }
  }
  for (; i < n4; i += stride) {
    const float4 a = ldg_stream(s4 + i), b = ldg_stream(t4 + i);
    float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    d.x *= gscale; d.y *= gscale; d.z *= gscale; d.w *= gscale;
    stg_stream(d4 + i, d);
  }
  if (blockIdx.x == 0) {   // ragged tail (n % 4)
    for (int64_t k = (n4 << 2) + threadIdx.x; k < n; k += blockDim.x) {
      const float d = s[k] - t[k];
      acc += d * d;
      ds[k] = gscale * d;
    }
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&w);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(256)
feat_mse_bf16_kernel(const __nv_bfloat16* __restrict__ s, const __nv_bfloat16* __restrict__ t,
                     __nv_bfloat16* __restrict__ ds, int64_t n, float gscale, float* __restrict__ partials) {
  __shared__ float scratch[32];
  const int64_t n8 = n >> 3;
  const uint4* s8 = reinterpret_cast<const uint4*>(s);
  const uint4* t8 = reinterpret_cast<const uint4*>(t);
  uint4* d8 = reinterpret_cast<uint4*>(ds);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (kMseUnroll - 1) * stride < n8; i += kMseUnroll * stride) {
    uint4 a[kMseUnroll], b[kMseUnroll];
#pragma unroll
    for (int u = 0; u < kMseUnroll; ++u) a[u] = ldg_stream_u4(s8 + i + u * stride);
#pragma unroll
    for (int u = 0; u < kMseUnroll; ++u) b[u] = ldg_stream_u4(t8 + i + u * stride);
#pragma unroll
    for (int u = 0; u < kMseUnroll; ++u) {
      const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, bw[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
      uint32_t ow[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x = bf2_to_f2(aw[k]), y = bf2_to_f2(bw[k]);
        const float d0 = x.x - y.x, d1 = x.y - y.y;
        acc += d0 * d0 + d1 * d1;
        ow[k] = f2_to_bf2(d0 * gscale, d1 * gscale);
      }
      stg_stream_u4(d8 + i + u * stride, make_uint4(ow[0], ow[1], ow[2], ow[3]));
    }
  }
  for (; i < n8; i += stride) {
    const uint4 a = ldg_stream_u4(s8 + i), b = ldg_stream_u4(t8 + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = bf2_to_f2(aw[k]), y = bf2_to_f2(bw[k]);
      const float d0 = x.x - y.x, d1 = x.y - y.y;
      acc += d0 * d0 + d1 * d1;
      ow[k] = f2_to_bf2(d0 * gscale, d1 * gscale);
    }
    stg_stream_u4(d8 + i, make_uint4(ow[0], ow[1], ow[2], ow[3]));
  }
  if (blockIdx.x == 0) {
    for (int64_t k = (n8 << 3) + threadIdx.x; k < n; k += blockDim.x) {
      const float d = __bfloat162float(s[k]) - __bfloat162float(t[k]);
      acc += d * d;
      ds[k] = __float2bfloat16_rn(gscale * d);
    }
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void mse_finish_kernel(const float* __restrict__ partials, int n, float lscale, float* __restrict__ out,
                                  int accumulate) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + lscale * acc;
}

__global__ void scale_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ g) {
  const float gv = *g;
  if (gv == 1.f) return;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    x[i] *= gv;
}

// ---- logits-level terms ---------------------------------------------------------------------
__device__ __forceinline__ float row_lse(const float* v, int C, float scale) {
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, v[c] * scale);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += __expf(v[c] * scale - mx);
  return mx + __logf(s);
}

// value of one row of a term; if g != nullptr also d(value)/d(s row) (unscaled by 1/rows)
__device__ float term_row(int kind, const float* s, const float* t, int64_t y, int C, float T, float* g) {
  if (kind == TERM_CE) {
    const float lse = row_lse(s, C, 1.f);
    if (g)
      for (int c = 0; c < C; ++c) g[c] = __expf(s[c] - lse) - (c == y ? 1.f : 0.f);
    return lse - s[y];
  }
  if (kind == TERM_KD) {
    const float invT = 1.f / T;
    const float ls = row_lse(s, C, invT), lt = row_lse(t, C, invT);
    float v = 0.f;
    for (int c = 0; c < C; ++c) {
      const float lps = s[c] * invT - ls, lpt = t[c] * invT - lt;
      const float pt = __expf(lpt);
      v += pt * (lpt - lps);
      if (g) g[c] = T * (__expf(lps) - pt);       // T^2 * (1/T) * (p_s - p_t)
    }
    return v * T * T;
  }
  // TERM_ICR: 1 - pearson(softmax(s), softmax(t)); returns -pearson (the +1 is added by the caller)
  float a[kMaxCols], b[kMaxCols];
  const float ls = row_lse(s, C, 1.f), lt = row_lse(t, C, 1.f);
  float ma = 0.f, mb = 0.f;
  for (int c = 0; c < C; ++c) {
    a[c] = __expf(s[c] - ls);
    b[c] = __expf(t[c] - lt);
    ma += a[c];
    mb += b[c];
  }
  ma /= C;
  mb /= C;
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (int c = 0; c < C; ++c) {
    const float ac = a[c] - ma, bc = b[c] - mb;
    dot += ac * bc;
    na += ac * ac;
    nb += bc * bc;
  }
  na = sqrtf(na);
  nb = sqrtf(nb);
  const float den = na * nb + 1e-8f;
  const float r = dot / den;
  if (g) {
    // dr/d(ac) = bc/den - dot*nb*ac/(den^2*na); centre; then through the softmax
    float gm = 0.f;
    for (int c = 0; c < C; ++c) {
      const float ac = a[c] - ma, bc = b[c] - mb;
      const float ga = bc / den - (na > 0.f ? dot * nb * ac / (den * den * na) : 0.f);
      g[c] = ga;
      gm += ga;
    }
    gm /= C;
    float sa = 0.f;
    for (int c = 0; c < C; ++c) {
      g[c] -= gm;
      sa += a[c] * g[c];
    }
    for (int c = 0; c < C; ++c) g[c] = -(a[c] * (g[c] - sa));   // d(-r)/ds
  }
  return -r;
}

__global__ void __launch_bounds__(kLossThreads)
d2m_logit_loss_kernel(const LossSpec spec, float* __restrict__ loss, float* __restrict__ values,
                      float* __restrict__ focal_out) {
  __shared__ float scratch[32];
  __shared__ float tv[kMaxTerms];
  const int64_t b = blockIdx.x;
  const float T = spec.temperature;
  // phase A: unweighted term values
  for (int i = 0; i < spec.nterms; ++i) {
    const LossTerm& tm = spec.terms[i];
    float acc = 0.f;
    for (int r = threadIdx.x; r < tm.rows; r += blockDim.x) {
      const int64_t off = (b * tm.rows + r) * tm.cols;
      acc += term_row(tm.kind, tm.s + off, tm.t ? tm.t + off : nullptr, tm.y ? tm.y[b * tm.rows + r] : 0,
                      tm.cols, T, nullptr);
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) tv[i] = acc / tm.rows + (tm.kind == TERM_ICR ? 1.f : 0.f);
  }
  // focal weight (WSL family, distillers.py:86-93): detached CE ratio
  float fw = 0.f;
  if (spec.fnum != nullptr) {
    float a = 0.f, d = 0.f;
    for (int r = threadIdx.x; r < spec.frows; r += blockDim.x) {
      const int64_t off = (b * spec.frows + r) * spec.fcols;
      const int64_t y = spec.fy[b * spec.frows + r];
      a += term_row(TERM_CE, spec.fnum + off, nullptr, y, spec.fcols, T, nullptr);
      d += term_row(TERM_CE, spec.fden + off, nullptr, y, spec.fcols, T, nullptr);
    }
    a = block_sum(a, scratch) / spec.frows;
    d = block_sum(d, scratch) / spec.frows;
    fw = 1.f - __expf(-fmaxf(a / (d + 1e-8f), 0.f));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < spec.nterms; ++i) {
      const LossTerm& tm = spec.terms[i];
      tot += tm.w * (tm.fa + tm.fb * fw) * tv[i];
      if (values) values[b * spec.nterms + i] = tv[i];
    }
    loss[b] = tot;
    if (focal_out) focal_out[b] = fw;
  }
  // phase B: gradients (each thread owns the same rows for every term, so += is race-free)
  for (int i = 0; i < spec.nterms; ++i) {
    const LossTerm& tm = spec.terms[i];
    if (tm.grad == nullptr) continue;
    const float coef = tm.w * (tm.fa + tm.fb * fw) / tm.rows;
    float g[kMaxCols];
    for (int r = threadIdx.x; r < tm.rows; r += blockDim.x) {
      const int64_t off = (b * tm.rows + r) * tm.cols;
      term_row(tm.kind, tm.s + off, tm.t ? tm.t + off : nullptr, tm.y ? tm.y[b * tm.rows + r] : 0, tm.cols, T, g);
      float* dst = tm.grad + off;
      for (int c = 0; c < tm.cols; ++c) dst[c] = (tm.grad_accumulate ? dst[c] : 0.f) + coef * g[c];
    }
  }
}

// ---- SupportDK ----------------------------------------------------------------------------------
__global__ void protos_kernel(const float* __restrict__ support, float* __restrict__ protos, int64_t total4,
                              int shot, int64_t vid4) {
  // protos[(b, i)][e] = mean_k support[(b, i, k)][e]; vid4 = L*D/4
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total4;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = t % vid4, bi = t / vid4;
    const float4* src = reinterpret_cast<const float4*>(support) + bi * shot * vid4 + e;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < shot; ++k) {
      const float4 v = __ldg(src + k * vid4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float inv = 1.f / shot;
    reinterpret_cast<float4*>(protos)[t] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

// block per (b, unordered pair i < n)
__global__ void __launch_bounds__(256)
support_dk_pairs_kernel(const float* __restrict__ protos, float* __restrict__ out, int way, int L, int64_t vid) {
  __shared__ float scratch[32];
  const int npairs = way * (way - 1) / 2;
  const int64_t b = blockIdx.x / npairs;
  int pr = blockIdx.x % npairs, i = 0;
  while (pr >= way - 1 - i) { pr -= way - 1 - i; ++i; }
  const int n = i + 1 + pr;
  const float4* pi = reinterpret_cast<const float4*>(protos + (b * way + i) * vid);
  const float4* pn = reinterpret_cast<const float4*>(protos + (b * way + n) * vid);
  float acc = 0.f;
  for (int64_t e = threadIdx.x; e < vid / 4; e += blockDim.x) {
    const float4 x = __ldg(pi + e), y = __ldg(pn + e);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    const float v = -acc / L;
    float* o = out + b * way * (way - 1);
    o[i * (way - 1) + (n - 1)] = v;   // row i, n > i sits at column n-1
    o[n * (way - 1) + i] = v;         // row n, i < n sits at column i
  }
}

__global__ void support_dk_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ protos,
                                      float* __restrict__ gsupport, int64_t total4, int way, int shot, int L,
                                      int64_t vid4) {
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total4;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = t % vid4, bi = t / vid4;
    const int i = static_cast<int>(bi % way);
    const int64_t b = bi / way;
    const float4 pi = __ldg(reinterpret_cast<const float4*>(protos) + bi * vid4 + e);
    const float* g = gout + b * way * (way - 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = 0; n < way; ++n) {
      if (n == i) continue;
      const float gi = g[i * (way - 1) + (n > i ? n - 1 : n)] + g[n * (way - 1) + (i > n ? i - 1 : i)];
      const float4 pn = __ldg(reinterpret_cast<const float4*>(protos) + (b * way + n) * vid4 + e);
      const float c = -2.f * gi / L;
      acc.x += c * (pi.x - pn.x); acc.y += c * (pi.y - pn.y); acc.z += c * (pi.z - pn.z); acc.w += c * (pi.w - pn.w);
    }
    const float inv = 1.f / shot;
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    float4* dst = reinterpret_cast<float4*>(gsupport) + bi * shot * vid4 + e;
    for (int k = 0; k < shot; ++k) dst[k * vid4] = acc;
  }
}

__global__ void accuracy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t rows,
                                int cols, int* __restrict__ correct) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int hit = 0;
  if (r < rows) {
    const float* v = logits + r * cols;
    int best = 0;
    for (int c = 1; c < cols; ++c)
      if (v[c] > v[best]) best = c;
    hit = (best == labels[r]);
  }
  const unsigned m = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(correct, __popc(m));
}

int stream_blocks() { return sm_count() * 8; }

}  // namespace

int feat_mse_fwdbwd(const float* s, const float* t, float* ds, int64_t n, float gscale, float* partials,
                    int max_partials, int* npartials, cudaStream_t st) {
  LMKD_CHECK(n > 0, "feat_mse: empty input");
  LMKD_CHECK(((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(ds)) & 15) == 0,
             "feat_mse: pointers must be 16-byte aligned");
  int64_t blocks = ceil_div(n / 4 + 1, 256);
  if (blocks > stream_blocks()) blocks = stream_blocks();
  if (blocks > max_partials) blocks = max_partials;
  *npartials = static_cast<int>(blocks);
  KernelTimingScope timing(TIME_LOSS, st, 3.0 * n * sizeof(float));      // read s, read t, write ds
  if (int rc = timing.begin()) return rc;
  feat_mse_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(s, t, ds, n, gscale, partials);
  LMKD_LAUNCH_CHECK("feat_mse_kernel");
  if (int rc = timing.end()) return rc;
  return 0;
}

int feat_mse_fwdbwd_bf16(const __nv_bfloat16* s, const __nv_bfloat16* t, __nv_bfloat16* ds, int64_t n, float gscale,
                         float* partials, int max_partials, int* npartials, cudaStream_t st) {
  LMKD_CHECK(n > 0, "feat_mse: empty input");
  LMKD_CHECK(((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(ds)) & 15) == 0,
             "feat_mse: pointers must be 16-byte aligned");
  int64_t blocks = ceil_div(n / 8 + 1, 256);
  if (blocks > stream_blocks()) blocks = stream_blocks();
  if (blocks > max_partials) blocks = max_partials;
  *npartials = static_cast<int>(blocks);
  KernelTimingScope timing(TIME_LOSS, st, 3.0 * n * sizeof(__nv_bfloat16));
  if (int rc = timing.begin()) return rc;
  feat_mse_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(s, t, ds, n, gscale, partials);
  LMKD_LAUNCH_CHECK("feat_mse_bf16_kernel");
  if (int rc = timing.end()) return rc;
  return 0;
}

int mse_finish(const float* partials, int npartials, float lscale, float* loss_out, int accumulate, cudaStream_t st) {
  mse_finish_kernel<<<1, 256, 0, st>>>(partials, npartials, lscale, loss_out, accumulate);
  LMKD_LAUNCH_CHECK("mse_finish_kernel");
  return 0;
}

int scale_by_device_scalar(float* x, int64_t n, const float* g, cudaStream_t st) {
  int64_t blocks = ceil_div(n, 256);
  if (blocks > stream_blocks()) blocks = stream_blocks();
  scale_kernel<<<static_cast<unsigned>(blocks > 0 ? blocks : 1), 256, 0, st>>>(x, n, g);
  LMKD_LAUNCH_CHECK("scale_kernel");
  return 0;
}

int d2m_logit_loss(const LossSpec& spec, int B, float* loss, float* values, float* focal, cudaStream_t st) {
  LMKD_CHECK(spec.nterms >= 1 && spec.nterms <= kMaxTerms, "d2m: %d terms (max %d)", spec.nterms, kMaxTerms);
  for (int i = 0; i < spec.nterms; ++i) {
    const LossTerm& t = spec.terms[i];
    LMKD_CHECK(t.cols >= 1 && t.cols <= kMaxCols && t.rows >= 1, "d2m: term %d has %d x %d logits (cols <= %d)", i,
               t.rows, t.cols, kMaxCols);
    LMKD_CHECK(t.s != nullptr, "d2m: term %d has no student logits", i);
    LMKD_CHECK(t.kind == TERM_CE ? t.y != nullptr : t.t != nullptr, "d2m: term %d lacks its target", i);
  }
  if (spec.fnum) LMKD_CHECK(spec.fden && spec.fy && spec.fcols <= kMaxCols, "d2m: incomplete focal spec");
  d2m_logit_loss_kernel<<<B, kLossThreads, 0, st>>>(spec, loss, values, focal);
  LMKD_LAUNCH_CHECK("d2m_logit_loss_kernel");
  return 0;
}

int support_dk_fwd(const float* support, float* protos, float* out, int B, int way, int shot, int L, int D,
                   cudaStream_t st) {
  LMKD_CHECK(D % 4 == 0 && way >= 2, "support_dk: D %% 4 and way >= 2 required");
  const int64_t vid = static_cast<int64_t>(L) * D;
  const int64_t total4 = static_cast<int64_t>(B) * way * vid / 4;
  int64_t blocks = ceil_div(total4, 256);
  if (blocks > stream_blocks()) blocks = stream_blocks();
  protos_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(support, protos, total4, shot, vid / 4);
  LMKD_LAUNCH_CHECK("protos_kernel");
  support_dk_pairs_kernel<<<static_cast<unsigned>(B * way * (way - 1) / 2), 256, 0, st>>>(protos, out, way, L, vid);
  LMKD_LAUNCH_CHECK("support_dk_pairs_kernel");
  return 0;
}

int support_dk_bwd(const float* gout, const float* protos, float* gsupport, int B, int way, int shot, int L, int D,
                   cudaStream_t st) {
  const int64_t vid = static_cast<int64_t>(L) * D;
  const int64_t total4 = static_cast<int64_t>(B) * way * vid / 4;
  int64_t blocks = ceil_div(total4, 256);
  if (blocks > stream_blocks()) blocks = stream_blocks();
  support_dk_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(gout, protos, gsupport, total4, way, shot, L,
                                                                      vid / 4);
  LMKD_LAUNCH_CHECK("support_dk_bwd_kernel");
  return 0;
}

int accuracy_count(const float* logits, const int64_t* labels, int64_t rows, int cols, int* correct,
                   cudaStream_t st) {
  accuracy_kernel<<<static_cast<unsigned>(ceil_div(rows, 128)), 128, 0, st>>>(logits, labels, rows, cols, correct);
  LMKD_LAUNCH_CHECK("accuracy_kernel");
  return 0;
}

}  // namespace lmkd
