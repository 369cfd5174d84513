#include "prep.cuh"

namespace lmkd {

namespace {

// one warp per row; D % 4 == 0
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// warp per feature row; 8 independent 16-byte loads per lane are in flight before the first is consumed
// (4 KB per warp), the fp32 input is read once and bypasses L1
__global__ void __launch_bounds__(256)
feat_cast_norm_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                      float* __restrict__ norms, int* __restrict__ nanflag, int64_t rows,
                      int D, int64_t rows_per_flag) {
  constexpr int U = 8;
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(x + row * D);
  uint2* dst = reinterpret_cast<uint2*>(xb + row * D);
  const int D4 = D >> 2;
  float acc = 0.f;
  for (int i0 = lane; i0 < D4; i0 += 32 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + 32 * u;
      v[u] = i < D4 ? ld_stream_f4(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + 32 * u;
      if (i < D4) {
        acc += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
        __nv_bfloat162 a = __floats2bfloat162_rn(v[u].x, v[u].y), b = __floats2bfloat162_rn(v[u].z, v[u].w);
        dst[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    norms[row] = sqrtf(acc);
    if (nanflag != nullptr && isnan(acc)) atomicOr(nanflag + row / rows_per_flag, 1);
  }
}

// one block per frame row (b, n, l): the row decomposition is done once per block, threads stride
// over the D/4 float4 columns
__global__ void __launch_bounds__(256)
trx_pe_cast_kernel(const float* __restrict__ support, const float* __restrict__ query,
                   const float* __restrict__ pe, __nv_bfloat16* __restrict__ out, int Ns, int Nq, int L, int D,
                   float p, float inv_keep, uint64_t seed_host, const uint64_t* __restrict__ seed_dev,
                   uint64_t* __restrict__ seed_used) {
  const uint64_t seed = seed_host + (seed_dev != nullptr ? *seed_dev : 0ull);
  if (blockIdx.x == 0 && threadIdx.x == 0 && seed_used != nullptr) *seed_used = seed;
  const int D4 = D >> 2;
  const int N = Ns + Nq;
  const int64_t row = blockIdx.x;                   // (b, n, l)
  const int l = static_cast<int>(row % L);
  const int64_t bn = row / L;
  const int n = static_cast<int>(bn % N);
  const int64_t b = bn / N;
  const float4* src = reinterpret_cast<const float4*>(
      n < Ns ? support + ((b * Ns + n) * L + l) * static_cast<int64_t>(D)
             : query + ((b * Nq + (n - Ns)) * L + l) * static_cast<int64_t>(D));
  const float4* pe4 = reinterpret_cast<const float4*>(pe + static_cast<int64_t>(l) * D);
  uint2* dst = reinterpret_cast<uint2*>(out + row * D);
  const uint64_t base = static_cast<uint64_t>(row) * D;
  for (int c4 = threadIdx.x; c4 < D4; c4 += blockDim.x) {
    float4 v = __ldg(src + c4);
    const float4 e = __ldg(pe4 + c4);
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    if (p > 0.f) {
      const uint64_t i0 = base + 4ull * c4;
      float sc[4];
      dropout_scale4(seed, i0 >> 2, dropout_threshold(p), inv_keep, sc);
      v.x *= sc[0]; v.y *= sc[1]; v.z *= sc[2]; v.w *= sc[3];
    }
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), bb = __floats2bfloat162_rn(v.z, v.w);
    dst[c4] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&bb));
  }
}

__global__ void __launch_bounds__(256)
trx_dx_scatter_kernel(const float* __restrict__ dx, float* __restrict__ gs, float* __restrict__ gq, int Ns, int Nq,
                      int L, int D, float p, float inv_keep, const uint64_t* __restrict__ seed_used, int accumulate) {
  const uint64_t seed = p > 0.f ? *seed_used : 0ull;
  const int D4 = D >> 2;
  const int N = Ns + Nq;
  const int64_t row = blockIdx.x;
  const int l = static_cast<int>(row % L);
  const int64_t bn = row / L;
  const int n = static_cast<int>(bn % N);
  const int64_t b = bn / N;
  float4* dst = reinterpret_cast<float4*>(n < Ns ? gs + ((b * Ns + n) * L + l) * static_cast<int64_t>(D)
                                                 : gq + ((b * Nq + (n - Ns)) * L + l) * static_cast<int64_t>(D));
  const float4* src = reinterpret_cast<const float4*>(dx + row * D);
  const uint64_t base = static_cast<uint64_t>(row) * D;
  for (int c4 = threadIdx.x; c4 < D4; c4 += blockDim.x) {
    float4 v = __ldg(src + c4);
    if (p > 0.f) {
      const uint64_t i0 = base + 4ull * c4;
      float sc[4];
      dropout_scale4(seed, i0 >> 2, dropout_threshold(p), inv_keep, sc);
      v.x *= sc[0]; v.y *= sc[1]; v.z *= sc[2]; v.w *= sc[3];
    }
    if (accumulate) {
      const float4 o = dst[c4];
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    dst[c4] = v;
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

// bf16 -> fp32, 8 elements per thread (16-byte load, two 16-byte stores); n8 = n / 8
__global__ void __launch_bounds__(256)
upcast_bf16_kernel(const uint4* __restrict__ x, float4* __restrict__ y, int64_t n8) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 q = __ldg(x + i);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.z));
    const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.w));
    y[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    y[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
  }
}

__global__ void dropout_mask_kernel(float* out, int64_t n, float p, float inv_keep, uint64_t seed) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = p > 0.f ? dropout_scale(seed, static_cast<uint64_t>(i), p, inv_keep) : 1.f;
}

int stream_grid(int64_t work_items, int threads) {
  int64_t blocks = ceil_div(work_items, threads);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace

int feat_cast_norm(const float* x, __nv_bfloat16* xb, float* norms, int* nanflag, int64_t rows, int D,
                   int64_t rows_per_flag, cudaStream_t stream) {
  LMKD_CHECK(D % 8 == 0, "feature dim %d must be a multiple of 8", D);
  const int threads = 256;
  const int64_t blocks = ceil_div(rows * 32, threads);
  feat_cast_norm_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(x, xb, norms, nanflag, rows, D,
                                                                              rows_per_flag);
  LMKD_LAUNCH_CHECK("feat_cast_norm_kernel");
  return 0;
}

int trx_pe_cast(const float* support, const float* query, const float* pe, __nv_bfloat16* out, int B, int Ns,
                int Nq, int L, int D, float p, uint64_t seed, const uint64_t* seed_dev, uint64_t* seed_used,
                cudaStream_t stream) {
  LMKD_CHECK(D % 8 == 0, "feature dim %d must be a multiple of 8", D);
  LMKD_CHECK(p >= 0.f && p < 1.f, "dropout p %f out of range", p);
  const int64_t rows = static_cast<int64_t>(B) * (Ns + Nq) * L;
  LMKD_CHECK(rows < (1ll << 31), "too many frame rows");
  const int threads = D / 4 >= 256 ? 256 : (D / 4 >= 128 ? 128 : 64);
  trx_pe_cast_kernel<<<static_cast<unsigned>(rows), threads, 0, stream>>>(support, query, pe, out, Ns, Nq, L, D, p,
                                                                         1.f / (1.f - p), seed, seed_dev, seed_used);
  LMKD_LAUNCH_CHECK("trx_pe_cast_kernel");
  return 0;
}

int trx_dx_scatter(const float* dx, float* gsupport, float* gquery, int B, int Ns, int Nq, int L, int D, float p,
                   const uint64_t* seed_used, int accumulate, cudaStream_t stream) {
  const int64_t rows = static_cast<int64_t>(B) * (Ns + Nq) * L;
  LMKD_CHECK(rows < (1ll << 31), "too many frame rows");
  const int threads = D / 4 >= 256 ? 256 : (D / 4 >= 128 ? 128 : 64);
  trx_dx_scatter_kernel<<<static_cast<unsigned>(rows), threads, 0, stream>>>(dx, gsupport, gquery, Ns, Nq, L, D, p,
                                                                            1.f / (1.f - p), seed_used, accumulate);
  LMKD_LAUNCH_CHECK("trx_dx_scatter_kernel");
  return 0;
}

int cast_bf16(const float* x, __nv_bfloat16* y, int64_t n, cudaStream_t stream) {
  cast_bf16_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(x, y, n);
  LMKD_LAUNCH_CHECK("cast_bf16_kernel");
  return 0;
}

int upcast_bf16(const __nv_bfloat16* x, float* y, int64_t n, cudaStream_t stream) {
  LMKD_CHECK(n % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             "upcast_bf16: n must be a multiple of 8 and both pointers 16-byte aligned");
  upcast_bf16_kernel<<<stream_grid(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x),
                                                                 reinterpret_cast<float4*>(y), n / 8);
  LMKD_LAUNCH_CHECK("upcast_bf16_kernel");
  return 0;
}

int dropout_mask(float* out, int64_t n, float p, uint64_t seed, cudaStream_t stream) {
  dropout_mask_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(out, n, p, 1.f / (1.f - p), seed);
  LMKD_LAUNCH_CHECK("dropout_mask_kernel");
  return 0;
}

}  // namespace lmkd
