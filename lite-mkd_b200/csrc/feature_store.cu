// Teacher-feature store kernels: both are pure HBM streams over whole video rows (L*D contiguous elements),
// 16-byte accesses, no reuse -> L1 bypass on the loads, streaming stores.
//   tile = (row, chunk of kTileVec 16-byte vectors); blocks stride over the tile list.
#include "feature_store.cuh"

namespace lmkd {

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kTileVec = kThreads * kUnroll;     // float4 outputs per tile

__device__ __forceinline__ float4 ld_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_f4(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 bf4_to_f4(uint2 w) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// four consecutive elements (one output float4) number v of store row `src`
template <bool BF16>
__device__ __forceinline__ float4 load_store_vec(const void* store, int64_t src, int64_t row_vecs, int64_t v) {
  if (BF16) return bf4_to_f4(ld_u2(reinterpret_cast<const uint2*>(store) + src * row_vecs + v));
  return ld_f4(reinterpret_cast<const float4*>(store) + src * row_vecs + v);
}

template <bool BF16, bool MSE>
__global__ void __launch_bounds__(kThreads)
store_stream_kernel(const float* __restrict__ s, const void* __restrict__ store, int64_t store_rows,
                    const int64_t* __restrict__ index, int64_t count, int64_t row_vecs, int tiles_per_row,
                    float* __restrict__ out, float gscale, float* __restrict__ partials, int* __restrict__ status) {
  __shared__ float scratch[32];
  const int64_t ntiles = count * tiles_per_row;
  float acc = 0.f;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row = tile / tiles_per_row;
    const int chunk = static_cast<int>(tile - row * tiles_per_row);
    int64_t src = __ldg(index + row);
    const bool ok = src >= 0 && src < store_rows;
    if (!ok) {
      if (threadIdx.x == 0 && status != nullptr) atomicOr(status, 4);
      src = 0;
    }
    const int64_t v0 = static_cast<int64_t>(chunk) * kTileVec + threadIdx.x;
    float4 t[kUnroll], a[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = v0 + u * kThreads;
      t[u] = (ok && v < row_vecs) ? load_store_vec<BF16>(store, src, row_vecs, v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (MSE) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t v = v0 + u * kThreads;
        a[u] = v < row_vecs ? ld_f4(reinterpret_cast<const float4*>(s) + row * row_vecs + v) : t[u];
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t v = v0 + u * kThreads;
      if (v < row_vecs) {
        float4 o = t[u];
        if (MSE) {
          o = make_float4(a[u].x - t[u].x, a[u].y - t[u].y, a[u].z - t[u].z, a[u].w - t[u].w);
          acc += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
          o.x *= gscale; o.y *= gscale; o.z *= gscale; o.w *= gscale;
        }
        st_f4(reinterpret_cast<float4*>(out) + row * row_vecs + v, o);
      }
    }
  }
  if (MSE) {
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
  }
}

int launch(bool mse, const float* s, const void* store, int bf16, int64_t store_rows, const int64_t* index,
           int64_t count, int64_t row_elems, float* out, float gscale, float* partials, int max_blocks, int* nblocks,
           int* status, cudaStream_t st) {
  LMKD_CHECK(count > 0 && store_rows > 0 && row_elems > 0, "feature store: empty input");
  LMKD_CHECK(row_elems % 8 == 0, "feature store: row of %lld elements is not a multiple of 8", (long long)row_elems);
  LMKD_CHECK(((reinterpret_cast<uintptr_t>(store) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(s)) & 15) == 0,
             "feature store: pointers must be 16-byte aligned");
  const int64_t row_vecs = row_elems / 4;
  const int tiles_per_row = static_cast<int>(ceil_div(row_vecs, kTileVec));
  int64_t blocks = count * tiles_per_row;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > max_blocks) blocks = max_blocks;
  if (nblocks) *nblocks = static_cast<int>(blocks);
  const unsigned g = static_cast<unsigned>(blocks);
  if (mse) {
    if (bf16) store_stream_kernel<true, true><<<g, kThreads, 0, st>>>(s, store, store_rows, index, count, row_vecs, tiles_per_row, out, gscale, partials, status);
    else store_stream_kernel<false, true><<<g, kThreads, 0, st>>>(s, store, store_rows, index, count, row_vecs, tiles_per_row, out, gscale, partials, status);
  } else {
    if (bf16) store_stream_kernel<true, false><<<g, kThreads, 0, st>>>(s, store, store_rows, index, count, row_vecs, tiles_per_row, out, 0.f, nullptr, status);
    else store_stream_kernel<false, false><<<g, kThreads, 0, st>>>(s, store, store_rows, index, count, row_vecs, tiles_per_row, out, 0.f, nullptr, status);
  }
  LMKD_LAUNCH_CHECK("store_stream_kernel");
  return 0;
}

}  // namespace

int episode_gather(const void* store, int store_bf16, int64_t store_rows, const int64_t* index, int64_t count,
                   int64_t row_elems, float* out, int* status, cudaStream_t st) {
  return launch(false, nullptr, store, store_bf16, store_rows, index, count, row_elems, out, 0.f, nullptr, 1 << 30,
                nullptr, status, st);
}

int feat_mse_store_fwdbwd(const float* s, const void* store, int store_bf16, int64_t store_rows, const int64_t* index,
                          int64_t count, int64_t row_elems, float* ds, float gscale, float* partials, int max_partials,
                          int* npartials, int* status, cudaStream_t st) {
  return launch(true, s, store, store_bf16, store_rows, index, count, row_elems, ds, gscale, partials, max_partials,
                npartials, status, st);
}

}  // namespace lmkd
