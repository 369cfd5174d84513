// Kernels around the student feature heads (Linear(512 -> 2048) x heads on pooled frame features).
// The two contractions per direction run on the tcgen05 GEMM (bias epilogue forward; dX, dW backward);
// this file holds the patch pooling in front of them and the cast + bias-gradient pass behind them.
#include "feature_head.cuh"

namespace lmkd {

namespace {

// PyTorch adaptive pooling window i of `out` over `in` elements: [floor(i*in/out), ceil((i+1)*in/out))
__device__ __forceinline__ int win_lo(int i, int in, int out) { return (i * in) / out; }
__device__ __forceinline__ int win_hi(int i, int in, int out) { return ((i + 1) * in + out - 1) / out; }

// A block takes kPoolMaps consecutive (row, channel) maps -- one contiguous kPoolMaps*H*W slab of the NCHW trunk
// output -- through shared memory with 16-byte coalesced loads; thread t then reduces map t (stride H*W words
// between threads: conflict-free when H*W is odd, as for 7x7).
constexpr int kPoolMaps = 128;

__device__ __forceinline__ void pool_load_slab(const float* __restrict__ src, float* dst, int64_t count) {
  // src is 16-byte aligned (slabs start at multiples of kPoolMaps words)
  const int64_t n4 = count >> 2;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
  for (int64_t i = (n4 << 2) + threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
}

// TH, TW, TO > 0 fix the geometry at compile time (7x7 -> 4x4 is the reference's): windows and loops unroll
template <int TH, int TW, int TO>
__global__ void __launch_bounds__(kPoolMaps)
frame_pool_fwd_kernel(const float* __restrict__ fmap, float* __restrict__ pooled, int64_t n, int rH, int rW,
                      int r_out) {
  extern __shared__ __align__(16) float slab[];
  const int H = TH > 0 ? TH : rH, W = TW > 0 ? TW : rW, out_hw = TO > 0 ? TO : r_out;
  const int HW = H * W;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kPoolMaps;
  const int nm = static_cast<int>(n - m0 < kPoolMaps ? n - m0 : kPoolMaps);
  pool_load_slab(fmap + m0 * HW, slab, static_cast<int64_t>(nm) * HW);
  __syncthreads();
  if (static_cast<int>(threadIdx.x) >= nm) return;
  const float* m = slab + threadIdx.x * HW;
  float acc = 0.f;
#pragma unroll
  for (int oh = 0; oh < out_hw; ++oh) {
    const int h0 = win_lo(oh, H, out_hw), h1 = win_hi(oh, H, out_hw);
#pragma unroll
    for (int ow = 0; ow < out_hw; ++ow) {
      const int w0 = win_lo(ow, W, out_hw), w1 = win_hi(ow, W, out_hw);
      float mx = -INFINITY;
#pragma unroll
      for (int h = h0; h < h1; ++h)
#pragma unroll
        for (int w = w0; w < w1; ++w) mx = fmaxf(mx, m[h * W + w]);
      acc += mx;
    }
  }
  pooled[m0 + threadIdx.x] = acc / (out_hw * out_hw);
}

template <int TH, int TW, int TO>
__global__ void __launch_bounds__(kPoolMaps)
frame_pool_bwd_kernel(const float* __restrict__ fmap, const float* __restrict__ gp, float* __restrict__ gmap,
                      int64_t n, int rH, int rW, int r_out) {
  extern __shared__ __align__(16) float slab[];      // maps, then their gradients
  const int H = TH > 0 ? TH : rH, W = TW > 0 ? TW : rW, out_hw = TO > 0 ? TO : r_out;
  const int HW = H * W;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kPoolMaps;
  const int nm = static_cast<int>(n - m0 < kPoolMaps ? n - m0 : kPoolMaps);
  float* gslab = slab + kPoolMaps * HW;
  pool_load_slab(fmap + m0 * HW, slab, static_cast<int64_t>(nm) * HW);
  for (int i = threadIdx.x; i < nm * HW; i += blockDim.x) gslab[i] = 0.f;
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < nm) {
    const float* m = slab + threadIdx.x * HW;
    float* g = gslab + threadIdx.x * HW;
    const float share = gp[m0 + threadIdx.x] / (out_hw * out_hw);
#pragma unroll
    for (int oh = 0; oh < out_hw; ++oh) {
      const int h0 = win_lo(oh, H, out_hw), h1 = win_hi(oh, H, out_hw);
#pragma unroll
      for (int ow = 0; ow < out_hw; ++ow) {
        const int w0 = win_lo(ow, W, out_hw), w1 = win_hi(ow, W, out_hw);
        float mx = -INFINITY;
        int at = h0 * W + w0;
#pragma unroll
        for (int h = h0; h < h1; ++h)
#pragma unroll
          for (int w = w0; w < w1; ++w) {
            const float v = m[h * W + w];
            if (v > mx) {            // first maximum in scan order, as torch's max-pool backward picks it
              mx = v;
              at = h * W + w;
            }
          }
        g[at] += share;              // windows overlap: a cell can win several of them
      }
    }
  }
  __syncthreads();
  float* dst = gmap + m0 * HW;
  const int64_t count = static_cast<int64_t>(nm) * HW, n4 = count >> 2;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(gslab)[i];
  for (int64_t i = (n4 << 2) + threadIdx.x; i < count; i += blockDim.x) dst[i] = gslab[i];
}

// block = 128 threads x 4 columns = 512 columns, kRows rows; column sums leave through one atomic per thread
constexpr int kCastRows = 64;
__global__ void __launch_bounds__(128)
cast_colsum_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ yb, float* __restrict__ colsum,
                   int64_t rows, int cols) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c >= cols) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kCastRows;
  const int64_t r1 = r0 + kCastRows < rows ? r0 + kCastRows : rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = r0; r < r1; ++r) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(y + r * cols + c));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(yb + r * cols + c) =
        make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
  atomicAdd(colsum + c, acc.x);
  atomicAdd(colsum + c + 1, acc.y);
  atomicAdd(colsum + c + 2, acc.z);
  atomicAdd(colsum + c + 3, acc.w);
}

int pool_attrs() {
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(frame_pool_fwd_kernel<7, 7, 4>), 100 * 1024)) return rc;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(frame_pool_fwd_kernel<0, 0, 0>), 100 * 1024)) return rc;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(frame_pool_bwd_kernel<7, 7, 4>), 100 * 1024)) return rc;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(frame_pool_bwd_kernel<0, 0, 0>), 100 * 1024)) return rc;
  return 0;
}

}  // namespace

int frame_pool_fwd(const float* fmap, float* pooled, int64_t rows, int C, int H, int W, int out_hw, cudaStream_t st) {
  LMKD_CHECK(H > 0 && W > 0 && out_hw > 0 && out_hw <= H && out_hw <= W, "frame_pool: bad geometry %dx%d -> %d", H, W,
             out_hw);
  const int64_t n = rows * C;
  LMKD_CHECK(H * W <= 96, "frame_pool: %dx%d maps do not fit the shared-memory slab", H, W);
  LMKD_CHECK((reinterpret_cast<uintptr_t>(fmap) & 15) == 0, "frame_pool: input not 16-byte aligned");
  const size_t smem = sizeof(float) * kPoolMaps * H * W;
  if (int rc = pool_attrs()) return rc;
  const unsigned grid = static_cast<unsigned>(ceil_div(n, kPoolMaps));
  if (H == 7 && W == 7 && out_hw == 4)
    frame_pool_fwd_kernel<7, 7, 4><<<grid, kPoolMaps, smem, st>>>(fmap, pooled, n, H, W, out_hw);
  else
    frame_pool_fwd_kernel<0, 0, 0><<<grid, kPoolMaps, smem, st>>>(fmap, pooled, n, H, W, out_hw);
  LMKD_LAUNCH_CHECK("frame_pool_fwd_kernel");
  return 0;
}

int frame_pool_bwd(const float* fmap, const float* grad_pooled, float* grad_fmap, int64_t rows, int C, int H, int W,
                   int out_hw, cudaStream_t st) {
  LMKD_CHECK(H > 0 && W > 0 && out_hw > 0 && out_hw <= H && out_hw <= W, "frame_pool: bad geometry %dx%d -> %d", H, W,
             out_hw);
  const int64_t n = rows * C;
  LMKD_CHECK(H * W <= 96, "frame_pool: %dx%d maps do not fit the shared-memory slab", H, W);
  LMKD_CHECK(((reinterpret_cast<uintptr_t>(fmap) | reinterpret_cast<uintptr_t>(grad_fmap)) & 15) == 0,
             "frame_pool: maps not 16-byte aligned");
  const size_t smem = 2 * sizeof(float) * kPoolMaps * H * W;
  if (int rc = pool_attrs()) return rc;
  const unsigned grid = static_cast<unsigned>(ceil_div(n, kPoolMaps));
  if (H == 7 && W == 7 && out_hw == 4)
    frame_pool_bwd_kernel<7, 7, 4><<<grid, kPoolMaps, smem, st>>>(fmap, grad_pooled, grad_fmap, n, H, W, out_hw);
  else
    frame_pool_bwd_kernel<0, 0, 0><<<grid, kPoolMaps, smem, st>>>(fmap, grad_pooled, grad_fmap, n, H, W, out_hw);
  LMKD_LAUNCH_CHECK("frame_pool_bwd_kernel");
  return 0;
}

int cast_colsum(const float* y, __nv_bfloat16* yb, float* colsum, int64_t rows, int cols, cudaStream_t st) {
  LMKD_CHECK(cols % 4 == 0, "cast_colsum: %d columns not a multiple of 4", cols);
  dim3 grid(static_cast<unsigned>(ceil_div(cols, 512)), static_cast<unsigned>(ceil_div(rows, kCastRows)));
  cast_colsum_kernel<<<grid, 128, 0, st>>>(y, yb, colsum, rows, cols);
  LMKD_LAUNCH_CHECK("cast_colsum_kernel");
  return 0;
}

}  // namespace lmkd
