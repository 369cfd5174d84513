// Teacher multi-modal fusion forward: the encoders that turn the per-modality frame features
// into the 2048-d `mm_features` the D2M feature loss is trained against.
//
// Reference: TrainablePositionalEncoding (teacher/code/model.py:1135-1151), ThreeTransforTemproal / TwoTransforFusion
// (:1361-1392, :1300-1331: torch nn.TransformerEncoderLayer, post-norm, ReLU, batch_first) and
// ThreeTRXShiftLoopTime.extract_feature (:1648-1664), in eval() (extract_multi_feature.py:114).
//
// Layout: a "row" is one frame of one video, M = videos * L rows.  Per encoder the residual stream X is kept in fp32
// [M, d] next to a bf16 copy that is the A operand of the next contraction; d = modalities * 2048 (6144 / 4096).
// Every Linear runs on the tcgen05 GEMM with a bias (+ReLU) epilogue; what is left for the kernels in this file is
// HBM-bound row work: positional encoding + LayerNorm + concatenation, residual + LayerNorm, and the 8-token
// self-attention (L x L scores per head: far too small for tensor cores, one block per (video, head)).
#include "fusion.cuh"

#include <cmath>

#include "gemm.cuh"

namespace lmkd {

namespace {

constexpr int kRowThreads = 256;

// block-wide (sum, sum of squares) of per-thread partials
__device__ __forceinline__ float2 block_sum2(float a, float b, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __syncthreads();
  if (lane == 0) {
    scratch[2 * warp] = a;
    scratch[2 * warp + 1] = b;
  }
  __syncthreads();
  a = lane < nw ? scratch[2 * lane] : 0.f;
  b = lane < nw ? scratch[2 * lane + 1] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  return make_float2(a, b);
}

// X[row][m * dmod + :] = LayerNorm(x_m[video][(l + shift_m) % L][:] + emb_m[l][:]); also the bf16 copy.
// grid (M, nmod); NV float4 per thread (dmod = 4 * NV * blockDim)
struct PeArgs {
  const float* x[4];
  const float* emb[4];
  const float* g[4];
  const float* b[4];
  int shift[4];
};
template <int NV>
__global__ void __launch_bounds__(kRowThreads)
fusion_pe_ln_kernel(const PeArgs a, int L, int dmod, int d, float eps, float* __restrict__ X,
                    __nv_bfloat16* __restrict__ Xb) {
  __shared__ float scratch[64];
  const int64_t row = blockIdx.x;
  const int m = blockIdx.y;
  const int l = static_cast<int>(row % L);
  const int64_t vid = row / L;
  const int ls = (l + a.shift[m]) % L;
  const float4* src = reinterpret_cast<const float4*>(a.x[m] + (vid * L + ls) * dmod);
  const float4* emb = reinterpret_cast<const float4*>(a.emb[m] + static_cast<int64_t>(l) * dmod);
  float4 v[NV];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    const float4 xv = __ldg(src + i), ev = __ldg(emb + i);
    v[k] = make_float4(xv.x + ev.x, xv.y + ev.y, xv.z + ev.z, xv.w + ev.w);
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mean = block_sum2(s, 0.f, scratch).x / dmod;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
    q += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  }
  const float rstd = rsqrtf(block_sum2(q, 0.f, scratch).x / dmod + eps);
  const float4* g4 = reinterpret_cast<const float4*>(a.g[m]);
  const float4* b4 = reinterpret_cast<const float4*>(a.b[m]);
  float4* out = reinterpret_cast<float4*>(X + row * d + static_cast<int64_t>(m) * dmod);
  uint2* outb = reinterpret_cast<uint2*>(Xb + row * d + static_cast<int64_t>(m) * dmod);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    const float4 g = __ldg(g4 + i), b = __ldg(b4 + i);
    const float4 y = make_float4(fmaf(v[k].x * rstd, g.x, b.x), fmaf(v[k].y * rstd, g.y, b.y),
                                 fmaf(v[k].z * rstd, g.z, b.z), fmaf(v[k].w * rstd, g.w, b.w));
    out[i] = y;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
    outb[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
  }
}

// X = LayerNorm(X + Y) in place (post-norm residual), plus the bf16 copy.  One block per row; the row is held in
// registers (NV float4 per thread, d <= 4 * NV * blockDim, ragged tail masked).
template <int NV>
__global__ void __launch_bounds__(kRowThreads)
residual_ln_kernel(float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int d, float eps, __nv_bfloat16* __restrict__ Xb) {
  __shared__ float scratch[64];
  const int64_t row = blockIdx.x;
  const int d4 = d >> 2;
  float4* x4 = reinterpret_cast<float4*>(X + row * d);
  const float4* y4 = reinterpret_cast<const float4*>(Y + row * d);
  float4 v[NV];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    if (i < d4) {
      const float4 a = x4[i], b = __ldg(y4 + i);
      v[k] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    } else {
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = block_sum2(s, 0.f, scratch).x / d;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    if (i < d4) {
      v[k].x -= mean; v[k].y -= mean; v[k].z -= mean; v[k].w -= mean;
      q += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
  }
  const float rstd = rsqrtf(block_sum2(q, 0.f, scratch).x / d + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* outb = reinterpret_cast<uint2*>(Xb + row * d);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * kRowThreads;
    if (i < d4) {
      const float4 g = __ldg(g4 + i), b = __ldg(b4 + i);
      const float4 y = make_float4(fmaf(v[k].x * rstd, g.x, b.x), fmaf(v[k].y * rstd, g.y, b.y),
                                   fmaf(v[k].z * rstd, g.z, b.z), fmaf(v[k].w * rstd, g.w, b.w));
      x4[i] = y;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
      outb[i] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
  }
}

// Self-attention over the L frames of one video for one head.  qkv bf16 [M, 3d] (q | k | v, head h at columns
// h * dh inside each third).  Block = (video, head), 256 threads, 16-byte accesses throughout: Q and K of the head are
// staged in shared memory as bf16 (2 x L x dh x 2 bytes: 64 KB at L = 8, dh = 2048, three blocks per SM); warp w owns
// query rows w, w + 8, ... and keeps the L scores of a row in registers while it sweeps dh once (the Q piece is read
// once per sweep step, the L key pieces against it); softmax of a row is L numbers in every lane; the context rows
// are written by one thread per 8 output columns reading V straight from global memory.
// (First version: 4-byte loads into fp32 staging, one (i, j) pair per warp pass, one block per SM: 400 us per launch
// at 800 videos x 3 heads for 211 MB, a quarter of the whole fusion forward.)
__device__ __forceinline__ void bf8_to_f8(const uint4 w, float (&f)[8]) {
  const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(u[i] << 16);
    f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

template <int LMAX>
__global__ void __launch_bounds__(256)
encoder_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int L, int d, int dh, float scale,
                         __nv_bfloat16* __restrict__ ctx) {
  extern __shared__ uint4 att_smem[];          // q [L][dh/8], k [L][dh/8] as 8 x bf16, then float p [L][LMAX]
  const int dh8 = dh >> 3;
  uint4* sq = att_smem;
  uint4* sk = sq + static_cast<size_t>(L) * dh8;
  float* sp = reinterpret_cast<float*>(sk + static_cast<size_t>(L) * dh8);
  const int64_t vid = blockIdx.x;
  const int h = blockIdx.y;
  const int64_t row_stride8 = (3 * static_cast<int64_t>(d)) >> 3;              // uint4 per qkv row
  const uint4* base = reinterpret_cast<const uint4*>(qkv + vid * L * 3 * static_cast<int64_t>(d) + static_cast<int64_t>(h) * dh);
  const int d8 = d >> 3;
  for (int i = threadIdx.x; i < L * dh8; i += blockDim.x) {
    const int l = i / dh8, c = i % dh8;
    const uint4* r = base + l * row_stride8 + c;
    sq[i] = __ldg(r);
    sk[i] = __ldg(r + d8);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < L; i += nw) {
    float acc[LMAX];
#pragma unroll
    for (int j = 0; j < LMAX; ++j) acc[j] = 0.f;
    for (int c = lane; c < dh8; c += 32) {
      float qf[8];
      bf8_to_f8(sq[i * dh8 + c], qf);
#pragma unroll
      for (int j = 0; j < LMAX; ++j) {
        if (j < L) {
          float kf[8];
          bf8_to_f8(sk[j * dh8 + c], kf);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j] = fmaf(qf[e], kf[e], acc[j]);
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < LMAX; ++j) {
      if (j < L) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        acc[j] *= scale;
        mx = fmaxf(mx, acc[j]);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < LMAX; ++j) {
      if (j < L) {
        acc[j] = __expf(acc[j] - mx);
        sum += acc[j];
      }
    }
    const float inv = 1.f / sum;
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < LMAX; ++j)
        if (j < L) sp[i * LMAX + j] = acc[j] * inv;
    }
  }
  __syncthreads();
  const uint4* vb = base + 2 * d8;
  uint4* out = reinterpret_cast<uint4*>(ctx + vid * L * static_cast<int64_t>(d) + static_cast<int64_t>(h) * dh);
  for (int c = threadIdx.x; c < dh8; c += blockDim.x) {
    float o[LMAX][8];
#pragma unroll
    for (int i = 0; i < LMAX; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) o[i][e] = 0.f;
#pragma unroll
    for (int j = 0; j < LMAX; ++j) {
      if (j < L) {
        float vf[8];
        bf8_to_f8(__ldg(vb + j * row_stride8 + c), vf);
#pragma unroll
        for (int i = 0; i < LMAX; ++i) {
          if (i < L) {
            const float pij = sp[i * LMAX + j];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[i][e] = fmaf(pij, vf[e], o[i][e]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < LMAX; ++i) {
      if (i < L) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 hv = __floats2bfloat162_rn(o[i][2 * e], o[i][2 * e + 1]);
          w[e] = *reinterpret_cast<uint32_t*>(&hv);
        }
        out[static_cast<int64_t>(i) * d8 + c] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

// out[row][:] += bias[:]   (the accumulating output GEMM carries no bias term of its own)
__global__ void add_bias_rows_kernel(float* __restrict__ out, const float* __restrict__ bias, int64_t rows, int n4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * n4) return;
  float4 v = reinterpret_cast<float4*>(out)[i];
  const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (i % n4));
  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  reinterpret_cast<float4*>(out)[i] = v;
}

int linear(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, int64_t M, int N, int K, void* C,
           int kind, int relu, cudaStream_t st) {
  GemmDesc g;
  g.M = static_cast<int>(M); g.N = N; g.K = K;
  g.A.ptr = A; g.A.ld = K;
  g.B.ptr = W; g.B.ld = K;
  g.epi.kind = kind; g.epi.alpha = 1.f; g.epi.relu = relu;
  g.epi.C = C; g.epi.ldc = N;
  g.epi.colv = bias;
  return gemm_bf16(g, st);
}

template <int NV>
int launch_residual_ln(float* X, const float* Y, const float* g, const float* b, int64_t M, int d, float eps,
                       __nv_bfloat16* Xb, cudaStream_t st) {
  residual_ln_kernel<NV><<<static_cast<unsigned>(M), kRowThreads, 0, st>>>(X, Y, g, b, d, eps, Xb);
  LMKD_LAUNCH_CHECK("residual_ln_kernel");
  return 0;
}
int residual_ln(float* X, const float* Y, const float* g, const float* b, int64_t M, int d, float eps,
                __nv_bfloat16* Xb, cudaStream_t st) {
  const int need = static_cast<int>(ceil_div(d / 4, kRowThreads));
  if (need <= 2) return launch_residual_ln<2>(X, Y, g, b, M, d, eps, Xb, st);
  if (need <= 4) return launch_residual_ln<4>(X, Y, g, b, M, d, eps, Xb, st);
  if (need <= 6) return launch_residual_ln<6>(X, Y, g, b, M, d, eps, Xb, st);
  return launch_residual_ln<8>(X, Y, g, b, M, d, eps, Xb, st);
}

}  // namespace

int fusion_check(const FusionEncoder& e, int64_t nvideos, int L) {
  LMKD_CHECK(e.nmod >= 1 && e.nmod <= 4, "fusion: %d modalities unsupported (1..4)", e.nmod);
  LMKD_CHECK(e.dmod > 0 && e.dmod % (4 * kRowThreads) == 0 && e.dmod / (4 * kRowThreads) <= 4,
             "fusion: features per modality (%d) must be a multiple of %d, at most %d", e.dmod, 4 * kRowThreads,
             16 * kRowThreads);
  const int d = e.nmod * e.dmod;
  LMKD_CHECK(d / 4 <= 8 * kRowThreads, "fusion: model width %d too large", d);
  LMKD_CHECK(e.nhead >= 1 && d % e.nhead == 0 && (d / e.nhead) % 8 == 0, "fusion: %d heads do not divide width %d", e.nhead, d);
  LMKD_CHECK(e.dff > 0 && e.dff % 8 == 0 && e.dout > 0 && e.dout % 8 == 0, "fusion: dff (%d) / dout (%d) must be multiples of 8",
             e.dff, e.dout);
  LMKD_CHECK(e.nlayers >= 0 && (e.nlayers == 0 || e.layers != nullptr), "fusion: missing layer table");
  LMKD_CHECK(L >= 1 && L <= 16, "fusion: %d frames unsupported (1..16)", L);
  LMKD_CHECK(nvideos > 0 && nvideos * L < (1ll << 31), "fusion: bad video count");
  const size_t att_smem = 2 * 2 * static_cast<size_t>(L) * (d / e.nhead) + sizeof(float) * static_cast<size_t>(L) * 16;
  LMKD_CHECK(att_smem <= 200 * 1024, "fusion: head width %d x %d frames does not fit shared memory", d / e.nhead, L);
  return 0;
}

FusionWs fusion_layout(void* ws, const FusionEncoder& e, int64_t nvideos, int L) {
  FusionWs w;
  const int64_t M = nvideos * L;
  const int64_t d = static_cast<int64_t>(e.nmod) * e.dmod;
  uint8_t* base = static_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  };
  w.X = static_cast<float*>(take(sizeof(float) * M * d));
  w.Y = static_cast<float*>(take(sizeof(float) * M * d));
  w.Xb = static_cast<__nv_bfloat16*>(take(2 * M * d));
  w.qkv = static_cast<__nv_bfloat16*>(take(2 * M * 3 * d));
  w.ctx = static_cast<__nv_bfloat16*>(take(2 * M * d));
  w.H = static_cast<__nv_bfloat16*>(take(2 * M * e.dff));
  w.bytes = (off + 255) & ~static_cast<size_t>(255);
  return w;
}

int fusion_forward(const FusionEncoder& e, const float* const* x, const int* shift, int64_t nvideos, int L, float* out,
                   int accumulate, void* workspace, cudaStream_t st) {
  if (int rc = fusion_check(e, nvideos, L)) return rc;
  LMKD_CHECK(x != nullptr && out != nullptr && workspace != nullptr, "fusion_fwd: null pointer");
  const FusionWs w = fusion_layout(workspace, e, nvideos, L);
  const int64_t M = nvideos * L;
  const int d = e.nmod * e.dmod, dh = d / e.nhead;
  // ---- positional encodings + LayerNorm, concatenated (teacher/code/model.py:1143-1151, 1385-1389) ----
  PeArgs pa{};
  for (int m = 0; m < e.nmod; ++m) {
    LMKD_CHECK(x[m] && e.pe_emb[m] && e.pe_g[m] && e.pe_b[m], "fusion_fwd: null input / positional table %d", m);
    pa.x[m] = x[m]; pa.emb[m] = e.pe_emb[m]; pa.g[m] = e.pe_g[m]; pa.b[m] = e.pe_b[m];
    pa.shift[m] = shift ? ((shift[m] % L) + L) % L : 0;
  }
  {
    const dim3 grid(static_cast<unsigned>(M), e.nmod);
    switch (e.dmod / (4 * kRowThreads)) {
      case 1: fusion_pe_ln_kernel<1><<<grid, kRowThreads, 0, st>>>(pa, L, e.dmod, d, e.ln_eps, w.X, w.Xb); break;
      case 2: fusion_pe_ln_kernel<2><<<grid, kRowThreads, 0, st>>>(pa, L, e.dmod, d, e.ln_eps, w.X, w.Xb); break;
      case 3: fusion_pe_ln_kernel<3><<<grid, kRowThreads, 0, st>>>(pa, L, e.dmod, d, e.ln_eps, w.X, w.Xb); break;
      default: fusion_pe_ln_kernel<4><<<grid, kRowThreads, 0, st>>>(pa, L, e.dmod, d, e.ln_eps, w.X, w.Xb); break;
    }
    LMKD_LAUNCH_CHECK("fusion_pe_ln_kernel");
  }
  // ---- encoder stack (post-norm) ----
  const size_t att_smem = 2 * 2 * static_cast<size_t>(L) * dh + sizeof(float) * static_cast<size_t>(L) * 16;   // bf16 Q, K + p
  auto att_kern = L <= 8 ? encoder_attention_kernel<8> : encoder_attention_kernel<16>;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(att_kern), 200 * 1024)) return rc;
  for (int i = 0; i < e.nlayers; ++i) {
    const FusionLayer& ly = e.layers[i];
    LMKD_CHECK(ly.w_qkv && ly.b_qkv && ly.w_o && ly.b_o && ly.w_ff1 && ly.b_ff1 && ly.w_ff2 && ly.b_ff2 && ly.ln1_g &&
                   ly.ln1_b && ly.ln2_g && ly.ln2_b,
               "fusion_fwd: null parameter in layer %d", i);
    if (int rc = linear(w.Xb, static_cast<const __nv_bfloat16*>(ly.w_qkv), ly.b_qkv, M, 3 * d, d, w.qkv, EPI_BIAS_BF16, 0, st))
      return rc;
    att_kern<<<dim3(static_cast<unsigned>(nvideos), e.nhead), 256, att_smem, st>>>(w.qkv, L, d, dh, 1.f / sqrtf(static_cast<float>(dh)),
                                                                                    w.ctx);
    LMKD_LAUNCH_CHECK("encoder_attention_kernel");
    if (int rc = linear(w.ctx, static_cast<const __nv_bfloat16*>(ly.w_o), ly.b_o, M, d, d, w.Y, EPI_BIAS_F32, 0, st)) return rc;
    if (int rc = residual_ln(w.X, w.Y, ly.ln1_g, ly.ln1_b, M, d, e.ln_eps, w.Xb, st)) return rc;
    if (int rc = linear(w.Xb, static_cast<const __nv_bfloat16*>(ly.w_ff1), ly.b_ff1, M, e.dff, d, w.H, EPI_BIAS_BF16, 1, st))
      return rc;
    if (int rc = linear(w.H, static_cast<const __nv_bfloat16*>(ly.w_ff2), ly.b_ff2, M, d, e.dff, w.Y, EPI_BIAS_F32, 0, st)) return rc;
    if (int rc = residual_ln(w.X, w.Y, ly.ln2_g, ly.ln2_b, M, d, e.ln_eps, w.Xb, st)) return rc;
  }
  // ---- f1: Linear down to `dout`, written or added to the running sum of the streams (:1391, :1664) ----
  LMKD_CHECK(e.w_out && e.b_out, "fusion_fwd: null output projection");
  if (!accumulate) return linear(w.Xb, static_cast<const __nv_bfloat16*>(e.w_out), e.b_out, M, e.dout, d, out, EPI_BIAS_F32, 0, st);
  if (int rc = linear(w.Xb, static_cast<const __nv_bfloat16*>(e.w_out), nullptr, M, e.dout, d, out, EPI_ACCUM_F32, 0, st)) return rc;
  const int64_t n = M * (e.dout / 4);
  add_bias_rows_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(out, e.b_out, M, e.dout / 4);
  LMKD_LAUNCH_CHECK("add_bias_rows_kernel");
  return 0;
}

}  // namespace lmkd
