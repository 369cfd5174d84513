// OTAM ordered temporal alignment (reference: teacher/code/model.py:3271-3343).
//
// otam_dp_fwd : anti-diagonal wavefront over the zero-padded L x (M+2) cumulative-distance table.
//               One lane per table row, a group of G lanes per (query, support, direction) task,
//               32/G tasks per warp.  Lane r needs C[r-1][m-1], C[r-1][m] (two shuffles from lane
//               r-1) and its own C[r][m-1]; soft-min in the stabilised log-sum-exp form, which is
//               identical to the reference wherever the reference is finite.
// otam_dp_bwd : recomputes the table into shared memory, walks the anti-diagonals in reverse,
//               emits d(loss)/d(numerator) in bf16 for the two feature-gradient GEMMs and
//               accumulates the row/column norm gradients of the cosine similarity.
// otam_class  : per-class mean over supports + softmax(-d) over classes (and its backward).
#include "otam.cuh"

#include <algorithm>

namespace lmkd {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// shared-memory accesses of the sweeps by 32-bit shared address: one LDS/STS each, no generic-pointer arithmetic
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v));
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct DpShape {
  int B, Nq, Ns, L, M;   // L query frames, M support frames
  int G;                 // lanes per task (power of two >= max(L, M), >= 8)
  int pairs_per_warp;    // 1 (G >= 16) or 2 (G == 8)
  int dirs_concurrent;   // 2 if both directions fit in the warp at once, else 1
  int npass;             // sequential passes over directions (1 or 2)
  int single_dir;        // 1: only direction 0 (raw OTAM_cum_dist test hook)
  int64_t ld;            // pitch of dist / dnum
  float lbda;
};

// Loads the L x M block of pair (q, s) into smem (row-major, pitch M), multiplied by `scale`.
__device__ __forceinline__ void load_block(const float* __restrict__ dist, float* blk, int L, int M, int64_t ld,
                                           int lane_in, int nlanes, float scale) {
  if (((M | static_cast<int>(ld)) & 3) == 0 && (reinterpret_cast<uintptr_t>(dist) & 15) == 0) {
    const int vpr = M >> 2;   // float4 per row
    for (int i = lane_in; i < L * vpr; i += nlanes) {
      const int l = i / vpr, j = i - l * vpr;
      float4 v = __ldg(reinterpret_cast<const float4*>(dist + static_cast<int64_t>(l) * ld) + j);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      *reinterpret_cast<float4*>(blk + l * M + 4 * j) = v;
    }
  } else {
    for (int i = lane_in; i < L * M; i += nlanes) {
      const int l = i / M, m = i - l * M;
      blk[i] = __ldg(dist + static_cast<int64_t>(l) * ld + m) * scale;
    }
  }
}

// One DP sweep of one task by a group of G lanes, in the scaled domain C' = C * log2(e) / lambda, where the
// soft-min is  min - log2(sum 2^(min - x)).  `R` rows (this lane is row r), `Cn` columns;
// d'(r, c) = blk[r * sr + c * sc].  Every cell is evaluated as a three-way soft-min over
// (diagonal, up, left); the inputs a cell does not have (top row: diagonal and up; interior columns: up) are
// +inf, which contributes exactly 0 to the sum, so all lanes run the same instruction stream.
// STORE: the soft-min weights of every cell towards its diagonal and left inputs (= d C[r][m] / d input) go
// to Wd / Wl ([R][W]); the weight of the "up" input is 1 - wd - wl on the two three-way columns, 0 elsewhere.
// Returns C'[R-1][Cn+1] in lane R-1 (garbage elsewhere).
template <bool STORE>
__device__ __forceinline__ float dp_sweep(const float* blk, int sr, int sc, int R, int Cn, int r, int G, int steps,
                                          float* Wd, float* Wl, int W) {
  // `steps` (= L + M for either direction) is warp-uniform; lanes without a task pass R = 0.
  const float inf = __int_as_float(0x7f800000);
  const bool row_ok = r < R, top = r == 0;
  // running shared addresses of d'(r, m - 1), Wd[r][m], Wl[r][m] for m = t - r
  uint32_t a_blk = static_cast<uint32_t>(__cvta_generic_to_shared(blk)) + 4u * (r * sr - r * sc);
  uint32_t a_wd = STORE ? static_cast<uint32_t>(__cvta_generic_to_shared(Wd)) + 4u * (r * W + 1 - r) : 0u;
  uint32_t a_wl = STORE ? static_cast<uint32_t>(__cvta_generic_to_shared(Wl)) + 4u * (r * W + 1 - r) : 0u;
  const uint32_t blk_step = 4u * sc;
  float cur = 0.f, prevcur = 0.f;
  for (int t = 1; t <= steps; ++t) {
    const float up1 = __shfl_up_sync(0xffffffffu, cur, 1, G);      // C[r-1][m]
    const float up2 = __shfl_up_sync(0xffffffffu, prevcur, 1, G);  // C[r-1][m-1]
    const int m = t - r;
    const bool act = row_ok && static_cast<unsigned>(m - 1) <= static_cast<unsigned>(Cn);   // 1 <= m <= Cn + 1
    const bool edge = (m == 1) || (m == Cn + 1);
    float dv = 0.f;
    if (act && m <= Cn) dv = lds_f32(a_blk);
    const float xd = top ? inf : up2;
    const float xu = (edge && !top) ? up1 : inf;
    const float mn = fminf(fminf(xd, xu), cur);
    const float ed = ex2_approx(mn - xd), eu = ex2_approx(mn - xu), el = ex2_approx(mn - cur);
    const float ssum = (ed + eu) + el;
    const float nv = (dv + mn) - lg2_approx(ssum);
    if (STORE) {
      const float inv = rcp_approx(ssum);
      if (act) {
        sts_f32(a_wd, ed * inv);
        sts_f32(a_wl, el * inv);
      }
      a_wd += 4u;
      a_wl += 4u;
    }
    if (act) {
      prevcur = cur;
      cur = nv;
    }
    a_blk += blk_step;
  }
  return cur;
}

// Reverse sweep: lane r walks its row right to left.  The gradient of cell (r, m) is the sum of the messages
// g * weight sent by its three consumers: (r, m+1) (own lane, previous step), (r+1, m) and (r+1, m+1)
// (lane r+1: its latest "up" message and its previous "diagonal" message, two shuffles).  The gradient
// overwrites the cell's Wd slot: for m in 1..Cn it is d result / d d[r][m-1].  Lanes that are not yet active
// carry all-zero state (g is forced to 0), lanes past their row are never read again, so only the shared
// memory accesses are predicated.
__device__ __forceinline__ void dp_reverse(float* Wd, const float* Wl, int R, int Cn, int r, int G, int steps,
                                           float gout, int W) {
  const bool row_ok = r < R, has_below = r + 1 < R;
  // running shared addresses of Wd[r][m], Wl[r][m] for m = t - r, t = steps .. 1
  uint32_t a_wd = static_cast<uint32_t>(__cvta_generic_to_shared(Wd)) + 4u * (r * W + steps - r);
  uint32_t a_wl = static_cast<uint32_t>(__cvta_generic_to_shared(Wl)) + 4u * (r * W + steps - r);
  float left = (row_ok && r == R - 1) ? gout : 0.f;   // seeds cell (R-1, Cn+1), the first one visited
  float up_msg = 0.f, diag_cur = 0.f, diag_prev = 0.f;
  for (int t = steps; t >= 1; --t) {
    const float ru = __shfl_down_sync(0xffffffffu, up_msg, 1, G);
    const float rd = __shfl_down_sync(0xffffffffu, diag_prev, 1, G);
    const int m = t - r;
    const bool act = row_ok && static_cast<unsigned>(m - 1) <= static_cast<unsigned>(Cn);
    const bool edge = (m == 1) || (m == Cn + 1);
    float wd = 0.f, wl = 0.f;
    if (act) {
      wd = lds_f32(a_wd);
      wl = lds_f32(a_wl);
    }
    const float below = has_below ? ru + rd : 0.f;
    const float g = act ? left + below : 0.f;
    const float wu = edge ? fmaxf((1.f - wd) - wl, 0.f) : 0.f;
    if (act) sts_f32(a_wd, g);
    left = g * wl;
    diag_prev = diag_cur;
    diag_cur = g * wd;
    up_msg = g * wu;
    a_wd -= 4u;
    a_wl -= 4u;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
otam_dp_fwd_kernel(const float* __restrict__ dist, float* __restrict__ pair, const DpShape p) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk_elems = p.L * p.M;
  float* wblk = smem + warp * p.pairs_per_warp * blk_elems;
  const uint32_t npairs = static_cast<uint32_t>(p.B) * p.Nq * p.Ns;   // < 2^31, checked by the host
  const int lanes_per_pair = 32 / p.pairs_per_warp;
  const int pw = lane / lanes_per_pair;          // which pair of this warp
  const int lp = lane - pw * lanes_per_pair;     // lane within the pair's lanes
  const float k = kLog2e / p.lbda;
  float* blk = wblk + pw * blk_elems;
  const int dir_of_lane = lp / p.G;   // 0 or 1 when both directions run concurrently
  const int r = lp - dir_of_lane * p.G;
  // persistent warps: a warp strides over the pair list (block launch rate would otherwise bound the kernel)
  const uint32_t stride = gridDim.x * kWarpsPerBlock * p.pairs_per_warp;
  for (uint32_t first = (blockIdx.x * kWarpsPerBlock + warp) * p.pairs_per_warp; first < npairs; first += stride) {
  const uint32_t pid = first + pw;
  const bool valid = pid < npairs;
  int64_t b = 0;
  int q = 0, s = 0;
  if (valid) {
    const uint32_t qs = pid / p.Ns;
    s = static_cast<int>(pid - qs * p.Ns);
    b = qs / p.Nq;
    q = static_cast<int>(qs - static_cast<uint32_t>(b) * p.Nq);
  }
  if (valid) {
    const float* src = dist + (b * p.Nq * p.L + static_cast<int64_t>(q) * p.L) * p.ld + static_cast<int64_t>(s) * p.M;
    load_block(src, blk, p.L, p.M, p.ld, lp, lanes_per_pair, k);
  }
  __syncwarp();
  float total = 0.f;
  for (int pass = 0; pass < p.npass; ++pass) {
    const int dir = p.dirs_concurrent == 2 ? dir_of_lane : pass;
    // dir 0: rows = query frames (L), cols = support frames (M); dir 1: transposed
    const int R = dir == 0 ? p.L : p.M, Cn = dir == 0 ? p.M : p.L;
    const int sr = dir == 0 ? p.M : 1, sc = dir == 0 ? 1 : p.M;
    const bool lane_has_task = lp < p.G * p.dirs_concurrent;
    const float res = dp_sweep<false>(blk, sr, sc, lane_has_task ? R : 0, Cn, r, p.G, p.L + p.M, nullptr, nullptr, 0);
    // fetch the result from the last row's lane of each direction group
    const int src_lane = pw * lanes_per_pair + dir_of_lane * p.G + (R - 1);
    const float got = __shfl_sync(0xffffffffu, res, p.dirs_concurrent == 2 ? src_lane
                                                                           : pw * lanes_per_pair + (R - 1));
    if (lane_has_task) total += got;
  }
  if (!(lp < p.G * p.dirs_concurrent)) total = 0.f;
  // lanes of direction 0 and 1 hold their own direction's value; combine
  if (p.dirs_concurrent == 2) {
    const float other = __shfl_xor_sync(0xffffffffu, total, p.G);
    total += other;
  }
  if (valid && lp == 0) pair[pid] = total * (p.lbda / kLog2e);
  __syncwarp();   // the block buffer is reloaded by the next iteration
  }
}

// shared memory of one pair in the backward, in floats:
//   blk [L*M] (scaled distances) | per direction: Wd [tab] (weights, then gradients), Wl [tab]
// Lanes walk an anti-diagonal, i.e. lane r touches offset r * (pitch - 1) + t of its direction's table.  The
// pitch is padded so that the lanes of one group fall on distinct banks 32/G apart, and the (up to four)
// groups of a warp are shifted by 0, 1, 2, 3 banks: no bank conflicts in either sweep.
__host__ __device__ inline int tab_pitch(int Cn, int G) {
  const int mod = G == 8 ? 8 : (G == 16 ? 4 : 2);
  int w = Cn + 2;
  while ((w - 1) % mod != mod / 2) ++w;
  return w;
}
__host__ __device__ inline int bwd_tab(int L, int M, int G) {
  const int a = L * tab_pitch(M, G), b = M * tab_pitch(L, G);
  return a > b ? a : b;
}
__host__ __device__ inline int bwd_dir_floats(int L, int M, int G) {
  return ((2 * bwd_tab(L, M, G) + 3) & ~3) + 1;
}
__host__ __device__ inline int bwd_pair_floats(int L, int M, int G) {
  return ((L * M + 3) & ~3) + 2 * bwd_dir_floats(L, M, G) + 6;   // multiple of 4: blk stays 16-byte aligned
}

// backward: one pair per warp-slot as in the forward.  Forward sweep again (storing the soft-min weights),
// reverse sweep, then d dist -> d numerator (bf16, for the two feature-gradient GEMMs) and norm gradients.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
otam_dp_bwd_kernel(const float* __restrict__ dist, const float* __restrict__ gpair,
                   const float* __restrict__ nq, const float* __restrict__ ns,
                   __nv_bfloat16* __restrict__ dnum, float* __restrict__ gnq, float* __restrict__ gns,
                   float* __restrict__ ddist_raw, const DpShape p, float eps) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk_elems = p.L * p.M;
  const int tab = bwd_tab(p.L, p.M, p.G);
  const int dir_floats = bwd_dir_floats(p.L, p.M, p.G);
  const int W0 = tab_pitch(p.M, p.G), W1 = tab_pitch(p.L, p.G);   // table pitch of direction 0 / 1
  const uint32_t npairs = static_cast<uint32_t>(p.B) * p.Nq * p.Ns;
  const int lanes_per_pair = 32 / p.pairs_per_warp;
  const int pw = lane / lanes_per_pair;
  const int lp = lane - pw * lanes_per_pair;
  float* base = smem + (warp * p.pairs_per_warp + pw) * bwd_pair_floats(p.L, p.M, p.G);
  float* blk = base;
  float* dir_base = base + ((blk_elems + 3) & ~3) + 2 * (pw & 1);
  const float k = kLog2e / p.lbda, inv_k = p.lbda / kLog2e;
  const int dir_of_lane = lp / p.G;
  const int r = lp - dir_of_lane * p.G;
  const bool lane_has_task = lp < p.G * p.dirs_concurrent;
  const bool both = !p.single_dir;
  const float* G0 = dir_base;                  // [L][W0], entry (l, m + 1)
  const float* G1 = dir_base + dir_floats;     // [M][W1], entry (m, l + 1)
  float* tx = blk;                   // contribution of (l, m) to d|x_l|   (blk is consumed element by element)
  float* ty = dir_base + tab;        // contribution to d|y_m|             (direction 0's Wl, no longer needed)
  // element walk of the epilogue: lane lp owns elements lp, lp + nl, ... of the L x M block
  const int nl = lanes_per_pair;
  const int l0 = lp / p.M, m0 = lp - l0 * p.M, dl = nl / p.M, dm = nl - dl * p.M;
  const int ld32 = static_cast<int>(p.ld);
  const uint32_t stride = gridDim.x * kWarpsPerBlock * p.pairs_per_warp;
  for (uint32_t first = (blockIdx.x * kWarpsPerBlock + warp) * p.pairs_per_warp; first < npairs; first += stride) {
    const uint32_t pid = first + pw;
    const bool valid = pid < npairs;
    int64_t b = 0;
    int q = 0, s = 0;
    if (valid) {
      const uint32_t qs = pid / p.Ns;
      s = static_cast<int>(pid - qs * p.Ns);
      b = qs / p.Nq;
      q = static_cast<int>(qs - static_cast<uint32_t>(b) * p.Nq);
    }
    const int64_t row0 = (b * p.Nq + q) * p.L;                                  // first query-frame row
    const int64_t scol0 = (b * p.Ns + s) * static_cast<int64_t>(p.M);           // first support-frame row
    const int64_t blk0 = row0 * p.ld + static_cast<int64_t>(s) * p.M;           // block origin in dist / dnum
    if (valid) load_block(dist + blk0, blk, p.L, p.M, p.ld, lp, lanes_per_pair, k);
    __syncwarp();
    const float gout = valid ? __ldg(gpair + pid) : 0.f;
    for (int pass = 0; pass < p.npass; ++pass) {
      const int dir = p.dirs_concurrent == 2 ? dir_of_lane : pass;
      const int R = dir == 0 ? p.L : p.M, Cn = dir == 0 ? p.M : p.L;
      const int sr = dir == 0 ? p.M : 1, sc = dir == 0 ? 1 : p.M;
      float* Wd = dir_base + (lane_has_task ? dir : 0) * dir_floats;
      float* Wl = Wd + tab;
      const int Rl = lane_has_task ? R : 0;
      const int W = dir == 0 ? W0 : W1;
      dp_sweep<true>(blk, sr, sc, Rl, Cn, r, p.G, p.L + p.M, Wd, Wl, W);
      dp_reverse(Wd, Wl, Rl, Cn, r, p.G, p.L + p.M, gout, W);
    }
    __syncwarp();
    if (ddist_raw != nullptr) {   // test hook: raw d loss / d dist, no cosine chain
      if (valid) {
        int l = l0, m = m0;
        for (int i = lp; i < blk_elems; i += nl) {
          float g = G0[l * W0 + m + 1];
          if (both) g += G1[m * W1 + l + 1];
          ddist_raw[blk0 + l * ld32 + m] = g;
          m += dm; l += dl;
          if (m >= p.M) { m -= p.M; ++l; }
        }
      }
      __syncwarp();
      continue;
    }
    // d dist -> d numerator (bf16) and norm gradients.  dist = 1 - num / (|x||y| + eps)
    if (valid) {
      const float* nqp = nq + row0;
      const float* nsp = ns + scol0;
      __nv_bfloat16* dnp = dnum + blk0;
      int l = l0, m = m0;
      for (int i = lp; i < blk_elems; i += nl) {
        float g = G0[l * W0 + m + 1];
        if (both) g += G1[m * W1 + l + 1];
        const float nx = __ldg(nqp + l), ny = __ldg(nsp + m);
        const float gr = g * rcp_approx(fmaf(nx, ny, eps));   // g / den
        const float sim = 1.f - blk[i] * inv_k;
        dnp[l * ld32 + m] = __float2bfloat16_rn(-gr);
        const float tt = gr * sim;        // dL/d(den)
        tx[i] = tt * ny;
        ty[i] = tt * nx;
        m += dm; l += dl;
        if (m >= p.M) { m -= p.M; ++l; }
      }
    }
    __syncwarp();
    if (valid) {
      // lanes 0..L-1 sum rows of tx (d|x_l|), lanes L..L+M-1 sum columns of ty (d|y_m|): one strided loop
      for (int j = lp; j < p.L + p.M; j += nl) {
        const bool is_row = j < p.L;
        const float* src = is_row ? tx + j * p.M : ty + (j - p.L);
        const int step = is_row ? 1 : p.M, cnt = is_row ? p.M : p.L;
        float acc = 0.f;
        for (int c = 0; c < cnt; ++c) acc += src[c * step];
        atomicAdd(is_row ? gnq + row0 + j : gns + scol0 + (j - p.L), acc);
      }
    }
    __syncwarp();   // the pair's shared memory is reused by the next iteration
  }
}

// ---- class mean + softmax over classes ----------------------------------------------------
// one warp per (b, q); class id = (int)label; classes with no support are excluded (prob 0)
__global__ void otam_class_fwd_kernel(const float* __restrict__ pair, const float* __restrict__ labels,
                                      const int* __restrict__ nanflag, float* __restrict__ probs, int B,
                                      int Nq, int Ns, int way, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= static_cast<int64_t>(B) * Nq) return;
  const int64_t b = wid / Nq;
  const float* pr = pair + wid * Ns;
  const float* lb = labels + b * Ns;
  float* out = probs + wid * way;
  if (nanflag != nullptr && nanflag[b] != 0) {   // reference NaN guard: all-zero logits
    for (int c = lane; c < way; c += 32) out[c] = 0.f;
    return;
  }
  // lane c (strided) owns class c
  float mx = -INFINITY;
  for (int c = lane; c < way; c += 32) {
    float sum = 0.f;
    int cnt = 0;
    for (int s = 0; s < Ns; ++s) {
      const int cls = static_cast<int>(lb[s]);
      if (cls < 0 || cls >= way) { if (status) atomicOr(status, 1); continue; }
      if (cls == c) { sum += pr[s]; ++cnt; }
    }
    const float z = cnt > 0 ? -(sum / cnt) : -INFINITY;
    out[c] = z;
    mx = fmaxf(mx, z);
  }
  mx = warp_max(mx);
  float den = 0.f;
  for (int c = lane; c < way; c += 32) {
    const float e = (out[c] == -INFINITY) ? 0.f : __expf(out[c] - mx);
    out[c] = e;
    den += e;
  }
  den = warp_sum(den);
  for (int c = lane; c < way; c += 32) out[c] = out[c] / den;
}

__global__ void otam_class_bwd_kernel(const float* __restrict__ gprobs, const float* __restrict__ probs,
                                      const float* __restrict__ labels, const int* __restrict__ nanflag,
                                      float* __restrict__ gpair, int B, int Nq, int Ns, int way) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= static_cast<int64_t>(B) * Nq) return;
  const int64_t b = wid / Nq;
  const float* lb = labels + b * Ns;
  const float* p = probs + wid * way;
  const float* g = gprobs + wid * way;
  float* out = gpair + wid * Ns;
  if (nanflag != nullptr && nanflag[b] != 0) {
    for (int s = lane; s < Ns; s += 32) out[s] = 0.f;
    return;
  }
  float dot = 0.f;
  for (int c = lane; c < way; c += 32) dot += p[c] * g[c];
  dot = warp_sum(dot);
  for (int s = lane; s < Ns; s += 32) {
    const int cls = static_cast<int>(lb[s]);
    float v = 0.f;
    if (cls >= 0 && cls < way) {
      int cnt = 0;
      for (int j = 0; j < Ns; ++j) cnt += (static_cast<int>(lb[j]) == cls);
      const float gz = p[cls] * (g[cls] - dot);   // d/d z_c, z = -class mean
      v = -gz / cnt;
    }
    out[s] = v;
  }
}

__global__ void div_safe_kernel(const float* __restrict__ num, const float* __restrict__ den,
                                float* __restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = den[i] > 0.f ? num[i] / den[i] : 0.f;
}

int make_shape(DpShape* p, int B, int Nq, int Ns, int L, int M, int64_t ld, float lbda, int single_dir) {
  LMKD_CHECK(L >= 1 && M >= 1 && L <= 32 && M <= 32, "OTAM supports 1..32 frames per clip (got %d, %d)", L, M);
  int G = 8;
  while (G < L || G < M) G <<= 1;
  p->B = B; p->Nq = Nq; p->Ns = Ns; p->L = L; p->M = M; p->G = G; p->ld = ld; p->lbda = lbda;
  p->pairs_per_warp = G == 8 ? 2 : 1;
  p->single_dir = single_dir;
  p->dirs_concurrent = (G <= 16 && !single_dir) ? 2 : 1;
  p->npass = single_dir ? 1 : 2 / p->dirs_concurrent;
  return 0;
}

}  // namespace

int otam_dp_fwd(const float* dist, float* pair, int B, int Nq, int Ns, int L, int M, int64_t ld, float lbda,
                int single_dir, cudaStream_t stream) {
  DpShape p;
  if (int rc = make_shape(&p, B, Nq, Ns, L, M, ld, lbda, single_dir)) return rc;
  const int64_t npairs = static_cast<int64_t>(B) * Nq * Ns;
  LMKD_CHECK(npairs < (1ll << 31), "OTAM: too many (query, support) pairs");
  int64_t blocks = ceil_div(npairs, static_cast<int64_t>(kWarpsPerBlock) * p.pairs_per_warp);
  if (blocks > 16ll * sm_count()) blocks = 16ll * sm_count();
  const size_t smem = sizeof(float) * kWarpsPerBlock * p.pairs_per_warp * L * M;
  // work = DP cells: pairs x directions x L x (M + 1)   (column 0 excluded)
  KernelTimingScope timing(TIME_OTAM_DP, stream, static_cast<double>(npairs) * (single_dir ? 1 : 2) * L * (M + 1));
  if (int rc = timing.begin()) return rc;
  otam_dp_fwd_kernel<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, smem, stream>>>(dist, pair, p);
  LMKD_LAUNCH_CHECK("otam_dp_fwd_kernel");
  return timing.end();
}

int otam_dp_bwd(const float* dist, const float* gpair, const float* nq, const float* ns, __nv_bfloat16* dnum,
                float* gnq, float* gns, float* ddist_raw, int B, int Nq, int Ns, int L, int M, int64_t ld, float lbda,
                float eps, int single_dir, cudaStream_t stream) {
  DpShape p;
  if (int rc = make_shape(&p, B, Nq, Ns, L, M, ld, lbda, single_dir)) return rc;
  const int64_t npairs = static_cast<int64_t>(B) * Nq * Ns;
  LMKD_CHECK(npairs < (1ll << 31), "OTAM: too many (query, support) pairs");
  int64_t blocks = ceil_div(npairs, static_cast<int64_t>(kWarpsPerBlock) * p.pairs_per_warp);
  const size_t smem = sizeof(float) * kWarpsPerBlock * p.pairs_per_warp * bwd_pair_floats(L, M, p.G);
  {  // as many blocks as fit at once: shared memory or 64 warps per SM, whichever binds
    const int64_t per_sm = std::max<int64_t>(1, std::min<int64_t>(16, (200 * 1024) / static_cast<int64_t>(smem + 1024)));
    if (blocks > per_sm * sm_count()) blocks = per_sm * sm_count();
  }
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(otam_dp_bwd_kernel), 160 * 1024)) return rc;
  // the backward repeats the forward sweep (storing the soft-min weights) and then walks it in reverse: 2 x cells
  KernelTimingScope timing(TIME_OTAM_DP, stream, 2.0 * static_cast<double>(npairs) * (single_dir ? 1 : 2) * L * (M + 1));
  if (int rc = timing.begin()) return rc;
  otam_dp_bwd_kernel<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, smem, stream>>>(dist, gpair, nq, ns, dnum,
                                                                                          gnq, gns, ddist_raw, p, eps);
  LMKD_LAUNCH_CHECK("otam_dp_bwd_kernel");
  return timing.end();
}

int otam_class_fwd(const float* pair, const float* labels, const int* nanflag, float* probs, int B, int Nq, int Ns,
                   int way, int* status, cudaStream_t stream) {
  const int64_t warps = static_cast<int64_t>(B) * Nq;
  otam_class_fwd_kernel<<<static_cast<unsigned>(ceil_div(warps * 32, 128)), 128, 0, stream>>>(
      pair, labels, nanflag, probs, B, Nq, Ns, way, status);
  LMKD_LAUNCH_CHECK("otam_class_fwd_kernel");
  return 0;
}

int otam_class_bwd(const float* gprobs, const float* probs, const float* labels, const int* nanflag, float* gpair,
                   int B, int Nq, int Ns, int way, cudaStream_t stream) {
  const int64_t warps = static_cast<int64_t>(B) * Nq;
  otam_class_bwd_kernel<<<static_cast<unsigned>(ceil_div(warps * 32, 128)), 128, 0, stream>>>(
      gprobs, probs, labels, nanflag, gpair, B, Nq, Ns, way);
  LMKD_LAUNCH_CHECK("otam_class_bwd_kernel");
  return 0;
}

// grid (B, blocks per episode): blocks of unflagged episodes exit on their first instruction
__global__ void __launch_bounds__(256)
zero_flagged_kernel(const int* __restrict__ nanflag, float* __restrict__ gq, float* __restrict__ gs, int64_t nq,
                    int64_t ns) {
  const int64_t b = blockIdx.x;
  if (nanflag[b] == 0) return;
  float4* q4 = reinterpret_cast<float4*>(gq + b * nq);
  float4* s4 = reinterpret_cast<float4*>(gs + b * ns);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t stride = static_cast<int64_t>(gridDim.y) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.y) * blockDim.x + threadIdx.x; i < nq / 4; i += stride) q4[i] = z;
  for (int64_t i = static_cast<int64_t>(blockIdx.y) * blockDim.x + threadIdx.x; i < ns / 4; i += stride) s4[i] = z;
}

int otam_zero_flagged(const int* nanflag, float* gq, float* gs, int B, int64_t nq, int64_t ns, cudaStream_t stream) {
  LMKD_CHECK(nq % 4 == 0 && ns % 4 == 0, "otam: feature rows must be multiples of 4 floats");
  zero_flagged_kernel<<<dim3(static_cast<unsigned>(B), 8), 256, 0, stream>>>(nanflag, gq, gs, nq, ns);
  LMKD_LAUNCH_CHECK("zero_flagged_kernel");
  return 0;
}

int div_safe(const float* num, const float* den, float* out, int64_t n, cudaStream_t stream) {
  div_safe_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, stream>>>(num, den, out, n);
  LMKD_LAUNCH_CHECK("div_safe_kernel");
  return 0;
}

}  // namespace lmkd
