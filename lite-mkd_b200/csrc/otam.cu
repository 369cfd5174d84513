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

namespace lmkd {

namespace {

constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ float softmin2(float a, float b, float inv_l, float lbda) {
  const float mn = fminf(a, b);
  const float s = __expf(-(a - mn) * inv_l) + __expf(-(b - mn) * inv_l);
  return mn - lbda * __logf(s);
}
__device__ __forceinline__ float softmin3(float a, float b, float c, float inv_l, float lbda) {
  const float mn = fminf(a, fminf(b, c));
  const float s = __expf(-(a - mn) * inv_l) + __expf(-(b - mn) * inv_l) + __expf(-(c - mn) * inv_l);
  return mn - lbda * __logf(s);
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

// Loads the L x M block of pair (q, s) into smem (row-major, pitch M).
__device__ __forceinline__ void load_block(const float* __restrict__ dist, float* blk, int L, int M,
                                           int64_t ld, int lane_in, int nlanes) {
  for (int i = lane_in; i < L * M; i += nlanes) {
    const int l = i / M, m = i - l * M;
    blk[i] = __ldg(dist + static_cast<int64_t>(l) * ld + m);
  }
}

// One DP sweep of one task by a group of G lanes.  `R` rows (this lane is row r), `Cn` columns;
// dval(r, c) = blk[r * sr + c * sc].  If `table` != nullptr the full padded table
// [R][Cn + 2] is stored (for the backward).  Returns C[R-1][Cn+1] in lane R-1 (garbage elsewhere).
__device__ __forceinline__ float dp_sweep(const float* blk, int sr, int sc, int R, int Cn, int r, int G,
                                          float lbda, float* table, int steps) {
  // `steps` (= L + M for either direction) is warp-uniform; lanes without a task pass R = 0.
  const float inv_l = 1.f / lbda;
  float cur = 0.f, prevcur = 0.f;
  if (table != nullptr && r < R) table[r * (Cn + 2)] = 0.f;
  for (int t = 1; t <= steps; ++t) {
    const float up1 = __shfl_up_sync(0xffffffffu, cur, 1, G);      // C[r-1][m]
    const float up2 = __shfl_up_sync(0xffffffffu, prevcur, 1, G);  // C[r-1][m-1]
    const int m = t - r;
    if (r < R && m >= 1 && m <= Cn + 1) {
      const float dv = (m <= Cn) ? blk[r * sr + (m - 1) * sc] : 0.f;
      float nv;
      if (r == 0) nv = dv + cur;
      else if (m == 1 || m == Cn + 1) nv = dv + softmin3(up2, up1, cur, inv_l, lbda);
      else nv = dv + softmin2(up2, cur, inv_l, lbda);
      prevcur = cur;
      cur = nv;
      if (table != nullptr) table[r * (Cn + 2) + m] = nv;
    }
  }
  return cur;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
otam_dp_fwd_kernel(const float* __restrict__ dist, float* __restrict__ pair, const DpShape p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk_elems = p.L * p.M;
  float* wblk = smem + warp * p.pairs_per_warp * blk_elems;
  const int64_t npairs = static_cast<int64_t>(p.B) * p.Nq * p.Ns;
  const int64_t first = (static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + warp) * p.pairs_per_warp;
  const int lanes_per_pair = 32 / p.pairs_per_warp;
  const int pw = lane / lanes_per_pair;          // which pair of this warp
  const int lp = lane - pw * lanes_per_pair;     // lane within the pair's lanes
  const int64_t pid = first + pw;
  const bool valid = pid < npairs;
  int64_t b = 0;
  int q = 0, s = 0;
  if (valid) {
    s = static_cast<int>(pid % p.Ns);
    q = static_cast<int>((pid / p.Ns) % p.Nq);
    b = pid / (static_cast<int64_t>(p.Ns) * p.Nq);
  }
  float* blk = wblk + pw * blk_elems;
  if (valid) {
    const float* src = dist + (b * p.Nq * p.L + static_cast<int64_t>(q) * p.L) * p.ld + static_cast<int64_t>(s) * p.M;
    load_block(src, blk, p.L, p.M, p.ld, lp, lanes_per_pair);
  }
  __syncwarp();
  float total = 0.f;
  const int dir_of_lane = lp / p.G;   // 0 or 1 when both directions run concurrently
  const int r = lp - dir_of_lane * p.G;
  for (int pass = 0; pass < p.npass; ++pass) {
    const int dir = p.dirs_concurrent == 2 ? dir_of_lane : pass;
    // dir 0: rows = query frames (L), cols = support frames (M); dir 1: transposed
    const int R = dir == 0 ? p.L : p.M, Cn = dir == 0 ? p.M : p.L;
    const int sr = dir == 0 ? p.M : 1, sc = dir == 0 ? 1 : p.M;
    const bool lane_has_task = lp < p.G * p.dirs_concurrent;
    const float res = dp_sweep(blk, sr, sc, lane_has_task ? R : 0, Cn, r, p.G, p.lbda, nullptr, p.L + p.M);
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
  if (valid && lp == 0) pair[pid] = total;
}

// backward: one pair per warp-slot as in the forward; smem per pair:
//   blk [L*M] | dd [L*M] (d loss / d dist, both directions) | per direction: C table, G table
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
otam_dp_bwd_kernel(const float* __restrict__ dist, const float* __restrict__ gpair,
                   const float* __restrict__ nq, const float* __restrict__ ns,
                   __nv_bfloat16* __restrict__ dnum, float* __restrict__ gnq, float* __restrict__ gns,
                   float* __restrict__ ddist_raw, const DpShape p, float eps) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk_elems = p.L * p.M;
  const int tabA = p.L * (p.M + 2), tabB = p.M * (p.L + 2);
  const int tab = tabA > tabB ? tabA : tabB;
  const int per_pair = 2 * blk_elems + 4 * tab;
  const int64_t npairs = static_cast<int64_t>(p.B) * p.Nq * p.Ns;
  const int64_t first = (static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + warp) * p.pairs_per_warp;
  const int lanes_per_pair = 32 / p.pairs_per_warp;
  const int pw = lane / lanes_per_pair;
  const int lp = lane - pw * lanes_per_pair;
  const int64_t pid = first + pw;
  const bool valid = pid < npairs;
  int64_t b = 0;
  int q = 0, s = 0;
  if (valid) {
    s = static_cast<int>(pid % p.Ns);
    q = static_cast<int>((pid / p.Ns) % p.Nq);
    b = pid / (static_cast<int64_t>(p.Ns) * p.Nq);
  }
  float* base = smem + (warp * p.pairs_per_warp + pw) * per_pair;
  float* blk = base;
  float* dd = base + blk_elems;
  const int64_t row0 = b * p.Nq * p.L + static_cast<int64_t>(q) * p.L;
  const int64_t col0 = static_cast<int64_t>(s) * p.M;
  if (valid) load_block(dist + row0 * p.ld + col0, blk, p.L, p.M, p.ld, lp, lanes_per_pair);
  for (int i = lp; i < blk_elems; i += lanes_per_pair) dd[i] = 0.f;
  __syncwarp();
  const float gout = valid ? __ldg(gpair + pid) : 0.f;
  const float inv_l = 1.f / p.lbda;
  const int dir_of_lane = lp / p.G;
  const int r = lp - dir_of_lane * p.G;
  for (int pass = 0; pass < p.npass; ++pass) {
    const int dir = p.dirs_concurrent == 2 ? dir_of_lane : pass;
    const int R = dir == 0 ? p.L : p.M, Cn = dir == 0 ? p.M : p.L;
    const int sr = dir == 0 ? p.M : 1, sc = dir == 0 ? 1 : p.M;
    const bool lane_has_task = lp < p.G * p.dirs_concurrent;
    const int slot = (p.dirs_concurrent == 2 && lane_has_task) ? dir_of_lane : 0;
    float* Ct = base + 2 * blk_elems + (2 * slot) * tab;
    float* Gt = Ct + tab;
    const int W = Cn + 2;
    dp_sweep(blk, sr, sc, lane_has_task ? R : 0, Cn, r, p.G, p.lbda, lane_has_task ? Ct : nullptr, p.L + p.M);
    if (lane_has_task && r < R)
      for (int m = 0; m < W; ++m) Gt[r * W + m] = 0.f;
    __syncwarp();
    if (lane_has_task && r == R - 1) Gt[r * W + Cn + 1] = gout;
    __syncwarp();
    for (int t = R + Cn; t >= 1; --t) {
      const int m = t - r;
      const bool act = lane_has_task && r < R && m >= 1 && m <= Cn + 1;
      float g = 0.f, wa = 0.f, wb = 0.f, wc = 0.f;
      bool three = false;
      if (act) {
        g = Gt[r * W + m];
        if (m <= Cn) atomicAdd(&dd[r * sr + (m - 1) * sc], g);
        if (r == 0) {
          wb = 1.f;  // plain running sum along the top row
        } else {
          three = (m == 1 || m == Cn + 1);
          const float a = Ct[(r - 1) * W + m - 1];
          const float bb = Ct[r * W + m - 1];
          const float c = three ? Ct[(r - 1) * W + m] : 0.f;
          float mn = fminf(a, bb);
          if (three) mn = fminf(mn, c);
          const float ea = __expf(-(a - mn) * inv_l), eb = __expf(-(bb - mn) * inv_l);
          const float ec = three ? __expf(-(c - mn) * inv_l) : 0.f;
          const float inv = 1.f / (ea + eb + ec);
          wa = ea * inv;
          wb = eb * inv;
          wc = ec * inv;
        }
        // phase 1: left neighbour (own row) and diagonal (row above): distinct cells across lanes
        Gt[r * W + m - 1] += g * wb;
        if (r > 0) Gt[(r - 1) * W + m - 1] += g * wa;
      }
      __syncwarp();
      // phase 2: the cell straight above (three-way cells only)
      if (act && three) Gt[(r - 1) * W + m] += g * wc;
      __syncwarp();
    }
  }
  __syncwarp();
  if (ddist_raw != nullptr) {   // test hook: raw d loss / d dist, no cosine chain
    if (valid)
      for (int i = lp; i < blk_elems; i += lanes_per_pair)
        ddist_raw[(row0 + i / p.M) * p.ld + col0 + i % p.M] = dd[i];
    return;
  }
  // d dist -> d numerator (bf16) and norm gradients.  dist = 1 - num / (|x||y| + eps)
  if (valid) {
    for (int i = lp; i < blk_elems; i += lanes_per_pair) {
      const int l = i / p.M, m = i - l * p.M;
      const float g = dd[i];
      const float nx = __ldg(nq + row0 + l), ny = __ldg(ns + b * p.Ns * p.M + col0 + m);
      const float den = nx * ny + eps;
      const float sim = 1.f - blk[i];
      dnum[(row0 + l) * p.ld + col0 + m] = __float2bfloat16_rn(-g / den);
      const float tt = g * sim / den;   // dL/d(den)
      blk[i] = tt * ny;                 // contribution to d|x_l|
      dd[i] = tt * nx;                  // contribution to d|y_m|
    }
  }
  __syncwarp();
  if (valid) {
    for (int l = lp; l < p.L; l += lanes_per_pair) {
      float acc = 0.f;
      for (int m = 0; m < p.M; ++m) acc += blk[l * p.M + m];
      atomicAdd(gnq + row0 + l, acc);
    }
    for (int m = lp; m < p.M; m += lanes_per_pair) {
      float acc = 0.f;
      for (int l = 0; l < p.L; ++l) acc += dd[l * p.M + m];
      atomicAdd(gns + b * p.Ns * p.M + col0 + m, acc);
    }
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
  const int64_t blocks = ceil_div(npairs, static_cast<int64_t>(kWarpsPerBlock) * p.pairs_per_warp);
  const size_t smem = sizeof(float) * kWarpsPerBlock * p.pairs_per_warp * L * M;
  otam_dp_fwd_kernel<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, smem, stream>>>(dist, pair, p);
  LMKD_LAUNCH_CHECK("otam_dp_fwd_kernel");
  return 0;
}

int otam_dp_bwd(const float* dist, const float* gpair, const float* nq, const float* ns, __nv_bfloat16* dnum,
                float* gnq, float* gns, float* ddist_raw, int B, int Nq, int Ns, int L, int M, int64_t ld, float lbda,
                float eps, int single_dir, cudaStream_t stream) {
  DpShape p;
  if (int rc = make_shape(&p, B, Nq, Ns, L, M, ld, lbda, single_dir)) return rc;
  const int64_t npairs = static_cast<int64_t>(B) * Nq * Ns;
  const int64_t blocks = ceil_div(npairs, static_cast<int64_t>(kWarpsPerBlock) * p.pairs_per_warp);
  const int tabA = L * (M + 2), tabB = M * (L + 2);
  const int tab = tabA > tabB ? tabA : tabB;
  const size_t smem = sizeof(float) * kWarpsPerBlock * p.pairs_per_warp * (2 * L * M + 4 * tab);
  static bool attr_set = false;
  if (!attr_set) {
    LMKD_CUDA(cudaFuncSetAttribute(otam_dp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  otam_dp_bwd_kernel<<<static_cast<unsigned>(blocks), kWarpsPerBlock * 32, smem, stream>>>(dist, gpair, nq, ns, dnum,
                                                                                          gnq, gns, ddist_raw, p, eps);
  LMKD_LAUNCH_CHECK("otam_dp_bwd_kernel");
  return 0;
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

int div_safe(const float* num, const float* den, float* out, int64_t n, cudaStream_t stream) {
  div_safe_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, stream>>>(num, den, out, n);
  LMKD_LAUNCH_CHECK("div_safe_kernel");
  return 0;
}

}  // namespace lmkd
