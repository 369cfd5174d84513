// extern "C" entry points declared in include/lmkd.h: argument checking, workspace carving and
// the kernel sequences of each operator.  No device memory is allocated here.
#include "../../include/lmkd.h"

#include <cmath>
#include <new>

#include "gemm.cuh"
#include "feature_head.cuh"
#include "feature_store.cuh"
#include "fusion.cuh"
#include "heads.cuh"
#include "loss.cuh"
#include "otam.cuh"
#include "prep.cuh"
#include "trx.cuh"
#include "strm.cuh"
#include "trx_attn.cuh"

using namespace lmkd;

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// bump allocator over a caller-provided workspace (256-byte aligned slices)
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <typename T>
  T* take(int64_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += static_cast<size_t>(count) * sizeof(T);
    return r;
  }
  size_t total() const { return (off + 255) & ~static_cast<size_t>(255); }
};

// LMKD_OTAM_AUX16=0: the OTAM gradient products read the fp32 features as their AXPY operand (A/B measurements)
const bool g_otam_aux16 = [] {
  const char* e = getenv("LMKD_OTAM_AUX16");
  return !(e && e[0] == '0');
}();

// ---------------------------------------------------------------------------------------------
struct OtamWs {
  int* nanflag;
  __nv_bfloat16 *xq, *xs, *dnum;
  float *nq, *ns, *dist, *pair, *gpair, *gnq, *gns, *rvq, *rvs;
  int64_t ld;
  size_t bytes;
};

OtamWs otam_layout(void* ws, int B, int Ns, int Nq, int L, int D) {
  Carver c(ws);
  OtamWs w;
  const int64_t rq = static_cast<int64_t>(B) * Nq * L, rs = static_cast<int64_t>(B) * Ns * L;
  w.ld = round_up(static_cast<int64_t>(Ns) * L, 8);
  w.nanflag = c.take<int>(B);
  w.xq = c.take<__nv_bfloat16>(rq * D);
  w.xs = c.take<__nv_bfloat16>(rs * D);
  w.nq = c.take<float>(rq);
  w.ns = c.take<float>(rs);
  w.dist = c.take<float>(rq * w.ld);
  w.pair = c.take<float>(static_cast<int64_t>(B) * Nq * Ns);
  w.gpair = c.take<float>(static_cast<int64_t>(B) * Nq * Ns);
  w.dnum = c.take<__nv_bfloat16>(rq * w.ld);
  w.gnq = c.take<float>(rq);
  w.gns = c.take<float>(rs);
  w.rvq = c.take<float>(rq);
  w.rvs = c.take<float>(rs);
  w.bytes = c.total();
  return w;
}

int sim_gemm(const __nv_bfloat16* xq, const __nv_bfloat16* xs, const float* nq, const float* ns, float* dist,
             int64_t ld, int B, int nx, int ny, int D, float eps, cudaStream_t st) {
  GemmDesc g;
  g.M = nx; g.N = ny; g.K = D; g.nb1 = 1; g.nb2 = B;
  g.A.ptr = xq; g.A.ld = D; g.A.stride_b2 = static_cast<int64_t>(nx) * D;
  g.B.ptr = xs; g.B.ld = D; g.B.stride_b2 = static_cast<int64_t>(ny) * D;
  g.epi.kind = EPI_COSDIST; g.epi.eps = eps;
  g.epi.C = dist; g.epi.ldc = ld; g.epi.c_b2 = static_cast<int64_t>(nx) * ld;
  g.epi.rowv = nq; g.epi.rv_b2 = nx;
  g.epi.colv = ns; g.epi.cv_b2 = ny;
  return gemm_bf16(g, st);
}

// ---------------------------------------------------------------------------------------------
struct TrxWs {
  int *slot, *cnt;
  int* flags;                  // device scratch ints (kernel hand-over flags)
  uint64_t* seed_used;
  __nv_bfloat16 *xb, *wcat, *kq, *vq, *ks, *vs, *patt, *dq;
  float *P, *stats, *scores, *rowred;
  // backward
  __nv_bfloat16 *ps, *dS, *dpcat;
  float *srow, *dP, *dKq, *dKs, *dVs, *dxk, *dxv, *partials, *dWcat, *dX, *gram, *lnred_q, *lnred_s;
  float *rowdot, *linv, *rs;   // fused attention: <diff, prototype>, 1 / rowsum, srow / rowsum per (b, class, row)
  __nv_bfloat16* E;
  bool fused;                  // scores / probabilities stay in tensor memory (trx_attn.cu)
  int qchunk;                  // materialised path: queries per pass (== Nq: one pass, probabilities kept)
  bool ln_fused;               // LayerNorm backward fused with the gather (needs one-pass dK products)
  bool g16;                    // dKq / dKs / dVs hold bf16 (one-pass products feeding the fused LayerNorm-backward kernels)
  int max_partial_blocks;
  size_t bytes;
};

// LMKD_TRX_ATTN_BYTES: byte budget of the materialised score / probability buffers (default 24 GB of the 180)
double g_attn_budget_override = 0.0;     // lmkd_trx_set_attn_budget(); read by layout, forward and backward alike
double trx_attn_budget_bytes() {
  static const double v = [] {
    const char* e = getenv("LMKD_TRX_ATTN_BYTES");
    const double x = e ? atof(e) : 0.0;
    return x > 0.0 ? x : 24.0e9;
  }();
  return g_attn_budget_override > 0.0 ? g_attn_budget_override : v;
}

// LMKD_TRX_DX_FUSED=0: dX through a buffer and the trx_dx_scatter kernel (A/B measurements)
const bool g_dx_fused = [] {
  const char* e = getenv("LMKD_TRX_DX_FUSED");
  return !(e && e[0] == '0');
}();
// LMKD_TRX_G16=0: fp32 dK / dV rows on every path (A/B measurements)
const bool g_grad_rows_bf16 = [] {
  const char* e = getenv("LMKD_TRX_G16");
  return !(e && e[0] == '0');
}();
// LMKD_TRX_DV_T=0: dV through the [KTp x d] formulation (padded row tiles) for A/B measurements
const bool g_dv_transposed = [] {
  const char* e = getenv("LMKD_TRX_DV_T");
  return !(e && e[0] == '0');
}();

int trx_dims(const lmkd_trx_shape* s, TrxDims* d) {
  LMKD_CHECK(s != nullptr, "null shape");
  LMKD_CHECK(s->B > 0 && s->Ns > 0 && s->Nq > 0 && s->L > 0 && s->D > 0 && s->d > 0, "trx: empty shape");
  LMKD_CHECK(s->card >= 1 && s->card <= 4 && s->card <= s->L, "trx: cardinality %d unsupported (L = %d)", s->card, s->L);
  LMKD_CHECK(s->way >= 1 && s->shot >= 1, "trx: way/shot must be positive");
  LMKD_CHECK(s->D % 8 == 0 && s->d % 8 == 0, "trx: D (%d) and d (%d) must be multiples of 8", s->D, s->d);
  double T = 1;
  for (int i = 0; i < s->card; ++i) T = T * (s->L - i) / (i + 1);
  d->B = s->B; d->Ns = s->Ns; d->Nq = s->Nq; d->L = s->L; d->D = s->D; d->d = s->d; d->card = s->card;
  d->way = s->way; d->shot = s->shot;
  d->T = static_cast<int>(std::llround(T));
  d->N = s->Ns + s->Nq;
  d->KT = s->shot * d->T;
  d->KTp = static_cast<int>(round_up(d->KT, 16));
  d->NqT = s->Nq * d->T;
  d->NqT_full = d->NqT;
  d->m_off = 0;
  d->M = static_cast<int64_t>(s->B) * d->N * s->L;
  d->R = static_cast<int64_t>(s->B) * d->N * d->T;
  // the GEMM descriptors carry 32-bit extents: refuse shapes that would truncate instead of mis-addressing
  const int64_t lim = (1ll << 31) - 1;
  LMKD_CHECK(T < 1e9 && d->M <= lim && 2ll * s->card * s->d <= lim && static_cast<int64_t>(s->way) * d->KTp <= lim &&
                 static_cast<int64_t>(s->Nq) * d->T <= lim && static_cast<int64_t>(s->shot) * d->T <= lim,
             "trx: shape too large for 32-bit tile extents (B*N*L = %lld rows)", (long long)d->M);
  return 0;
}

TrxWs trx_layout(void* ws, const TrxDims& s, int need_grad) {
  Carver c(ws);
  TrxWs w;
  memset(&w, 0, sizeof(w));
  const int64_t pcols = 2ll * s.card * s.d;
  const int64_t qrows = static_cast<int64_t>(s.B) * s.NqT;
  const int64_t srows = static_cast<int64_t>(s.B) * s.way * s.KTp;
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  w.slot = c.take<int>(static_cast<int64_t>(s.B) * s.Ns);
  w.cnt = c.take<int>(static_cast<int64_t>(s.B) * s.way);
  w.flags = c.take<int>(4);
  w.seed_used = c.take<uint64_t>(1);
  w.xb = c.take<__nv_bfloat16>(s.M * s.D);
  w.wcat = c.take<__nv_bfloat16>(pcols * s.D);
  w.P = c.take<float>(s.M * pcols);
  w.stats = c.take<float>(s.R * 2);
  w.kq = c.take<__nv_bfloat16>(qrows * s.d);
  w.vq = c.take<__nv_bfloat16>(qrows * s.d);
  w.ks = c.take<__nv_bfloat16>(srows * s.d);
  w.vs = c.take<__nv_bfloat16>(srows * s.d);
  // need_grad 2 (TRX_sup: gradient through the prototype similarities) keeps the materialised pipeline
  w.fused = trx_attn_fused_fits(s) && need_grad != 2;
  // Materialised path (class groups wider than tensor memory, or TRX_sup): scores / probabilities of ALL queries
  // would be B * Nq*T * way*KTp entries (6e10 per episode for 32-frame triples).  They are produced for `qchunk`
  // queries at a time inside a fixed byte budget; with more than one pass the backward recomputes them.
  w.qchunk = s.Nq;
  if (!w.fused) {
    const double per_query = static_cast<double>(s.B) * s.T * pitch * (need_grad ? 10.0 : 6.0);   // f32 + bf16 (+ 2 bf16)
    const double fit = trx_attn_budget_bytes() / per_query;
    if (fit < s.Nq) w.qchunk = fit < 1.0 ? 1 : static_cast<int>(fit);
  }
  const int64_t crows = static_cast<int64_t>(s.B) * w.qchunk * s.T;      // score rows held at a time
  w.ln_fused = trx_bwd_fused_fits(s) && w.qchunk == s.Nq;
  if (!w.fused) w.scores = c.take<float>(crows * pitch);
  if (!w.fused) w.patt = c.take<__nv_bfloat16>(crows * pitch);
  else if (need_grad) w.patt = c.take<__nv_bfloat16>(qrows * pitch);
  if (w.fused && need_grad) {
    w.rowdot = c.take<float>(static_cast<int64_t>(s.B) * s.way * s.NqT);
    w.linv = c.take<float>(static_cast<int64_t>(s.B) * s.way * s.NqT);
    w.rs = c.take<float>(static_cast<int64_t>(s.B) * s.way * s.NqT);
  }
  w.dq = c.take<__nv_bfloat16>(static_cast<int64_t>(s.B) * s.way * s.NqT * s.d);
  w.rowred = c.take<float>(static_cast<int64_t>(s.B) * s.way * s.NqT);
  w.gram = c.take<float>(static_cast<int64_t>(s.B) * s.Nq * s.way * s.way);
  if (need_grad) {
    w.srow = c.take<float>(static_cast<int64_t>(s.B) * s.way * s.NqT);
    w.ps = c.take<__nv_bfloat16>((w.fused ? qrows : crows) * pitch);
    w.dP = w.scores;  // the score buffer is dead after the softmax; reuse it for dP (null when fused)
    w.dS = c.take<__nv_bfloat16>((w.fused ? qrows : crows) * pitch);
    // The tuple-row gradients are written once by a GEMM epilogue and read once by the LayerNorm-backward + gather
    // kernel: on the one-pass path they are stored as bf16 (half the bytes of an HBM-bound hand-over; the LayerNorm row
    // reductions are taken from the fp32 accumulators before rounding).  Multi-pass paths accumulate and keep fp32.
    w.g16 = w.ln_fused && g_grad_rows_bf16;
    if (w.g16) {
      w.dKq = reinterpret_cast<float*>(c.take<__nv_bfloat16>(qrows * s.d));
      w.dKs = reinterpret_cast<float*>(c.take<__nv_bfloat16>(srows * s.d));
      w.dVs = reinterpret_cast<float*>(c.take<__nv_bfloat16>(srows * s.d));
    } else {
      w.dKq = c.take<float>(qrows * s.d);
      w.dKs = c.take<float>(srows * s.d);
      w.dVs = c.take<float>(srows * s.d);
    }
    w.lnred_q = c.take<float>(qrows * 2);
    w.lnred_s = c.take<float>(srows * 2);
    if (!w.ln_fused) {
      w.dxk = c.take<float>(s.R * s.d);
      w.dxv = c.take<float>(s.R * s.d);
    }
    w.max_partial_blocks = 4 * 160;
    w.partials = c.take<float>(static_cast<int64_t>(w.max_partial_blocks) * 4 * s.d);
    w.dpcat = c.take<__nv_bfloat16>(s.M * pcols);
    w.dWcat = c.take<float>(pcols * s.D);
    w.dX = c.take<float>(s.M * s.D);
    if (need_grad > 1)   // TRX_sup: total prototype gradient, same shape as dq
      w.E = c.take<__nv_bfloat16>(static_cast<int64_t>(s.B) * s.way * s.NqT * s.d);
  }
  w.bytes = c.total();
  return w;
}


// ---------------------------------------------------------------------------------------------
struct StrmWs {
  int *slot, *cnt;
  uint64_t* seed_used;
  float *pe0, *P, *nq2, *ns2, *dEq, *dEs, *dWcat, *dX;
  __nv_bfloat16 *xb, *wcat, *eq, *es, *dpcat;
  unsigned long long* best;
  size_t bytes;
};

StrmWs strm_layout(void* ws, const TrxDims& s, int need_grad) {
  Carver c(ws);
  StrmWs w;
  memset(&w, 0, sizeof(w));
  const int64_t pcols = static_cast<int64_t>(s.card) * s.d;
  const int64_t qrows = static_cast<int64_t>(s.B) * s.NqT;
  const int64_t srows = static_cast<int64_t>(s.B) * s.way * s.KTp;
  w.slot = c.take<int>(static_cast<int64_t>(s.B) * s.Ns);
  w.cnt = c.take<int>(static_cast<int64_t>(s.B) * s.way);
  w.seed_used = c.take<uint64_t>(1);
  w.pe0 = c.take<float>(static_cast<int64_t>(s.L) * s.D);
  w.xb = c.take<__nv_bfloat16>(s.M * s.D);
  w.wcat = c.take<__nv_bfloat16>(pcols * s.D);
  w.P = c.take<float>(s.M * pcols);
  w.eq = c.take<__nv_bfloat16>(qrows * s.d);
  w.es = c.take<__nv_bfloat16>(srows * s.d);
  w.nq2 = c.take<float>(qrows);
  w.ns2 = c.take<float>(srows);
  w.best = c.take<unsigned long long>(static_cast<int64_t>(s.B) * s.way * s.NqT);
  if (need_grad) {
    w.dEq = c.take<float>(qrows * s.d);
    w.dEs = c.take<float>(srows * s.d);
    w.dpcat = c.take<__nv_bfloat16>(s.M * pcols);
    w.dWcat = c.take<float>(pcols * s.D);
    w.dX = c.take<float>(s.M * s.D);
  }
  w.bytes = c.total();
  return w;
}

// Materialised attention for queries [q0, q0 + nq) of every episode: scores -> class-grouped softmax ->
// prototype distance.  Row offsets address the chunk inside the full [B, Nq*T, .] tensors; the score and
// probability buffers hold only the chunk.
TrxDims trx_chunk_dims(const TrxDims& s, int q0, int nq) {
  TrxDims c = s;
  c.Nq = nq;
  c.NqT = nq * s.T;
  c.NqT_full = s.NqT;
  c.m_off = q0 * s.T;
  return c;
}

int trx_scores_softmax(const TrxWs& w, const TrxDims& s, const TrxDims& c, cudaStream_t st) {
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  {  // scores[b][m][(c, kt)] = <kq, ks> / sqrt(d)       (TRX.py:125)
    GemmDesc g;
    g.M = c.NqT; g.N = static_cast<int>(pitch); g.K = s.d; g.nb2 = s.B;
    g.A.ptr = w.kq + static_cast<int64_t>(c.m_off) * s.d; g.A.ld = s.d; g.A.stride_b2 = static_cast<int64_t>(s.NqT) * s.d;
    g.B.ptr = w.ks; g.B.ld = s.d; g.B.stride_b2 = pitch * s.d;
    g.epi.kind = EPI_STORE_F32; g.epi.alpha = 1.f / sqrtf(static_cast<float>(s.d));
    g.epi.C = w.scores; g.epi.ldc = pitch; g.epi.c_b2 = static_cast<int64_t>(c.NqT) * pitch;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  return trx_softmax_fwd(w.scores, w.cnt, w.patt, c, st);
}

int trx_attn_chunk_fwd(const TrxWs& w, const TrxDims& s, int q0, int nq, __nv_bfloat16* dq, cudaStream_t st) {
  const TrxDims c = trx_chunk_dims(s, q0, nq);
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  if (int rc = trx_scores_softmax(w, s, c, st)) return rc;
  // per class: proto = P_c . V_c ; diff = v_q - proto ; rowred = |diff|^2   (TRX.py:137-141)
  GemmDesc g;
  g.M = c.NqT; g.N = s.d; g.K = s.KTp; g.nb1 = s.way; g.nb2 = s.B;
  g.A.ptr = w.patt; g.A.ld = pitch; g.A.stride_b1 = s.KTp; g.A.stride_b2 = static_cast<int64_t>(c.NqT) * pitch;
  g.B.ptr = w.vs; g.B.mn_major = 1; g.B.ld = s.d; g.B.stride_b1 = static_cast<int64_t>(s.KTp) * s.d;
  g.B.stride_b2 = pitch * s.d;
  g.epi.kind = EPI_DIFF_SQ;
  g.epi.C = dq ? dq + static_cast<int64_t>(c.m_off) * s.d : nullptr;   // diff rows: backward and TRX_sup only
  g.epi.ldc = s.d; g.epi.c_b1 = static_cast<int64_t>(s.NqT) * s.d;
  g.epi.c_b2 = static_cast<int64_t>(s.way) * s.NqT * s.d;
  g.epi.aux = w.vq + static_cast<int64_t>(c.m_off) * s.d; g.epi.ldaux = s.d; g.epi.aux_b1 = 0;
  g.epi.aux_b2 = static_cast<int64_t>(s.NqT) * s.d;
  g.epi.rowred = w.rowred + c.m_off; g.epi.rr_b1 = s.NqT; g.epi.rr_b2 = static_cast<int64_t>(s.way) * s.NqT;
  return gemm_bf16(g, st);
}

}  // namespace

extern "C" {

const char* lmkd_last_error(void) { return get_error(); }
int lmkd_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
int64_t lmkd_sim_pitch(int64_t ny) { return round_up(ny, 8); }

size_t lmkd_sim_workspace_bytes(int B, int nx, int ny, int D) {
  Carver c(nullptr);
  c.take<__nv_bfloat16>(static_cast<int64_t>(B) * nx * D);
  c.take<__nv_bfloat16>(static_cast<int64_t>(B) * ny * D);
  c.take<float>(static_cast<int64_t>(B) * nx);
  c.take<float>(static_cast<int64_t>(B) * ny);
  return c.total();
}

int lmkd_sim_fwd(const float* x, const float* y, int B, int nx, int ny, int D, float eps, float* dist,
                 void* workspace, void* stream) {
  LMKD_CHECK(x && y && dist && workspace, "sim_fwd: null pointer");
  Carver c(workspace);
  __nv_bfloat16* xb = c.take<__nv_bfloat16>(static_cast<int64_t>(B) * nx * D);
  __nv_bfloat16* yb = c.take<__nv_bfloat16>(static_cast<int64_t>(B) * ny * D);
  float* nxv = c.take<float>(static_cast<int64_t>(B) * nx);
  float* nyv = c.take<float>(static_cast<int64_t>(B) * ny);
  if (int rc = feat_cast_norm(x, xb, nxv, nullptr, static_cast<int64_t>(B) * nx, D, 1, S(stream))) return rc;
  if (int rc = feat_cast_norm(y, yb, nyv, nullptr, static_cast<int64_t>(B) * ny, D, 1, S(stream))) return rc;
  return sim_gemm(xb, yb, nxv, nyv, dist, lmkd_sim_pitch(ny), B, nx, ny, D, eps, S(stream));
}

// ---------------------------------------------------------------------------------------------
size_t lmkd_otam_workspace_bytes(int B, int Ns, int Nq, int L, int D, int way) {
  (void)way;
  return otam_layout(nullptr, B, Ns, Nq, L, D).bytes;
}

int lmkd_otam_fwd(const float* support, const float* labels, const float* query, int B, int Ns, int Nq, int L,
                  int D, int way, float lambda, float eps, float* probs, float* pair_dists, void* workspace,
                  int* status, void* stream) {
  LMKD_CHECK(support && labels && query && probs && workspace, "otam_fwd: null pointer");
  LMKD_CHECK(B > 0 && Ns > 0 && Nq > 0 && way > 0, "otam_fwd: empty episode");
  LMKD_CHECK(lambda > 0.f, "otam_fwd: lambda must be positive");
  cudaStream_t st = S(stream);
  OtamWs w = otam_layout(workspace, B, Ns, Nq, L, D);
  const int64_t rq = static_cast<int64_t>(B) * Nq * L, rs = static_cast<int64_t>(B) * Ns * L;
  LMKD_CUDA(cudaMemsetAsync(w.nanflag, 0, sizeof(int) * B, st));
  // the reference's NaN guard looks at the support features only (model.py:3322)
  if (int rc = feat_cast_norm(support, w.xs, w.ns, w.nanflag, rs, D, static_cast<int64_t>(Ns) * L, st)) return rc;
  if (int rc = feat_cast_norm(query, w.xq, w.nq, nullptr, rq, D, 1, st)) return rc;
  if (int rc = sim_gemm(w.xq, w.xs, w.nq, w.ns, w.dist, w.ld, B, Nq * L, Ns * L, D, eps, st)) return rc;
  if (int rc = otam_dp_fwd(w.dist, w.pair, B, Nq, Ns, L, L, w.ld, lambda, 0, st)) return rc;
  if (pair_dists)
    LMKD_CUDA(cudaMemcpyAsync(pair_dists, w.pair, sizeof(float) * B * Nq * Ns, cudaMemcpyDeviceToDevice, st));
  return otam_class_fwd(w.pair, labels, w.nanflag, probs, B, Nq, Ns, way, status, st);
}

int lmkd_otam_bwd(const float* grad_probs, const float* probs, const float* support, const float* labels,
                  const float* query, int B, int Ns, int Nq, int L, int D, int way, float lambda, float eps,
                  float* grad_support, float* grad_query, void* workspace, void* stream) {
  LMKD_CHECK(grad_probs && probs && support && labels && query && grad_support && grad_query && workspace,
             "otam_bwd: null pointer");
  cudaStream_t st = S(stream);
  OtamWs w = otam_layout(workspace, B, Ns, Nq, L, D);
  const int64_t rq = static_cast<int64_t>(B) * Nq * L, rs = static_cast<int64_t>(B) * Ns * L;
  const int nx = Nq * L, ny = Ns * L;
  if (int rc = otam_class_bwd(grad_probs, probs, labels, w.nanflag, w.gpair, B, Nq, Ns, way, st)) return rc;
  LMKD_CUDA(cudaMemsetAsync(w.gnq, 0, sizeof(float) * rq, st));
  LMKD_CUDA(cudaMemsetAsync(w.gns, 0, sizeof(float) * rs, st));
  if (int rc = otam_dp_bwd(w.dist, w.gpair, w.nq, w.ns, w.dnum, w.gnq, w.gns, nullptr, B, Nq, Ns, L, L, w.ld, lambda,
                           eps, 0, st))
    return rc;
  // d|x| -> coefficient of x itself: d|x|/dx = x / |x|
  if (int rc = div_safe(w.gnq, w.nq, w.rvq, rq, st)) return rc;
  if (int rc = div_safe(w.gns, w.ns, w.rvs, rs, st)) return rc;
  {  // dQ = dnum . Xs + rvq * Q
    GemmDesc g;
    g.M = nx; g.N = D; g.K = ny; g.nb2 = B;
    g.A.ptr = w.dnum; g.A.ld = w.ld; g.A.stride_b2 = static_cast<int64_t>(nx) * w.ld;
    g.B.ptr = w.xs; g.B.mn_major = 1; g.B.ld = D; g.B.stride_b2 = static_cast<int64_t>(ny) * D;
    g.epi.kind = g_otam_aux16 ? EPI_AXPY_B16 : EPI_AXPY_F32; g.epi.alpha = 1.f;
    g.epi.C = grad_query; g.epi.ldc = D; g.epi.c_b2 = static_cast<int64_t>(nx) * D;
    g.epi.rowv = w.rvq; g.epi.rv_b2 = nx;
    // the norm term rv * x reads the bf16 copy of the features the forward made (half the bytes of the fp32 input)
    g.epi.aux = g_otam_aux16 ? static_cast<const void*>(w.xq) : static_cast<const void*>(query);
    g.epi.ldaux = D; g.epi.aux_b2 = static_cast<int64_t>(nx) * D;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  {  // dS = dnum^T . Xq + rvs * S
    GemmDesc g;
    g.M = ny; g.N = D; g.K = nx; g.nb2 = B;
    g.A.ptr = w.dnum; g.A.mn_major = 1; g.A.ld = w.ld; g.A.stride_b2 = static_cast<int64_t>(nx) * w.ld;
    g.B.ptr = w.xq; g.B.mn_major = 1; g.B.ld = D; g.B.stride_b2 = static_cast<int64_t>(nx) * D;
    g.epi.kind = g_otam_aux16 ? EPI_AXPY_B16 : EPI_AXPY_F32; g.epi.alpha = 1.f;
    g.epi.C = grad_support; g.epi.ldc = D; g.epi.c_b2 = static_cast<int64_t>(ny) * D;
    g.epi.rowv = w.rvs; g.epi.rv_b2 = ny;
    g.epi.aux = g_otam_aux16 ? static_cast<const void*>(w.xs) : static_cast<const void*>(support);
    g.epi.ldaux = D; g.epi.aux_b2 = static_cast<int64_t>(ny) * D;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  // NaN guard (model.py:3322-3324): the reference returns detached zeros for such an episode, so no gradient
  // reaches its features; here the NaN supports went through the products above -- overwrite both gradients
  return otam_zero_flagged(w.nanflag, grad_query, grad_support, B, static_cast<int64_t>(nx) * D,
                           static_cast<int64_t>(ny) * D, st);
}

int lmkd_otam_cum_dist(const float* dists, int64_t P, int L, int M, float lambda, float* out, const float* grad_out,
                       float* grad_dists, void* stream) {
  LMKD_CHECK(dists && out, "otam_cum_dist: null pointer");
  LMKD_CHECK(P > 0 && P < (1ll << 31), "otam_cum_dist: bad pair count");
  // each "pair" is its own batch entry: B = P, Nq = Ns = 1, pitch = M
  if (int rc = otam_dp_fwd(dists, out, static_cast<int>(P), 1, 1, L, M, M, lambda, 1, S(stream))) return rc;
  if (grad_out && grad_dists)
    return otam_dp_bwd(dists, grad_out, nullptr, nullptr, nullptr, nullptr, nullptr, grad_dists, static_cast<int>(P), 1, 1,
                       L, M, M, lambda, 0.f, 1, S(stream));
  return 0;
}

// ---------------------------------------------------------------------------------------------
size_t lmkd_trx_workspace_bytes(const lmkd_trx_shape* s, int need_grad) {
  TrxDims d;
  if (trx_dims(s, &d)) return 0;
  return trx_layout(nullptr, d, need_grad).bytes;
}

int lmkd_trx_fwd(const lmkd_trx_shape* sh, const float* support, const float* labels, const float* query,
                 const float* pe, const int32_t* tuples, const float* Wk, const float* bk, const float* Wv,
                 const float* bv, const float* gamma, const float* beta, float* logits, float* proto_sim,
                 void* workspace, int need_grad, int* status, void* stream) {
  TrxDims s;
  if (int rc = trx_dims(sh, &s)) return rc;
  LMKD_CHECK(support && labels && query && pe && tuples && Wk && bk && Wv && bv && gamma && beta && logits && workspace,
             "trx_fwd: null pointer");
  cudaStream_t st = S(stream);
  TrxWs w = trx_layout(workspace, s, need_grad);
  const int64_t pcols = 2ll * s.card * s.d;
  const float ln_eps = sh->ln_eps > 0.f ? sh->ln_eps : 1e-5f;

  if (int rc = trx_class_slots(labels, w.slot, w.cnt, status, s, st)) return rc;
  if (int rc = trx_pe_cast(support, query, pe, w.xb, s.B, s.Ns, s.Nq, s.L, s.D, sh->dropout_p, sh->seed, sh->seed_dev,
                              w.seed_used, st)) return rc;
  if (int rc = trx_pack_weights(Wk, Wv, w.wcat, s, st)) return rc;
  {  // per-frame partial projections: P[M, 2cd] = X~[M, D] . Wcat[2cd, D]^T
    GemmDesc g;
    g.M = static_cast<int>(s.M); g.N = static_cast<int>(pcols); g.K = s.D;
    g.A.ptr = w.xb; g.A.ld = s.D;
    g.B.ptr = w.wcat; g.B.ld = s.D;
    g.epi.kind = EPI_STORE_F32; g.epi.C = w.P; g.epi.ldc = pcols;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  // class-sorted support keys/values carry zero rows for padding and missing shots; every other row
  // is written by the tuple kernel (slots 0..cnt-1 of a class are always filled)
  if (int rc = trx_zero_pad_rows(w.cnt, w.ks, w.vs, s, st)) return rc;
  if (int rc = trx_tuple_ln_fwd(w.P, bk, bv, gamma, beta, tuples, w.slot, w.kq, w.vq, w.ks, w.vs, w.stats, ln_eps, w.flags, s, st))
    return rc;
  if (w.fused) {
    // scores, per-class softmax, prototype and distance in one kernel (TRX.py:125-141); scores and
    // probabilities stay in tensor memory.  Training passes keep exp(score - max) and 1/rowsum for the backward.
    const size_t rbytes = sizeof(float) * s.B * s.way * s.NqT;
    LMKD_CUDA(cudaMemsetAsync(w.rowred, 0, rbytes, st));
    if (need_grad) LMKD_CUDA(cudaMemsetAsync(w.rowdot, 0, rbytes, st));
    TrxAttnFwd a{};
    a.kq = w.kq; a.vq = w.vq; a.ks = w.ks; a.vs = w.vs; a.cnt = w.cnt;
    a.dq = (need_grad || proto_sim) ? w.dq : nullptr;
    a.patt = need_grad ? w.patt : nullptr;
    a.rowred = w.rowred;
    a.rowdot = need_grad ? w.rowdot : nullptr;
    a.linv = need_grad ? w.linv : nullptr;
    if (int rc = trx_attn_fwd(a, s, st)) return rc;
    if (proto_sim)
      if (int rc = trx_proto_sim_fwd(w.vq, w.dq, w.cnt, w.gram, proto_sim, s, st)) return rc;
    return trx_logits_fwd(w.rowred, w.cnt, logits, s, st);
  }
  LMKD_CUDA(cudaMemsetAsync(w.rowred, 0, sizeof(float) * s.B * s.way * s.NqT, st));
  for (int q0 = 0; q0 < s.Nq; q0 += w.qchunk) {
    const int nq = s.Nq - q0 < w.qchunk ? s.Nq - q0 : w.qchunk;
    if (int rc = trx_attn_chunk_fwd(w, s, q0, nq, (need_grad || proto_sim) ? w.dq : nullptr, st)) return rc;
  }
  if (proto_sim)
    if (int rc = trx_proto_sim_fwd(w.vq, w.dq, w.cnt, w.gram, proto_sim, s, st)) return rc;
  return trx_logits_fwd(w.rowred, w.cnt, logits, s, st);
}

int lmkd_trx_bwd(const lmkd_trx_shape* sh, const float* grad_logits, const float* grad_proto_sim,
                 const int32_t* tuples, const int32_t* inv_off,
                 const int32_t* inv_idx, const float* bk, const float* gamma, const float* beta, float* grad_support,
                 float* grad_query,
                 float* gWk, float* gbk, float* gWv, float* gbv, float* ggamma, float* gbeta, void* workspace,
                 int need_grad, int accumulate, void* stream) {
  const int accumulate_param_grads = accumulate & 1;      // bit 0: the six parameter gradients are added to
  const int accumulate_feature_grads = (accumulate >> 1) & 1;   // bit 1: grad_support / grad_query are added to
  TrxDims s;
  if (int rc = trx_dims(sh, &s)) return rc;
  LMKD_CHECK(need_grad == 1 || need_grad == 2, "trx_bwd: need_grad must be the value (1 or 2) the forward ran with");
  LMKD_CHECK(grad_proto_sim == nullptr || need_grad == 2, "trx_bwd: grad_proto_sim needs a forward run with need_grad = 2");
  LMKD_CHECK(grad_logits && tuples && inv_off && inv_idx && bk && gamma && beta && grad_support && grad_query && gWk && gbk &&
                 gWv && gbv && ggamma && gbeta && workspace,
             "trx_bwd: null pointer");
  cudaStream_t st = S(stream);
  TrxWs w = trx_layout(workspace, s, need_grad);
  const int64_t pcols = 2ll * s.card * s.d;
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  const float inv_sqrt_d = 1.f / sqrtf(static_cast<float>(s.d));

  if (int rc = trx_attn_bwd_prep(grad_logits, w.cnt, w.srow, w.fused ? w.linv : nullptr, w.fused ? w.rs : nullptr, s, st))
    return rc;
  // Gradient w.r.t. the class prototypes: srow_c * diff_c when only the logits carry gradient (the row
  // scale then rides in the GEMM epilogue / in Ps); a materialised tensor E when TRX_sup's prototype
  // similarities do too.
  const __nv_bfloat16* protograd = w.dq;
  if (grad_proto_sim) {
    if (int rc = trx_proto_sim_bwd(w.vq, w.dq, w.cnt, w.gram, grad_proto_sim, w.srow, w.E, s, st)) return rc;
    protograd = w.E;
  }
  const bool fused = w.ln_fused;
  if (fused) {
    LMKD_CUDA(cudaMemsetAsync(w.lnred_q, 0, sizeof(float) * 2 * s.B * s.NqT, st));
    LMKD_CUDA(cudaMemsetAsync(w.lnred_s, 0, sizeof(float) * 2 * s.B * pitch, st));
  }
  // One pass over all queries (fused attention, or a materialised path that fits its byte budget), otherwise
  // `qchunk` queries at a time: the chunk's probabilities are recomputed, dV_s / dK_s accumulate over the passes.
  const int qstep = w.fused ? s.Nq : w.qchunk;
  for (int q0 = 0; q0 < s.Nq; q0 += qstep) {
    const int nq = s.Nq - q0 < qstep ? s.Nq - q0 : qstep;
    const TrxDims c = trx_chunk_dims(s, q0, nq);
    const bool first = q0 == 0;
    const int64_t moff_d = static_cast<int64_t>(c.m_off) * s.d;
    const int64_t cstride = static_cast<int64_t>(c.NqT) * pitch;         // batch stride of the chunk-local buffers
    if (w.fused) {
      // dP = <diff_c[m], v_s[(c, kt)]> with the softmax backward in the epilogue: with p = P~ * srow / rowsum,
      // Ps = p and dS = p * (dP - <diff, prototype>) leave as bf16; neither dP nor the probabilities are re-read
      GemmDesc g;
      g.M = s.NqT; g.N = s.KTp; g.K = s.d; g.nb1 = s.way; g.nb2 = s.B;
      g.A.ptr = w.dq; g.A.ld = s.d; g.A.stride_b1 = static_cast<int64_t>(s.NqT) * s.d;
      g.A.stride_b2 = static_cast<int64_t>(s.way) * s.NqT * s.d;
      g.B.ptr = w.vs; g.B.ld = s.d; g.B.stride_b1 = static_cast<int64_t>(s.KTp) * s.d; g.B.stride_b2 = pitch * s.d;
      g.epi.kind = EPI_SMBWD_BF16;
      g.epi.C = w.dS; g.epi.C2 = w.ps; g.epi.ldc = pitch; g.epi.c_b1 = s.KTp;
      g.epi.c_b2 = static_cast<int64_t>(s.NqT) * pitch;
      g.epi.aux = w.patt; g.epi.ldaux = pitch; g.epi.aux_b1 = s.KTp; g.epi.aux_b2 = static_cast<int64_t>(s.NqT) * pitch;
      g.epi.rowv = w.rs; g.epi.rowv2 = w.rowdot; g.epi.rv_b1 = s.NqT; g.epi.rv_b2 = static_cast<int64_t>(s.way) * s.NqT;
      if (int rc = gemm_bf16(g, st)) return rc;
    } else {
      if (w.qchunk < s.Nq)
        if (int rc = trx_scores_softmax(w, s, c, st)) return rc;          // the forward kept no probabilities
      {  // dP[b][m][(c, kt)] = <dO_c[m], v_s[(c, kt)]>
        GemmDesc g;
        g.M = c.NqT; g.N = s.KTp; g.K = s.d; g.nb1 = s.way; g.nb2 = s.B;
        g.A.ptr = protograd + moff_d; g.A.ld = s.d; g.A.stride_b1 = static_cast<int64_t>(s.NqT) * s.d;
        g.A.stride_b2 = static_cast<int64_t>(s.way) * s.NqT * s.d;
        g.B.ptr = w.vs; g.B.ld = s.d; g.B.stride_b1 = static_cast<int64_t>(s.KTp) * s.d; g.B.stride_b2 = pitch * s.d;
        g.epi.kind = EPI_STORE_F32;
        g.epi.C = w.dP; g.epi.ldc = pitch; g.epi.c_b1 = s.KTp; g.epi.c_b2 = cstride;
        if (!grad_proto_sim) {
          g.epi.rowv = w.srow + c.m_off; g.epi.rv_b1 = s.NqT; g.epi.rv_b2 = static_cast<int64_t>(s.way) * s.NqT;
        }
        if (int rc = gemm_bf16(g, st)) return rc;
      }
      if (int rc = trx_softmax_bwd(w.patt, w.dP, w.cnt, w.srow, w.dS, w.ps, c, st)) return rc;
    }
    if (first && nq == s.Nq && g_dv_transposed) {
      // dV_s^T[:][(c, kt)] = sum_m dO_c[m][:] * P[m][(c, kt)]: the long dimension d is the tile-row dimension (9 full
      // 128-row tiles at d = 1152) and KTp the column dimension (any multiple of 16), instead of KTp = 288 / 144 rows
      // padded to 384 / 256; the epilogue stores the tile transposed, lanes writing consecutive d
      GemmDesc g;
      g.M = s.d; g.N = s.KTp; g.K = c.NqT; g.nb1 = s.way; g.nb2 = s.B;
      g.A.ptr = protograd + moff_d; g.A.mn_major = 1; g.A.ld = s.d; g.A.stride_b1 = static_cast<int64_t>(s.NqT) * s.d;
      g.A.stride_b2 = static_cast<int64_t>(s.way) * s.NqT * s.d;
      g.B.ptr = grad_proto_sim ? w.patt : w.ps; g.B.mn_major = 1; g.B.ld = pitch; g.B.stride_b1 = s.KTp;
      g.B.stride_b2 = cstride;
      g.epi.kind = w.g16 ? EPI_STORE_BF16 : EPI_STORE_F32; g.epi.c_transposed = 1;
      g.epi.C = w.dVs; g.epi.ldc = s.d; g.epi.c_b1 = static_cast<int64_t>(s.KTp) * s.d; g.epi.c_b2 = pitch * s.d;
      if (int rc = gemm_bf16(g, st)) return rc;
    } else {  // dV_s[(c, kt)][:] (+)= sum_m P[m][(c, kt)] * dO_c[m][:]
      GemmDesc g;
      g.M = s.KTp; g.N = s.d; g.K = c.NqT; g.nb1 = s.way; g.nb2 = s.B;
      g.A.ptr = grad_proto_sim ? w.patt : w.ps; g.A.mn_major = 1; g.A.ld = pitch; g.A.stride_b1 = s.KTp;
      g.A.stride_b2 = cstride;
      g.B.ptr = protograd + moff_d; g.B.mn_major = 1; g.B.ld = s.d; g.B.stride_b1 = static_cast<int64_t>(s.NqT) * s.d;
      g.B.stride_b2 = static_cast<int64_t>(s.way) * s.NqT * s.d;
      g.epi.kind = w.g16 ? EPI_STORE_BF16 : (first ? EPI_STORE_F32 : EPI_ACCUM_F32);
      g.epi.C = w.dVs; g.epi.ldc = s.d; g.epi.c_b1 = static_cast<int64_t>(s.KTp) * s.d; g.epi.c_b2 = pitch * s.d;
      if (int rc = gemm_bf16(g, st)) return rc;
    }
    {  // dK_q = dS . K_s / sqrt(d)   (+ the LayerNorm-backward row reductions in the epilogue)
      GemmDesc g;
      g.M = c.NqT; g.N = s.d; g.K = static_cast<int>(pitch); g.nb2 = s.B;
      g.A.ptr = w.dS; g.A.ld = pitch; g.A.stride_b2 = cstride;
      g.B.ptr = w.ks; g.B.mn_major = 1; g.B.ld = s.d; g.B.stride_b2 = pitch * s.d;
      g.epi.kind = EPI_STORE_F32; g.epi.alpha = inv_sqrt_d;
      g.epi.C = w.dKq + moff_d; g.epi.ldc = s.d; g.epi.c_b2 = static_cast<int64_t>(s.NqT) * s.d;
      if (fused) {
        g.epi.kind = w.g16 ? EPI_LNRED_BF16 : EPI_LNRED_F32;
        g.epi.aux = w.kq; g.epi.ldaux = s.d; g.epi.aux_b2 = static_cast<int64_t>(s.NqT) * s.d;
        g.epi.colv = gamma; g.epi.colv2 = beta;
        g.epi.rowred = w.lnred_q; g.epi.rr_b2 = s.NqT;
      }
      if (int rc = gemm_bf16(g, st)) return rc;
    }
    {  // dK_s (+)= dS^T . K_q / sqrt(d)
      GemmDesc g;
      g.M = static_cast<int>(pitch); g.N = s.d; g.K = c.NqT; g.nb2 = s.B;
      g.A.ptr = w.dS; g.A.mn_major = 1; g.A.ld = pitch; g.A.stride_b2 = cstride;
      g.B.ptr = w.kq + moff_d; g.B.mn_major = 1; g.B.ld = s.d; g.B.stride_b2 = static_cast<int64_t>(s.NqT) * s.d;
      g.epi.kind = first ? EPI_STORE_F32 : EPI_ACCUM_F32; g.epi.alpha = inv_sqrt_d;
      g.epi.C = w.dKs; g.epi.ldc = s.d; g.epi.c_b2 = pitch * s.d;
      if (fused) {
        g.epi.kind = w.g16 ? EPI_LNRED_BF16 : EPI_LNRED_F32;
        g.epi.aux = w.ks; g.epi.ldaux = s.d; g.epi.aux_b2 = pitch * s.d;
        g.epi.colv = gamma; g.epi.colv2 = beta;
        g.epi.rowred = w.lnred_s; g.epi.rr_b2 = pitch;
      }
      if (int rc = gemm_bf16(g, st)) return rc;
    }
  }
  int nblocks = 0;
  if (fused) {
    if (int rc = trx_ln_gather_bwd_fused(w.P, bk, gamma, w.stats, tuples, w.slot, w.dKq, w.dKs, w.dVs, w.g16 ? 1 : 0, w.lnred_q,
                                         w.lnred_s, w.srow, w.dq, w.dpcat, w.partials, w.max_partial_blocks, &nblocks,
                                         s, st))
      return rc;
    if (int rc = trx_reduce_partials(w.partials, nblocks, ggamma, gbeta, gbk, gbv, s.d, accumulate_param_grads, st)) return rc;
  } else {   // long clips: accumulators do not fit in shared memory -> materialise dx rows, then gather
    if (int rc = trx_ln_bwd(w.P, bk, gamma, w.stats, tuples, w.slot, w.dKq, w.dKs, w.dVs, w.srow, w.dq, w.dxk, w.dxv,
                            w.partials, w.max_partial_blocks, &nblocks, s, st))
      return rc;
    if (int rc = trx_reduce_partials(w.partials, nblocks, ggamma, gbeta, gbk, gbv, s.d, accumulate_param_grads, st)) return rc;
    if (int rc = trx_tuple_gather_bwd(w.dxk, w.dxv, inv_off, inv_idx, w.dpcat, s, st)) return rc;
  }
  {  // dX~[M, D] = dPcat[M, 2cd] . Wcat[2cd, D]
    GemmDesc g;
    g.M = static_cast<int>(s.M); g.N = s.D; g.K = static_cast<int>(pcols);
    g.A.ptr = w.dpcat; g.A.ld = pcols;
    g.B.ptr = w.wcat; g.B.mn_major = 1; g.B.ld = s.D;
    if (g_dx_fused) {
      // the dropout mask and the split into support / query gradients happen in the epilogue: no dX round trip
      g.epi.kind = EPI_DXSCATTER; g.epi.C = grad_support; g.epi.C2 = grad_query; g.epi.ldc = s.D;
      g.epi.group_rows = s.N * s.L; g.epi.split_rows = s.Ns * s.L;
      g.epi.drop_p = sh->dropout_p; g.epi.seed = reinterpret_cast<const unsigned long long*>(w.seed_used);
      g.epi.accumulate = accumulate_feature_grads;
    } else {
      g.epi.kind = EPI_STORE_F32; g.epi.C = w.dX; g.epi.ldc = s.D;
    }
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  {  // dWcat[2cd, D] = dPcat^T . X~
    GemmDesc g;
    g.M = static_cast<int>(pcols); g.N = s.D; g.K = static_cast<int>(s.M);
    g.A.ptr = w.dpcat; g.A.mn_major = 1; g.A.ld = pcols;
    g.B.ptr = w.xb; g.B.mn_major = 1; g.B.ld = s.D;
    g.epi.kind = EPI_STORE_F32; g.epi.C = w.dWcat; g.epi.ldc = s.D;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  if (int rc = trx_unpack_wgrad(w.dWcat, gWk, gWv, s, accumulate_param_grads, st)) return rc;
  if (g_dx_fused) return 0;
  return trx_dx_scatter(w.dX, grad_support, grad_query, s.B, s.Ns, s.Nq, s.L, s.D, sh->dropout_p, w.seed_used,
                        accumulate_feature_grads, st);
}

void lmkd_trx_set_attn_budget(double bytes) { g_attn_budget_override = bytes > 0.0 ? bytes : 0.0; }

int lmkd_trx_attn_fused_fits(const lmkd_trx_shape* sh) {
  TrxDims s;
  if (trx_dims(sh, &s)) return 0;
  return trx_attn_fused_fits(s) ? 1 : 0;
}

int lmkd_trx_attn_fwd(const lmkd_trx_shape* sh, const void* kq, const void* vq, const void* ks, const void* vs,
                      const int32_t* cnt, void* dq, void* patt, float* rowred, float* rowdot, float* linv,
                      void* stream) {
  TrxDims s;
  if (int rc = trx_dims(sh, &s)) return rc;
  LMKD_CHECK(kq && vq && ks && vs && cnt && rowred, "trx_attn_fwd: null pointer");
  TrxAttnFwd a{};
  a.kq = static_cast<const __nv_bfloat16*>(kq); a.vq = static_cast<const __nv_bfloat16*>(vq);
  a.ks = static_cast<const __nv_bfloat16*>(ks); a.vs = static_cast<const __nv_bfloat16*>(vs);
  a.cnt = cnt;
  a.dq = static_cast<__nv_bfloat16*>(dq); a.patt = static_cast<__nv_bfloat16*>(patt);
  a.rowred = rowred; a.rowdot = rowdot; a.linv = linv;
  return trx_attn_fwd(a, s, S(stream));
}

// ---------------------------------------------------------------------------------------------
size_t lmkd_strm_dist_workspace_bytes(const lmkd_trx_shape* s, int need_grad) {
  TrxDims d;
  if (trx_dims(s, &d)) return 0;
  return strm_layout(nullptr, d, need_grad).bytes;
}

int lmkd_strm_dist_fwd(const lmkd_trx_shape* sh, const float* support, const float* labels, const float* query,
                       const int32_t* tuples, const float* W, const float* bias, float* logits, void* workspace,
                       int need_grad, int* status, void* stream) {
  TrxDims s;
  if (int rc = trx_dims(sh, &s)) return rc;
  LMKD_CHECK(support && labels && query && tuples && W && bias && logits && workspace, "strm_dist_fwd: null pointer");
  cudaStream_t st = S(stream);
  StrmWs w = strm_layout(workspace, s, need_grad);
  const int64_t pcols = static_cast<int64_t>(s.card) * s.d;
  if (int rc = trx_class_slots(labels, w.slot, w.cnt, status, s, st)) return rc;
  // DistanceLoss applies dropout but NO positional encoding (strm_res18_sup.py:190-192): the cast kernel adds a zero table
  LMKD_CUDA(cudaMemsetAsync(w.pe0, 0, sizeof(float) * s.L * s.D, st));
  if (int rc = trx_pe_cast(support, query, w.pe0, w.xb, s.B, s.Ns, s.Nq, s.L, s.D, sh->dropout_p, sh->seed, sh->seed_dev,
                           w.seed_used, st)) return rc;
  if (int rc = strm_pack_weight(W, w.wcat, s, st)) return rc;
  {  // per-frame partial projections of the tuple MLP: P[M, c*dm] = X~[M, D] . Wcat[c*dm, D]^T   (:204, :221)
    GemmDesc g;
    g.M = static_cast<int>(s.M); g.N = static_cast<int>(pcols); g.K = s.D;
    g.A.ptr = w.xb; g.A.ld = s.D;
    g.B.ptr = w.wcat; g.B.ld = s.D;
    g.epi.kind = EPI_STORE_F32; g.epi.C = w.P; g.epi.ldc = pcols;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  if (int rc = strm_tuple_relu_fwd(w.P, bias, tuples, w.slot, w.cnt, w.eq, w.es, w.nq2, w.ns2, s, st)) return rc;
  LMKD_CUDA(cudaMemsetAsync(w.best, 0xFF, sizeof(unsigned long long) * s.B * s.way * s.NqT, st));
  {  // per class: |e_q - e_s|^2 = |e_q|^2 + |e_s|^2 - 2 <e_q, e_s>, arg-min over the class's support tuples (:227-230)
    GemmDesc g;
    g.M = s.NqT; g.N = s.KTp; g.K = s.d; g.nb1 = s.way; g.nb2 = s.B;
    g.A.ptr = w.eq; g.A.ld = s.d; g.A.stride_b1 = 0; g.A.stride_b2 = static_cast<int64_t>(s.NqT) * s.d;
    g.B.ptr = w.es; g.B.ld = s.d; g.B.stride_b1 = static_cast<int64_t>(s.KTp) * s.d;
    g.B.stride_b2 = static_cast<int64_t>(s.way) * s.KTp * s.d;
    g.epi.kind = EPI_MINDIST;
    g.epi.rowv = w.nq2; g.epi.rv_b1 = 0; g.epi.rv_b2 = s.NqT;
    g.epi.colv = w.ns2; g.epi.cv_b1 = s.KTp; g.epi.cv_b2 = static_cast<int64_t>(s.way) * s.KTp;
    g.epi.rowred = reinterpret_cast<float*>(w.best); g.epi.rr_b1 = s.NqT; g.epi.rr_b2 = static_cast<int64_t>(s.way) * s.NqT;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  return strm_logits_fwd(w.best, w.cnt, logits, s, st);
}

int lmkd_strm_dist_bwd(const lmkd_trx_shape* sh, const float* grad_logits, const int32_t* inv_off, const int32_t* inv_idx,
                       float* grad_support, float* grad_query, float* gW, float* gbias, void* workspace, void* stream) {
  TrxDims s;
  if (int rc = trx_dims(sh, &s)) return rc;
  LMKD_CHECK(grad_logits && inv_off && inv_idx && grad_support && grad_query && gW && gbias && workspace,
             "strm_dist_bwd: null pointer");
  cudaStream_t st = S(stream);
  StrmWs w = strm_layout(workspace, s, 1);
  const int64_t pcols = static_cast<int64_t>(s.card) * s.d;
  LMKD_CUDA(cudaMemsetAsync(w.dEs, 0, sizeof(float) * s.B * s.way * s.KTp * s.d, st));
  LMKD_CUDA(cudaMemsetAsync(gbias, 0, sizeof(float) * s.d, st));
  if (int rc = strm_dist_bwd(grad_logits, w.best, w.cnt, w.eq, w.es, w.dEq, w.dEs, s, st)) return rc;
  if (int rc = strm_relu_gather_bwd(w.dEq, w.dEs, w.eq, w.es, w.slot, inv_off, inv_idx, w.dpcat, gbias, s, st)) return rc;
  {  // dX~[M, D] = dPcat[M, c*dm] . Wcat[c*dm, D]
    GemmDesc g;
    g.M = static_cast<int>(s.M); g.N = s.D; g.K = static_cast<int>(pcols);
    g.A.ptr = w.dpcat; g.A.ld = pcols;
    g.B.ptr = w.wcat; g.B.mn_major = 1; g.B.ld = s.D;
    g.epi.kind = EPI_STORE_F32; g.epi.C = w.dX; g.epi.ldc = s.D;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  {  // dWcat[c*dm, D] = dPcat^T . X~
    GemmDesc g;
    g.M = static_cast<int>(pcols); g.N = s.D; g.K = static_cast<int>(s.M);
    g.A.ptr = w.dpcat; g.A.mn_major = 1; g.A.ld = pcols;
    g.B.ptr = w.xb; g.B.mn_major = 1; g.B.ld = s.D;
    g.epi.kind = EPI_STORE_F32; g.epi.C = w.dWcat; g.epi.ldc = s.D;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  if (int rc = strm_unpack_wgrad(w.dWcat, gW, s, st)) return rc;
  return trx_dx_scatter(w.dX, grad_support, grad_query, s.B, s.Ns, s.Nq, s.L, s.D, sh->dropout_p, w.seed_used, 0, st);
}

int lmkd_dropout_mask(float* out, int64_t n, float p, uint64_t seed, void* stream) {
  LMKD_CHECK(out && n > 0, "dropout_mask: bad arguments");
  return dropout_mask(out, n, p, seed, S(stream));
}

// ---------------------------------------------------------------------------------------------
int lmkd_support_dk_fwd(const float* support, int B, int way, int shot, int L, int D, float* protos, float* out,
                        void* stream) {
  LMKD_CHECK(support && protos && out, "support_dk_fwd: null pointer");
  return support_dk_fwd(support, protos, out, B, way, shot, L, D, S(stream));
}

int lmkd_support_dk_bwd(const float* grad_out, const float* protos, int B, int way, int shot, int L, int D,
                        float* grad_support, void* stream) {
  LMKD_CHECK(grad_out && protos && grad_support, "support_dk_bwd: null pointer");
  return support_dk_bwd(grad_out, protos, grad_support, B, way, shot, L, D, S(stream));
}

// ---------------------------------------------------------------------------------------------
int lmkd_d2m_logit_loss(const lmkd_loss_term* terms, int nterms, float temperature, const float* fnum,
                        const float* fden, const int64_t* fy, int frows, int fcols, int B, float* loss, float* values,
                        float* focal, void* stream) {
  LMKD_CHECK(terms && loss && B > 0, "d2m_logit_loss: bad arguments");
  LMKD_CHECK(nterms >= 1 && nterms <= kMaxTerms, "d2m_logit_loss: %d terms (max %d)", nterms, kMaxTerms);
  LMKD_CHECK(temperature > 0.f, "d2m_logit_loss: temperature must be positive");
  LossSpec spec;
  memset(&spec, 0, sizeof(spec));
  spec.nterms = nterms;
  for (int i = 0; i < nterms; ++i) {
    LossTerm& t = spec.terms[i];
    t.kind = terms[i].kind; t.rows = terms[i].rows; t.cols = terms[i].cols;
    t.s = terms[i].s; t.t = terms[i].t; t.y = terms[i].y; t.grad = terms[i].grad;
    t.grad_accumulate = terms[i].grad_accumulate;
    t.w = terms[i].w; t.fa = terms[i].fa; t.fb = terms[i].fb;
    LMKD_CHECK(t.kind >= 0 && t.kind <= 2, "d2m_logit_loss: term %d has unknown kind %d", i, t.kind);
  }
  spec.temperature = temperature;
  spec.fnum = fnum; spec.fden = fden; spec.fy = fy; spec.frows = frows; spec.fcols = fcols;
  return d2m_logit_loss(spec, B, loss, values, focal, S(stream));
}

int lmkd_mse_partials(void) { return 148 * 8 * 2; }

int lmkd_d2m_feature_mse_fwdbwd(const void* s, const void* t, void* ds, int64_t n, int dtype, float lscale,
                                float gscale, float* partials, float* loss, int accumulate, void* stream) {
  LMKD_CHECK(s && t && ds && partials && loss, "feature_mse: null pointer");
  int np = 0;
  int rc;
  if (dtype == 0)
    rc = feat_mse_fwdbwd(static_cast<const float*>(s), static_cast<const float*>(t), static_cast<float*>(ds), n, gscale,
                         partials, lmkd_mse_partials(), &np, S(stream));
  else if (dtype == 1)
    rc = feat_mse_fwdbwd_bf16(static_cast<const __nv_bfloat16*>(s), static_cast<const __nv_bfloat16*>(t),
                              static_cast<__nv_bfloat16*>(ds), n, gscale, partials, lmkd_mse_partials(), &np, S(stream));
  else {
    set_error("feature_mse: unknown dtype %d", dtype);
    return 1;
  }
  if (rc) return rc;
  return mse_finish(partials, np, lscale, loss, accumulate, S(stream));
}

int lmkd_episode_gather(const void* store, int store_dtype, int64_t store_rows, const int64_t* index, int64_t count,
                        int64_t row_elems, float* out, int* status, void* stream) {
  LMKD_CHECK(store && index && out, "episode_gather: null pointer");
  LMKD_CHECK(store_dtype == 0 || store_dtype == 1, "episode_gather: unknown store dtype %d", store_dtype);
  return episode_gather(store, store_dtype, store_rows, index, count, row_elems, out, status, S(stream));
}

int lmkd_d2m_feature_mse_store_fwdbwd(const float* s, const void* store, int store_dtype, int64_t store_rows,
                                      const int64_t* index, int64_t count, int64_t row_elems, float* ds, float lscale,
                                      float gscale, float* partials, float* loss, int accumulate, int* status,
                                      void* stream) {
  LMKD_CHECK(s && store && index && ds && partials && loss, "feature_mse_store: null pointer");
  LMKD_CHECK(store_dtype == 0 || store_dtype == 1, "feature_mse_store: unknown store dtype %d", store_dtype);
  int np = 0;
  if (int rc = feat_mse_store_fwdbwd(s, store, store_dtype, store_rows, index, count, row_elems, ds, gscale, partials,
                                     lmkd_mse_partials(), &np, status, S(stream)))
    return rc;
  return mse_finish(partials, np, lscale, loss, accumulate, S(stream));
}

// ---------------------------------------------------------------------------------------------
// teacher multi-modal fusion forward (fusion.cu); the C structs mirror FusionLayer / FusionEncoder field by field
static_assert(sizeof(lmkd_fusion_layer) == sizeof(FusionLayer), "lmkd_fusion_layer layout");
static_assert(sizeof(lmkd_fusion_encoder) == sizeof(FusionEncoder), "lmkd_fusion_encoder layout");

size_t lmkd_fusion_workspace_bytes(const lmkd_fusion_encoder* enc, int64_t nvideos, int L) {
  if (enc == nullptr) {
    set_error("fusion: null encoder");
    return 0;
  }
  const FusionEncoder& e = *reinterpret_cast<const FusionEncoder*>(enc);
  if (fusion_check(e, nvideos, L)) return 0;
  return fusion_layout(nullptr, e, nvideos, L).bytes;
}

int lmkd_fusion_fwd(const lmkd_fusion_encoder* enc, const float* const* x, const int32_t* shift, int64_t nvideos, int L,
                    float* out, int accumulate, void* workspace, void* stream) {
  LMKD_CHECK(enc != nullptr, "fusion_fwd: null encoder");
  return fusion_forward(*reinterpret_cast<const FusionEncoder*>(enc), x, shift, nvideos, L, out, accumulate, workspace,
                        S(stream));
}

int lmkd_scale_by_device_scalar(float* x, int64_t n, const float* g, void* stream) {
  LMKD_CHECK(x && g && n >= 0, "scale: bad arguments");
  if (n == 0) return 0;
  return scale_by_device_scalar(x, n, g, S(stream));
}

int lmkd_accuracy_count(const float* logits, const int64_t* labels, int64_t rows, int cols, int* correct,
                        void* stream) {
  LMKD_CHECK(logits && labels && correct && rows > 0 && cols > 0, "accuracy: bad arguments");
  return accuracy_count(logits, labels, rows, cols, correct, S(stream));
}

// ---------------------------------------------------------------------------------------------
int lmkd_gemm_bf16(int M, int N, int K, int batch, const void* A, int a_mn, int64_t lda, int64_t a_bs, const void* B,
                   int b_mn, int64_t ldb, int64_t b_bs, float* C, int64_t ldc, int64_t c_bs, float alpha,
                   int accumulate, int block_n, void* stream) {
  LMKD_CHECK(A && B && C, "gemm: null pointer");
  GemmDesc g;
  g.M = M; g.N = N; g.K = K; g.nb1 = 1; g.nb2 = batch;
  g.A.ptr = static_cast<const __nv_bfloat16*>(A); g.A.mn_major = a_mn; g.A.ld = lda; g.A.stride_b2 = a_bs;
  g.B.ptr = static_cast<const __nv_bfloat16*>(B); g.B.mn_major = b_mn; g.B.ld = ldb; g.B.stride_b2 = b_bs;
  g.block_n = block_n;
  g.epi.kind = accumulate ? EPI_ACCUM_F32 : EPI_STORE_F32;
  g.epi.alpha = alpha;
  g.epi.C = C; g.epi.ldc = ldc; g.epi.c_b2 = c_bs;
  return gemm_bf16(g, S(stream));
}

// ---------------------------------------------------------------------------------------------
int lmkd_frame_pool_fwd(const float* fmap, int64_t rows, int C, int H, int W, int out_hw, float* pooled,
                        void* stream) {
  LMKD_CHECK(fmap && pooled, "frame_pool_fwd: null pointer");
  LMKD_CHECK(rows > 0 && C > 0, "frame_pool_fwd: empty input");
  return frame_pool_fwd(fmap, pooled, rows, C, H, W, out_hw, S(stream));
}

int lmkd_frame_pool_bwd(const float* fmap, const float* grad_pooled, int64_t rows, int C, int H, int W, int out_hw,
                        float* grad_fmap, void* stream) {
  LMKD_CHECK(fmap && grad_pooled && grad_fmap, "frame_pool_bwd: null pointer");
  LMKD_CHECK(rows > 0 && C > 0, "frame_pool_bwd: empty input");
  return frame_pool_bwd(fmap, grad_pooled, grad_fmap, rows, C, H, W, out_hw, S(stream));
}

namespace {
struct FeatureHeadWs {
  __nv_bfloat16 *xb, *wb, *dyb;
  size_t bytes;
};
FeatureHeadWs feature_head_layout(void* ws, int64_t rows, int in_dim, int out_dim, int heads) {
  Carver c(ws);
  FeatureHeadWs w;
  w.xb = c.take<__nv_bfloat16>(rows * in_dim);
  w.wb = c.take<__nv_bfloat16>(static_cast<int64_t>(heads) * out_dim * in_dim);
  w.dyb = c.take<__nv_bfloat16>(static_cast<int64_t>(heads) * rows * out_dim);
  w.bytes = c.total();
  return w;
}
int feature_head_check(int64_t rows, int in_dim, int out_dim, int heads) {
  LMKD_CHECK(rows > 0 && rows < (1ll << 31) && heads > 0, "feature_head: bad row / head count");
  LMKD_CHECK(in_dim > 0 && out_dim > 0 && in_dim % 8 == 0 && out_dim % 8 == 0,
             "feature_head: dims %d -> %d must be positive multiples of 8", in_dim, out_dim);
  return 0;
}
}  // namespace

size_t lmkd_feature_head_workspace_bytes(int64_t rows, int in_dim, int out_dim, int heads) {
  return feature_head_layout(nullptr, rows, in_dim, out_dim, heads).bytes;
}

int lmkd_feature_head_fwd(const float* x, const float* weight, const float* bias, int64_t rows, int in_dim,
                          int out_dim, int heads, float* y, void* workspace, void* stream) {
  LMKD_CHECK(x && weight && bias && y && workspace, "feature_head_fwd: null pointer");
  if (int rc = feature_head_check(rows, in_dim, out_dim, heads)) return rc;
  cudaStream_t st = S(stream);
  FeatureHeadWs w = feature_head_layout(workspace, rows, in_dim, out_dim, heads);
  if (int rc = cast_bf16(x, w.xb, rows * in_dim, st)) return rc;
  if (int rc = cast_bf16(weight, w.wb, static_cast<int64_t>(heads) * out_dim * in_dim, st)) return rc;
  for (int h = 0; h < heads; ++h) {   // y[h] = x . W[h]^T + b[h]
    GemmDesc g;
    g.M = static_cast<int>(rows); g.N = out_dim; g.K = in_dim;
    g.A.ptr = w.xb; g.A.ld = in_dim;
    g.B.ptr = w.wb + static_cast<int64_t>(h) * out_dim * in_dim; g.B.ld = in_dim;
    g.epi.kind = EPI_BIAS_F32; g.epi.alpha = 1.f;
    g.epi.C = y + static_cast<int64_t>(h) * rows * out_dim; g.epi.ldc = out_dim;
    g.epi.colv = bias + static_cast<int64_t>(h) * out_dim;
    if (int rc = gemm_bf16(g, st)) return rc;
  }
  return 0;
}

int lmkd_feature_head_bwd(const float* grad_y, int64_t rows, int in_dim, int out_dim, int heads, float* grad_x,
                          float* grad_weight, float* grad_bias, void* workspace, void* stream) {
  LMKD_CHECK(grad_y && workspace, "feature_head_bwd: null pointer");
  if (int rc = feature_head_check(rows, in_dim, out_dim, heads)) return rc;
  cudaStream_t st = S(stream);
  FeatureHeadWs w = feature_head_layout(workspace, rows, in_dim, out_dim, heads);
  const int64_t ysz = rows * out_dim;
  if (grad_bias != nullptr) {
    LMKD_CUDA(cudaMemsetAsync(grad_bias, 0, sizeof(float) * heads * out_dim, st));
    for (int h = 0; h < heads; ++h)
      if (int rc = cast_colsum(grad_y + h * ysz, w.dyb + h * ysz, grad_bias + static_cast<int64_t>(h) * out_dim, rows,
                               out_dim, st))
        return rc;
  } else {
    if (int rc = cast_bf16(grad_y, w.dyb, heads * ysz, st)) return rc;
  }
  for (int h = 0; h < heads; ++h) {
    const __nv_bfloat16* dy = w.dyb + h * ysz;
    if (grad_x != nullptr) {   // dX (+)= dY[h] . W[h]      (W[h] is [out, in]: the contraction index is its row)
      GemmDesc g;
      g.M = static_cast<int>(rows); g.N = in_dim; g.K = out_dim;
      g.A.ptr = dy; g.A.ld = out_dim;
      g.B.ptr = w.wb + static_cast<int64_t>(h) * out_dim * in_dim; g.B.mn_major = 1; g.B.ld = in_dim;
      g.epi.kind = h == 0 ? EPI_STORE_F32 : EPI_ACCUM_F32; g.epi.alpha = 1.f;
      g.epi.C = grad_x; g.epi.ldc = in_dim;
      if (int rc = gemm_bf16(g, st)) return rc;
    }
    if (grad_weight != nullptr) {   // dW[h] = dY[h]^T . X
      GemmDesc g;
      g.M = out_dim; g.N = in_dim; g.K = static_cast<int>(rows);
      g.A.ptr = dy; g.A.mn_major = 1; g.A.ld = out_dim;
      g.B.ptr = w.xb; g.B.mn_major = 1; g.B.ld = in_dim;
      g.epi.kind = EPI_STORE_F32; g.epi.alpha = 1.f;
      g.epi.C = grad_weight + static_cast<int64_t>(h) * out_dim * in_dim; g.epi.ldc = in_dim;
      if (int rc = gemm_bf16(g, st)) return rc;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
size_t lmkd_edist_workspace_bytes(int B, int Ns, int Nq, int D) {
  Carver c(nullptr);
  c.take<float>(static_cast<int64_t>(B) * Ns * D);
  c.take<float>(static_cast<int64_t>(B) * Nq * D);
  c.take<float>(static_cast<int64_t>(B) * Nq * Ns);
  return c.total();
}

int lmkd_edist_fwd(const float* support, const float* labels, const float* query, int B, int Ns, int Nq, int L, int D,
                   int way, float* logits, void* workspace, int* status, void* stream) {
  LMKD_CHECK(support && labels && query && logits && workspace, "edist_fwd: null pointer");
  Carver c(workspace);
  float* sm = c.take<float>(static_cast<int64_t>(B) * Ns * D);
  float* qm = c.take<float>(static_cast<int64_t>(B) * Nq * D);
  float* pd = c.take<float>(static_cast<int64_t>(B) * Nq * Ns);
  return edist_fwd(support, labels, query, sm, qm, pd, logits, B, Ns, Nq, L, D, way, status, S(stream));
}

int lmkd_edist_bwd(const float* grad_logits, const float* labels, int B, int Ns, int Nq, int L, int D, int way,
                   float* grad_support, float* grad_query, void* workspace, void* stream) {
  LMKD_CHECK(grad_logits && labels && grad_support && grad_query && workspace, "edist_bwd: null pointer");
  Carver c(workspace);
  float* sm = c.take<float>(static_cast<int64_t>(B) * Ns * D);
  float* qm = c.take<float>(static_cast<int64_t>(B) * Nq * D);
  float* pd = c.take<float>(static_cast<int64_t>(B) * Nq * Ns);
  return edist_bwd(grad_logits, labels, sm, qm, pd, grad_support, grad_query, B, Ns, Nq, L, D, way, S(stream));
}

long long lmkd_launch_count(int reset) { return launch_count(reset); }
void lmkd_gemm_timing_enable(int on) { gemm_timing_enable(on); }
int lmkd_gemm_timing_read(double* ms, double* flops, int* launches) {
  LMKD_CHECK(ms && flops && launches, "gemm_timing_read: null pointer");
  return gemm_timing_read(ms, flops, launches);
}

int lmkd_upcast_bf16(const void* x, float* y, int64_t n, void* stream) {
  LMKD_CHECK(x && y && n > 0, "upcast: bad arguments");
  return upcast_bf16(static_cast<const __nv_bfloat16*>(x), y, n, S(stream));
}

int lmkd_kernel_timing_read(int category, double* ms, double* work, int* launches) {
  LMKD_CHECK(ms && work && launches, "kernel_timing_read: null pointer");
  return kernel_timing_read(category, ms, work, launches);
}

int lmkd_cast_bf16(const float* x, void* y, int64_t n, void* stream) {
  LMKD_CHECK(x && y && n > 0, "cast: bad arguments");
  return cast_bf16(x, static_cast<__nv_bfloat16*>(y), n, S(stream));
}

}  // extern "C"
