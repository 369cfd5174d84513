// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0      : TMA producer (one lane) — cp.async.bulk.tensor 4-D loads into a ring of
//                 128B-swizzled smem stages, completion on `full` mbarriers
//   warp 1      : TMEM allocator + MMA issuer (one lane) — tcgen05.mma.cta_group::1.kind::f16,
//                 128 x BN x 16 per instruction, accumulators in TMEM (2 stages x 256 columns),
//                 tcgen05.commit releases smem stages / publishes finished accumulators
//   warps 2..5  : epilogue — tcgen05.ld of the accumulator (warp w owns TMEM lanes
//                 32*(w%4)..+31, i.e. one output row per thread), fused epilogue math, stores
//
// Tiles: BM = 128 rows, BN = runtime multiple of 16 (<= 256), BK = 64 bf16 (= one 128-byte
// swizzle row).  One CTA per SM, static round-robin over (batch, m-tile, n-tile).
// Ragged edges rely on TMA out-of-bounds zero fill (loads) and per-element masks (stores).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <vector>

#include "gemm.cuh"
#include "prep.cuh"
#include "tmap.cuh"

namespace lmkd {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
// Epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, splitting the tile's column units).
// Measured on B200 per product (profiles/r01_notes.md): the epilogues that read an aux tile (P.V DIFF_SQ, dK LNRED,
// OTAM AXPY) are the critical path of their short contractions and gain 13-14 % from 8 warps; the plain store
// epilogues lose 7-11 % (the extra staging slabs cost an operand stage), so they keep 4.
// The kernel is instantiated with EW = 4 and EW = 8 epilogue warps; threads = TMA warp + MMA warp + EW warps.
constexpr int kMaxStages = 8;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccStride = 256;

struct KParams {
  int M, N, K;
  int nb1;
  int tiles_m, tiles_n, num_tiles;
  int block_n, stages, num_kb;
  int a_mn, b_mn;
  int a_b1;                    // 0: operand A is shared by the inner batch (stride_b1 == 0), its map has no b1 extent
  int b_boxes;
  uint32_t idesc;
  uint32_t a_tile_bytes, b_tile_bytes, tx_bytes;
  int vec_ok;
  int bm;                      // rows per scheduled tile: 128, or 256 for a CTA pair
  int aux_tma;                 // DIFF_SQ / LNRED: aux tile arrives through TMA into shared memory
  int aux_boxes, aux_use_b1;
  int aux_box_cols;            // columns per 128-byte aux box row: 64 (bf16 aux) or 32 (fp32 aux)
  int tma_store;               // epilogue stores go through smem staging + TMA (coalesced)
  int own_staging;             // ... from a dedicated staging slab (otherwise in place, from the aux tile)
  int aux_bufs;                // aux tiles in shared memory: 2 (one per accumulator stage) or 1 (frees an operand stage)
  int stage_bufs;              // staging slabs per epilogue warp: 2 lets a unit be filled while the previous one drains
  int out_bf16;
  uint32_t aux_tile_bytes;
  GemmEpilogue epi;
};

struct TileCoord {
  int b1, b2, m0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int tile) {
  const int per_batch = p.tiles_m * p.tiles_n;
  const int batch = tile / per_batch;
  const int r = tile - batch * per_batch;
  TileCoord t;
  t.b1 = batch % p.nb1;
  t.b2 = batch / p.nb1;
  t.m0 = (r / p.tiles_n) * p.bm;
  t.n0 = (r % p.tiles_n) * p.block_n;
  return t;
}

__device__ __forceinline__ void store16_f32(float* dst, const float (&v)[16], int nvalid, bool vec) {
  if (vec && nvalid == 16) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}

__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const float (&v)[16], int nvalid,
                                             bool vec) {
  if (vec && nvalid == 16) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d4[1] = make_uint4(w[4], w[5], w[6], w[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) dst[i] = __float2bfloat16_rn(v[i]);
  }
}

// 16 auxiliary values of one output row chunk, fetched ahead of the accumulator they combine with
struct AuxRegs {
  float v[16];
};

// aux tile in smem: boxes of [128 rows][64 bf16] with the TMA 128-byte swizzle
__device__ __forceinline__ void load_aux_smem(const uint8_t* aux_tile, int row, int c, AuxRegs& a) {
  const uint8_t* base = aux_tile + (c >> 6) * (128 * 128) + row * 128;
  const int j = (c & 63) >> 3;   // first of the two 16-byte chunks
  const uint4 q0 = *reinterpret_cast<const uint4*>(base + ((j ^ (row & 7)) << 4));
  const uint4 q1 = *reinterpret_cast<const uint4*>(base + (((j + 1) ^ (row & 7)) << 4));
  const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    a.v[2 * i] = f.x;
    a.v[2 * i + 1] = f.y;
  }
}

// fp32 aux tile (AXPY): boxes of [128 rows][32 fp32], same swizzle
__device__ __forceinline__ void load_aux_smem_f32(const uint8_t* aux_tile, int row, int c, AuxRegs& a) {
  const uint8_t* base = aux_tile + (c >> 5) * (128 * 128) + row * 128;
  const int j = (c & 31) >> 2;   // first of the four 16-byte chunks
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float4 q = *reinterpret_cast<const float4*>(base + (((j + q4) ^ (row & 7)) << 4));
    a.v[4 * q4] = q.x; a.v[4 * q4 + 1] = q.y; a.v[4 * q4 + 2] = q.z; a.v[4 * q4 + 3] = q.w;
  }
}

template <int KIND>
__device__ __forceinline__ void load_aux(const KParams& p, const GemmEpilogue& e, int64_t aux_off, const float* colv,
                                         int n, int nvalid, bool row_ok, AuxRegs& a) {
  if constexpr (KIND == EPI_SMBWD_BF16) {
    return;                  // always staged through shared memory by TMA
  } else if constexpr (KIND == EPI_DIFF_SQ || KIND == EPI_LNRED_F32 || KIND == EPI_LNRED_BF16 || KIND == EPI_AXPY_B16) {
    if (p.aux_tma) return;   // read from shared memory in the chunk loop instead
    if (!row_ok || nvalid <= 0) return;
    const __nv_bfloat16* ax = static_cast<const __nv_bfloat16*>(e.aux) + aux_off + n;
    if (p.vec_ok && nvalid == 16) {
      const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(ax));
      const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(ax) + 1);
      const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        const float2 f = __bfloat1622float2(h);
        a.v[2 * i] = f.x;
        a.v[2 * i + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) a.v[i] = i < nvalid ? __bfloat162float(ax[i]) : 0.f;
    }
  } else if constexpr (KIND == EPI_AXPY_F32 || KIND == EPI_ACCUM_F32) {
    if (KIND == EPI_AXPY_F32 && p.aux_tma) return;
    if (!row_ok || nvalid <= 0) return;
    const float* ax = (KIND == EPI_AXPY_F32 ? static_cast<const float*>(e.aux) + aux_off
                                            : static_cast<const float*>(e.C) + aux_off) + n;
    if (p.vec_ok && nvalid == 16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 q = *(reinterpret_cast<const float4*>(ax) + i);
        a.v[4 * i] = q.x; a.v[4 * i + 1] = q.y; a.v[4 * i + 2] = q.z; a.v[4 * i + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) a.v[i] = i < nvalid ? ax[i] : 0.f;
    }
  } else if constexpr (KIND == EPI_COSDIST) {
    if (nvalid <= 0) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) a.v[i] = i < nvalid ? __ldg(colv + n + i) : 1.f;
  } else if constexpr (KIND == EPI_MINDIST) {
    if (nvalid <= 0) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) a.v[i] = i < nvalid ? __ldg(colv + n + i) : 3.0e38f;
  }
}

// tile walk shared by the three warp roles: a CTA (or a CTA pair) strides over the tile list
struct Walker {
  int first, step, rank;
};

// Epilogue.  Warp w may only touch TMEM lanes 32*(w%4)..+31, so each lane quarter (32 output rows) is
// served by kHalves warps that split the tile's columns by 128-byte "units" (32 fp32 / 64 bf16 columns).
template <int KIND, int EW>
__device__ __forceinline__ void epilogue_loop(const KParams& p, const CUtensorMap* tma_c, const CUtensorMap* tma_c2,
                                              uint8_t* stage_smem,
                                              uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty,
                                              uint64_t* aux_full, uint64_t* aux_empty, const uint8_t* aux_smem,
                                              int warp, int lane, const Walker wk) {
  constexpr int kEpiWarps = EW;
  constexpr int kHalves = EW / 4;
  const int ew = warp - 2;       // 0..7
  const int quarter = warp & 3;  // TMEM lane quarter this warp may access
  const int half = ew >> 2;      // which of the quarter's two warps
  const int row_in_tile = quarter * 32 + lane;
  const GemmEpilogue& e = p.epi;
  // in a CTA pair the accumulator-free signal goes to the leader's barrier
  const uint32_t empty_remote = wk.rank != 0 ? mapa_shared(smem_u32(tmem_empty), 0) : 0u;
  // store staging: 32 rows x 128 bytes per warp, 128-byte swizzled like the TMA box that reads it
  constexpr bool kBf16Out = (KIND == EPI_STORE_BF16 || KIND == EPI_DIFF_SQ || KIND == EPI_SMBWD_BF16 ||
                             KIND == EPI_BIAS_BF16 || KIND == EPI_LNRED_BF16);
  constexpr bool kLnRed = (KIND == EPI_LNRED_F32 || KIND == EPI_LNRED_BF16);
  constexpr int kUnitCols = kBf16Out ? 64 : 32;            // columns per 128-byte staging row
  // DIFF_SQ / AXPY: the output tile has the shape, type and swizzle of the aux tile it is computed from, so it
  // is written in place over the aux tile (each thread overwrites exactly what it just read) and stored from there
  constexpr bool kInPlace = (KIND == EPI_DIFF_SQ || KIND == EPI_AXPY_F32 || KIND == EPI_SMBWD_BF16);
  // SMBWD has a second output of the same shape: it leaves through the per-warp staging slabs
  constexpr bool kSecond = (KIND == EPI_SMBWD_BF16);
  int sbuf = 0;                  // staging slab in use (double-buffered when p.stage_bufs == 2)
  const bool use_tma_store = p.tma_store && KIND != EPI_ACCUM_F32 && e.C != nullptr && (!kInPlace || p.aux_tma);
  // this warp's chunk walk: units half, half+2, ...; 16-column chunks inside a unit
  auto next_chunk = [&](int c) {
    const int u0 = (c / kUnitCols) * kUnitCols;
    const int uend = min(u0 + kUnitCols, p.block_n);
    if (c + 16 < uend) return c + 16;
    const int c2 = u0 + kHalves * kUnitCols;
    return c2 < p.block_n ? c2 : -1;
  };
  const int c_first = half * kUnitCols < p.block_n ? half * kUnitCols : -1;
  int it = 0;
  for (int tile = wk.first; tile < p.num_tiles; tile += wk.step, ++it) {
    TileCoord t = decode_tile(p, tile);
    t.m0 += wk.rank * BM;
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    // aux tile slot and its barrier phase: one slot per accumulator stage, or a single slot reused every tile
    const int xs = p.aux_bufs == 2 ? as : 0;
    const uint32_t xphase = p.aux_bufs == 2 ? aphase : static_cast<uint32_t>(it & 1);
    const int m = t.m0 + row_in_tile;
    const bool row_ok = m < p.M;
    const int64_t c_off = t.b2 * e.c_b2 + t.b1 * e.c_b1 + (e.c_transposed ? m : static_cast<int64_t>(m) * e.ldc);
    float rowv = 1.f, rowv2 = 0.f;
    if (e.rowv != nullptr && row_ok) rowv = e.rowv[t.b2 * e.rv_b2 + t.b1 * e.rv_b1 + m];
    if (e.rowv2 != nullptr && row_ok) rowv2 = e.rowv2[t.b2 * e.rv_b2 + t.b1 * e.rv_b1 + m];
    float* dx_row = nullptr;                       // EPI_DXSCATTER: this row's destination (support or query gradients)
    if constexpr (KIND == EPI_DXSCATTER) {
      if (row_ok) {
        const int grp = m / e.group_rows, r = m % e.group_rows;
        dx_row = r < e.split_rows
                     ? static_cast<float*>(e.C) + (static_cast<int64_t>(grp) * e.split_rows + r) * p.N
                     : static_cast<float*>(e.C2) +
                           (static_cast<int64_t>(grp) * (e.group_rows - e.split_rows) + (r - e.split_rows)) * p.N;
      }
    }
    const float* colv = e.colv ? e.colv + t.b2 * e.cv_b2 + t.b1 * e.cv_b1 : nullptr;
    const float* colv2 = e.colv2 ? e.colv2 + t.b2 * e.cv_b2 + t.b1 * e.cv_b1 : nullptr;
    // ACCUM reads the output itself, AXPY / DIFF_SQ read `aux`
    const int64_t aux_off = KIND == EPI_ACCUM_F32
                                ? c_off
                                : t.b2 * e.aux_b2 + t.b1 * e.aux_b1 + static_cast<int64_t>(m) * e.ldaux;
    // the operands that do not depend on the MMA are fetched before waiting for it
    AuxRegs cur, nxt;
    if (c_first >= 0)
      load_aux<KIND>(p, e, aux_off, colv, t.n0 + c_first, min(16, p.N - t.n0 - c_first), row_ok, cur);
    if constexpr (KIND == EPI_DIFF_SQ || kLnRed || KIND == EPI_AXPY_F32 || KIND == EPI_AXPY_B16 || KIND == EPI_SMBWD_BF16) {
      if (p.aux_tma) mbar_wait(&aux_full[xs], xphase);
    }
    const uint8_t* aux_tile = aux_smem + xs * p.aux_tile_bytes;
    mbar_wait(&tmem_full[as], aphase);
    tc_fence_after();
    float rsum = 0.f, rsum2 = 0.f;
    float best = 3.4e38f;          // MINDIST: smallest squared distance of this row in this tile and its column
    int best_n = 0;
    const float scale = e.alpha * rowv;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kAccStride;
    // the TMEM read of the next chunk is in flight while the current one is processed
    uint32_t r[16], rn[16];
    int c = c_first;
    if (c >= 0) tmem_ld16(t_row + c, rn);
    while (c >= 0) {
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = rn[i];
      const int cn = next_chunk(c);
      const int n = t.n0 + c;
      const int nvalid = min(16, p.N - n);
      if (cn >= 0) {
        tmem_ld16(t_row + cn, rn);
        load_aux<KIND>(p, e, aux_off, colv, t.n0 + cn, min(16, p.N - t.n0 - cn), row_ok, nxt);
      }
      // a full 128-byte unit that lies inside this tile goes out through TMA; the ragged last unit of a
      // tile (block_n not a multiple of the unit) keeps the direct per-row stores
      const int unit0 = (c / kUnitCols) * kUnitCols;
      const bool staged = use_tma_store && unit0 + kUnitCols <= p.block_n;
      uint8_t* slab = nullptr;
      uint8_t* unit_stage;
      if constexpr (kInPlace) {
        unit_stage = const_cast<uint8_t*>(aux_tile) + (unit0 / kUnitCols) * (128 * 128) + row_in_tile * 128;
        if constexpr (kSecond) slab = stage_smem + (sbuf * kEpiWarps + ew) * 4096;
      } else {
        slab = stage_smem + (sbuf * kEpiWarps + ew) * 4096;
        unit_stage = slab + lane * 128;
      }
      auto emit = [&](const float (&v)[16]) {
        if (staged) {
          const int jb = (c - unit0) * (kBf16Out ? 2 : 4) / 16;     // first 16-byte chunk of this piece
          if constexpr (kBf16Out) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(unit_stage + (((jb + 0) ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(unit_stage + (((jb + 1) ^ (lane & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
          } else {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              *reinterpret_cast<float4*>(unit_stage + (((jb + q4) ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
          }
        } else if (row_ok && nvalid > 0) {
          if ((KIND == EPI_STORE_F32 || KIND == EPI_STORE_BF16) && e.c_transposed) {
            // the warp's 32 lanes are 32 consecutive m: every column is one contiguous 128- / 64-byte run
            if constexpr (kBf16Out) {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(e.C) + c_off + static_cast<int64_t>(n) * e.ldc;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) dst[static_cast<int64_t>(i) * e.ldc] = __float2bfloat16_rn(v[i]);
            } else {
              float* dst = static_cast<float*>(e.C) + c_off + static_cast<int64_t>(n) * e.ldc;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) dst[static_cast<int64_t>(i) * e.ldc] = v[i];
            }
          } else if constexpr (kBf16Out) store16_bf16(static_cast<__nv_bfloat16*>(e.C) + c_off + n, v, nvalid, p.vec_ok);
          else store16_f32(static_cast<float*>(e.C) + c_off + n, v, nvalid, p.vec_ok);
        }
      };
      if ((row_ok && nvalid > 0) || staged) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if constexpr (KIND == EPI_STORE_F32 || KIND == EPI_STORE_BF16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= scale;
          emit(v);
        } else if constexpr (KIND == EPI_ACCUM_F32) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = cur.v[i] + v[i] * scale;
          store16_f32(static_cast<float*>(e.C) + c_off + n, v, nvalid, p.vec_ok);
        } else if constexpr (KIND == EPI_COSDIST) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 1.f - v[i] / (rowv * cur.v[i] + e.eps);
          emit(v);
        } else if constexpr (KIND == EPI_DIFF_SQ) {
          if (p.aux_tma) load_aux_smem(aux_tile, row_in_tile, c, cur);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float d = i < nvalid ? cur.v[i] - v[i] : 0.f;
            v[i] = d;
            rsum += d * d;
          }
          if (e.C != nullptr) emit(v);
        } else if constexpr (KIND == EPI_BIAS_F32 || KIND == EPI_BIAS_BF16) {
          if (nvalid == 16 && ((n & 3) == 0)) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(colv + n) + q4);
              v[4 * q4] = fmaf(v[4 * q4], scale, b4.x);
              v[4 * q4 + 1] = fmaf(v[4 * q4 + 1], scale, b4.y);
              v[4 * q4 + 2] = fmaf(v[4 * q4 + 2], scale, b4.z);
              v[4 * q4 + 3] = fmaf(v[4 * q4 + 3], scale, b4.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], scale, i < nvalid ? __ldg(colv + n + i) : 0.f);
          }
          if constexpr (KIND == EPI_BIAS_BF16) {
            if (e.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
          }
          emit(v);
        } else if constexpr (KIND == EPI_AXPY_F32) {
          if (p.aux_tma) load_aux_smem_f32(aux_tile, row_in_tile, c, cur);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = e.alpha * v[i] + rowv * cur.v[i];
          emit(v);
        } else if constexpr (KIND == EPI_AXPY_B16) {
          if (p.aux_tma) load_aux_smem(aux_tile, row_in_tile, c, cur);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = e.alpha * v[i] + rowv * cur.v[i];
          emit(v);
        } else if constexpr (KIND == EPI_DXSCATTER) {
          if (e.drop_p > 0.f) {
            const uint64_t seed = *e.seed;
            const uint32_t thr = dropout_threshold(e.drop_p);
            const float inv_keep = 1.f / (1.f - e.drop_p);
            const uint64_t g0 = (static_cast<uint64_t>(m) * p.N + n) >> 2;       // N and n are multiples of 4
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              float sc[4];
              dropout_scale4(seed, g0 + q4, thr, inv_keep, sc);
#pragma unroll
              for (int k = 0; k < 4; ++k) v[4 * q4 + k] *= scale * sc[k];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= scale;
          }
          if (row_ok && nvalid > 0) {
            float* dst = dx_row + n;
            if (e.accumulate) {
              if (p.vec_ok && nvalid == 16) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const float4 o = *(reinterpret_cast<const float4*>(dst) + q4);
                  v[4 * q4] += o.x; v[4 * q4 + 1] += o.y; v[4 * q4 + 2] += o.z; v[4 * q4 + 3] += o.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nvalid) v[i] += dst[i];
              }
            }
            store16_f32(dst, v, nvalid, p.vec_ok);
          }
        } else if constexpr (KIND == EPI_MINDIST) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float d2 = fmaxf(rowv + cur.v[i] - 2.f * v[i], 0.f);
            if (i < nvalid && d2 < best) {
              best = d2;
              best_n = n + i;
            }
          }
        } else if constexpr (KIND == EPI_SMBWD_BF16) {
          // p = P~ * (srow / rowsum);  dS = p * (dP_raw - delta)  (in place over P~);  Ps = p  (staging slab)
          load_aux_smem(aux_tile, row_in_tile, c, cur);
          float ps[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            ps[i] = i < nvalid ? cur.v[i] * rowv : 0.f;
            v[i] = ps[i] * (v[i] - rowv2);
          }
          emit(v);
          if (staged) {
            const int jb = (c - unit0) / 8;
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(ps[2 * i], ps[2 * i + 1]);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            uint8_t* srow = slab + lane * 128;
            *reinterpret_cast<uint4*>(srow + (((jb + 0) ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(srow + (((jb + 1) ^ (lane & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
          } else if (row_ok && nvalid > 0) {
            store16_bf16(static_cast<__nv_bfloat16*>(e.C2) + c_off + n, ps, nvalid, p.vec_ok);
          }
        } else if constexpr (kLnRed) {
          if (p.aux_tma) load_aux_smem(aux_tile, row_in_tile, c, cur);
          if (nvalid == 16 && ((n & 3) == 0)) {
            // 8 vector loads of the two column vectors instead of 32 scalar ones
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(colv + n) + q4);
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(colv2 + n) + q4);
              const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int i = 4 * q4 + k;
                v[i] *= scale;
                rsum = fmaf(v[i], gg[k], rsum);
                rsum2 = fmaf(v[i], cur.v[i] - bb[k], rsum2);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] *= scale;
              if (i < nvalid) {
                rsum = fmaf(v[i], __ldg(colv + n + i), rsum);
                rsum2 = fmaf(v[i], cur.v[i] - __ldg(colv2 + n + i), rsum2);
              }
            }
          }
          emit(v);
        }
      }
      if constexpr (!kInPlace || kSecond) {
        if (staged && c + 16 == unit0 + kUnitCols) {
          // the staging slab is complete: hand it to the TMA store
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(kSecond ? tma_c2 : tma_c, slab, t.n0 + unit0, t.m0 + quarter * 32, t.b1, t.b2);
            tma_store_commit();
            // one slab: wait until it has been read.  Two slabs: only the other one has to be free again.
            if (p.stage_bufs == 2) tma_store_wait_read_but_one(); else tma_store_wait_read();
          }
          __syncwarp();
          if (p.stage_bufs == 2) sbuf ^= 1;
        }
      }
      cur = nxt;
      c = cn;
    }
    if constexpr (kInPlace) {
      if (use_tma_store) {
        // one proxy fence per tile, then this warp's units leave straight from the aux tile.  The hand-off stays
        // out of the chunk loop: a fence + warp sync in that loop body cost the P.V products 10 %, even untaken
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          for (int u0 = half * kUnitCols; u0 + kUnitCols <= p.block_n; u0 += kHalves * kUnitCols)
            tma_store_4d(tma_c, aux_tile + (u0 / kUnitCols) * (128 * 128) + quarter * 4096, t.n0 + u0,
                         t.m0 + quarter * 32, t.b1, t.b2);
          tma_store_commit();
          tma_store_wait_read();          // the aux tile is refilled by the producer after this warp's release
        }
      }
    }
    if constexpr (KIND == EPI_DIFF_SQ) {
      if (row_ok && c_first >= 0) atomicAdd(e.rowred + t.b2 * e.rr_b2 + t.b1 * e.rr_b1 + m, rsum);
    }
    if constexpr (KIND == EPI_MINDIST) {
      if (row_ok && c_first >= 0) {
        // non-negative floats order like their bit patterns: one 64-bit atomicMin carries (distance, column)
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(best)) << 32) |
                                       static_cast<unsigned int>(best_n);
        atomicMin(reinterpret_cast<unsigned long long*>(e.rowred) + t.b2 * e.rr_b2 + t.b1 * e.rr_b1 + m, key);
      }
    }
    if constexpr (kLnRed) {
      if (row_ok && c_first >= 0) {
        float* rr = e.rowred + 2 * (t.b2 * e.rr_b2 + t.b1 * e.rr_b1 + m);
        atomicAdd(rr, rsum);
        atomicAdd(rr + 1, rsum2);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (p.aux_tma) mbar_arrive(&aux_empty[xs]);
      if (wk.rank == 0) mbar_arrive(&tmem_empty[as]);
      else mbar_arrive_cluster(empty_remote + as * 8);
    }
  }
  // shared memory must outlive the bulk stores that read it
  if (lane == 0 && p.tma_store) tma_store_wait_read();
}

template <bool CTA2, int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_aux, const __grid_constant__ CUtensorMap tma_c,
                    const __grid_constant__ CUtensorMap tma_c2, const KParams p) {
  constexpr int kEpiWarps = EW;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A tile | B tile)] [2 x aux tile] [4 x 4 KB store staging] [barriers] [tmem ptr]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t stage_bytes = p.a_tile_bytes + p.b_tile_bytes;
  uint8_t* aux_smem = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint8_t* stage_smem = aux_smem + (p.aux_tma ? p.aux_bufs * static_cast<size_t>(p.aux_tile_bytes) : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_smem + (p.own_staging ? p.stage_bufs * kEpiWarps * 4096 : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint64_t* aux_full = bars + 2 * kMaxStages + 4;
  uint64_t* aux_empty = bars + 2 * kMaxStages + 6;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 8);

  // warp index through a shuffle: provably warp-uniform, so the role branches are uniform control flow
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 (leader) issues the MMAs for both; each CTA owns 128 of the tile's 256 rows
  // and half of the B columns
  Walker wk;
  wk.rank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;
  wk.first = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  wk.step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader arms the transaction bytes of the whole pair
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CTA2 ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
      mbar_init(&aux_full[s], 1);
      mbar_init(&aux_empty[s], kEpiWarps);
    }
    if (p.aux_tma) tma_prefetch_desc(&tma_aux);
    if (p.tma_store) tma_prefetch_desc(&tma_c);
    if (p.tma_store && p.epi.kind == EPI_SMBWD_BF16) tma_prefetch_desc(&tma_c2);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CTA2) {
      tmem_alloc_2sm(tmem_ptr, kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------ TMA producer ---------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int n_half = CTA2 ? wk.rank * (p.block_n / 2) : 0;
      for (int tile = wk.first; tile < p.num_tiles; tile += wk.step, ++it) {
        TileCoord t = decode_tile(p, tile);
        t.m0 += wk.rank * BM;
        auto load_aux_tile = [&]() {
          // the epilogue operand tile rides along: same double buffering as the accumulator
          const int as = it & 1;
          const uint32_t aphase = (it >> 1) & 1;
          const int xs = p.aux_bufs == 2 ? as : 0;
          const uint32_t xphase = p.aux_bufs == 2 ? aphase : static_cast<uint32_t>(it & 1);
          mbar_wait(&aux_empty[xs], xphase ^ 1u);
          mbar_expect_tx(&aux_full[xs], p.aux_tile_bytes);
          for (int h = 0; h < p.aux_boxes; ++h)
            tma_load_4d(aux_smem + xs * p.aux_tile_bytes + h * (128 * 128), &tma_aux, &aux_full[xs], t.n0 + p.aux_box_cols * h,
                        t.m0, p.aux_use_b1 ? t.b1 : 0, t.b2);
        };
        // Two aux slots: the slot was released two tiles ago, load it first.  ONE slot: it is released only at the end
        // of the previous tile's epilogue -- waiting for it here would hold back this tile's operand loads and
        // serialise main loop and epilogue (measured: 12.6 us per tile instead of ~6), so it goes after them.
        if (p.aux_tma && p.aux_bufs == 2) load_aux_tile();
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          uint8_t* sb = sa + p.a_tile_bytes;
          uint64_t* fb = &full_bar[stage];
          // CTA pair: both CTAs' loads complete on the leader's barrier; only the leader arrives on it
          // (a per-stage remote arrive from the peer would cost a cluster-scope fence every k-block)
          if (wk.rank == 0) mbar_expect_tx(fb, CTA2 ? 2 * p.tx_bytes : p.tx_bytes);
          const int k0 = kb * BK;
          const int nb0 = t.n0 + n_half;
          if (!CTA2) {
            if (!p.a_mn) {
              tma_load_4d(sa, &tma_a, fb, k0, t.m0, t.b1 * p.a_b1, t.b2);
            } else {
              tma_load_4d(sa, &tma_a, fb, t.m0, k0, t.b1 * p.a_b1, t.b2);
              tma_load_4d(sa + BK * 128, &tma_a, fb, t.m0 + 64, k0, t.b1 * p.a_b1, t.b2);
            }
            if (!p.b_mn) {
              tma_load_4d(sb, &tma_b, fb, k0, nb0, t.b1, t.b2);
            } else {
              for (int h = 0; h < p.b_boxes; ++h)
                tma_load_4d(sb + h * (BK * 128), &tma_b, fb, nb0 + 64 * h, k0, t.b1, t.b2);
            }
          } else {
            if (!p.a_mn) {
              tma_load_4d_2sm(sa, &tma_a, fb, k0, t.m0, t.b1 * p.a_b1, t.b2);
            } else {
              tma_load_4d_2sm(sa, &tma_a, fb, t.m0, k0, t.b1 * p.a_b1, t.b2);
              tma_load_4d_2sm(sa + BK * 128, &tma_a, fb, t.m0 + 64, k0, t.b1 * p.a_b1, t.b2);
            }
            if (!p.b_mn) {
              tma_load_4d_2sm(sb, &tma_b, fb, k0, nb0, t.b1, t.b2);
            } else {
              for (int h = 0; h < p.b_boxes; ++h)
                tma_load_4d_2sm(sb + h * (BK * 128), &tma_b, fb, nb0 + 64 * h, k0, t.b1, t.b2);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (p.aux_tma && p.aux_bufs != 2) load_aux_tile();
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------
    // The whole warp walks the loop in uniform control flow and one elected lane issues: operands then live in
    // uniform registers.  Inside an `if (lane == 0)` every tcgen05.mma cost ~140 cycles of R2UR broadcast loops
    // (measured on the attention kernel, profiles/r02_attn_notes.md) -- more than a 128 x 192 x 16 product
    // occupies the tensor pipe.
    if (wk.rank == 0) {
      const bool leader = elect_one();
      // K-major: step 16 elements (32 B = 2 descriptor units) inside the 128-byte swizzle row.
      // MN-major: step 16 k-rows of 128 B (128 units); LBO = one [BK x 64] box, SBO = 8 k-rows.
      const uint64_t adesc0 = p.a_mn ? make_smem_desc_sw128(0, BK * 128, 1024) : make_smem_desc_sw128(0, 16, 1024);
      const uint64_t bdesc0 = p.b_mn ? make_smem_desc_sw128(0, BK * 128, 1024) : make_smem_desc_sw128(0, 16, 1024);
      const uint32_t astep = p.a_mn ? 128u : 2u, bstep = p.b_mn ? 128u : 2u;
      const uint32_t smem_addr = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = wk.first; tile < p.num_tiles; tile += wk.step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_addr + static_cast<uint32_t>(stage) * stage_bytes;
          const uint32_t sb = sa + p.a_tile_bytes;
          const uint64_t adesc = adesc0 | static_cast<uint64_t>((sa & 0x3FFFFu) >> 4);
          const uint64_t bdesc = bdesc0 | static_cast<uint64_t>((sb & 0x3FFFFu) >> 4);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              if (CTA2) umma_bf16_2sm(d_tmem, adesc + astep * kk, bdesc + bstep * kk, p.idesc, (kb | kk) != 0 ? 1u : 0u);
              else umma_bf16(d_tmem, adesc + astep * kk, bdesc + bstep * kk, p.idesc, (kb | kk) != 0 ? 1u : 0u);
            }
            // smem stage reusable (in both CTAs) once these MMAs retire
            if (CTA2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (leader) {
          if (CTA2) umma_commit_2sm(&tmem_full[as]); else umma_commit(&tmem_full[as]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue -------------------------------------------
#define LMKD_EPI(K) \
  epilogue_loop<K, EW>(p, &tma_c, &tma_c2, stage_smem, tmem_base, tmem_full, tmem_empty, aux_full, aux_empty, aux_smem, warp, lane, wk)
    switch (p.epi.kind) {
      case EPI_STORE_F32: LMKD_EPI(EPI_STORE_F32); break;
      case EPI_STORE_BF16: LMKD_EPI(EPI_STORE_BF16); break;
      case EPI_ACCUM_F32: LMKD_EPI(EPI_ACCUM_F32); break;
      case EPI_COSDIST: LMKD_EPI(EPI_COSDIST); break;
      case EPI_DIFF_SQ: LMKD_EPI(EPI_DIFF_SQ); break;
      case EPI_AXPY_F32: LMKD_EPI(EPI_AXPY_F32); break;
      case EPI_LNRED_F32: LMKD_EPI(EPI_LNRED_F32); break;
      case EPI_BIAS_F32: LMKD_EPI(EPI_BIAS_F32); break;
      case EPI_SMBWD_BF16: LMKD_EPI(EPI_SMBWD_BF16); break;
      case EPI_MINDIST: LMKD_EPI(EPI_MINDIST); break;
      case EPI_BIAS_BF16: LMKD_EPI(EPI_BIAS_BF16); break;
      case EPI_LNRED_BF16: LMKD_EPI(EPI_LNRED_BF16); break;
      case EPI_DXSCATTER: LMKD_EPI(EPI_DXSCATTER); break;
      case EPI_AXPY_B16: LMKD_EPI(EPI_AXPY_B16); break;
      default: break;
    }
#undef LMKD_EPI
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}


// =========================================================================================
// Resident-A variant for the short-K products with a DIFF_SQ epilogue (P.V of the TRX head):
//   D[b2][b1][m][n] = aux[m][n] - sum_k A[m][k] B[k][n],   rowred[m] += sum_n D^2
// The 128 x K operand A (the probability tile) is the same for every column tile of its row block, so it is
// loaded ONCE per (b2, b1, m-tile) and stays in shared memory while the column tiles sweep over N; only B
// (MN-major, BK x BN blocks) streams.  The aux tile no longer owns a double buffer: its upper and lower 64 rows
// travel through the same ring of BN x 64-row slots as the B blocks, behind the blocks of their column tile,
// and are released by the epilogue warps instead of the MMA.  Per column tile this moves K*BN*2 + 128*BN*2 bytes
// into the SM instead of K*(128+BN)*2 + 128*BN*2, and the ring is 6 slots deep instead of 3 stages.
struct RParams {
  int M, N, K, nb1;
  int tiles_m, tiles_n, num_items;
  int block_n, num_kb, slots;
  uint32_t idesc, slot_bytes, a_bytes;
  int aux_use_b1, vec_ok;
  int tma_out;                 // D leaves through the aux slot it was computed in (TMA store) instead of per-lane stores
  GemmEpilogue epi;
};

struct RingPos {
  int pos;
  uint32_t phase;
  __device__ __forceinline__ void advance(int slots) {
    if (++pos == slots) {
      pos = 0;
      phase ^= 1u;
    }
  }
};

constexpr int kREpiWarps = 8;
constexpr int kRThreads = 64 + kREpiWarps * 32;
constexpr int kRMaxSlots = 10;

__global__ void __launch_bounds__(kRThreads, 1)
gemm_resident_a_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                       const __grid_constant__ CUtensorMap tma_aux, const __grid_constant__ CUtensorMap tma_c,
                       const RParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [num_kb x 16 KB resident A] [slots x slot_bytes ring] [barriers] [tmem ptr]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* a_smem = smem;
  uint8_t* ring = smem + p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(p.slots) * p.slot_bytes);
  uint64_t* full_bar = bars;                      // [kRMaxSlots]
  uint64_t* empty_bar = bars + kRMaxSlots;        // [kRMaxSlots]
  uint64_t* tmem_full = bars + 2 * kRMaxSlots;    // [2]
  uint64_t* tmem_empty = bars + 2 * kRMaxSlots + 2;
  uint64_t* a_full = bars + 2 * kRMaxSlots + 4;
  uint64_t* a_empty = bars + 2 * kRMaxSlots + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kRMaxSlots + 6);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_aux);
    for (int i = 0; i < p.slots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);     // one arrival per use: tcgen05.commit (B block) or one epilogue thread (aux)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kREpiWarps);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_batch = p.tiles_m;                // items per (b2, b1)
  const int b_boxes = p.block_n >> 6;             // 64-column boxes per ring slot

  if (warp == 0) {
    // ------------------------------ TMA producer ---------------------------------------
    if (lane == 0) {
      RingPos r{0, 0u};
      int k = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++k) {
        const int batch = item / per_batch;
        const int m0 = (item - batch * per_batch) * BM;
        const int b1 = batch % p.nb1, b2 = batch / p.nb1;
        // the resident operand: free once the MMAs of the previous item have retired
        mbar_wait(a_empty, (static_cast<uint32_t>(k) & 1u) ^ 1u);
        mbar_expect_tx(a_full, p.a_bytes);
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_4d(a_smem + kb * (BM * 128), &tma_a, a_full, kb * BK, m0, b1, b2);
        for (int nt = 0; nt < p.tiles_n; ++nt) {
          const int n0 = nt * p.block_n;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[r.pos], r.phase ^ 1u);
            mbar_expect_tx(&full_bar[r.pos], p.slot_bytes);
            uint8_t* dst = ring + static_cast<size_t>(r.pos) * p.slot_bytes;
            for (int h = 0; h < b_boxes; ++h)
              tma_load_4d(dst + h * (BK * 128), &tma_b, &full_bar[r.pos], n0 + 64 * h, kb * BK, b1, b2);
            r.advance(p.slots);
          }
          for (int hf = 0; hf < 2; ++hf) {          // aux rows m0 + 64 hf .. + 63 of this column tile
            mbar_wait(&empty_bar[r.pos], r.phase ^ 1u);
            mbar_expect_tx(&full_bar[r.pos], p.slot_bytes);
            uint8_t* dst = ring + static_cast<size_t>(r.pos) * p.slot_bytes;
            for (int h = 0; h < b_boxes; ++h)
              tma_load_4d(dst + h * (64 * 128), &tma_aux, &full_bar[r.pos], n0 + 64 * h, m0 + 64 * hf,
                          p.aux_use_b1 ? b1 : 0, b2);
            r.advance(p.slots);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------------------
    // whole warp in uniform control flow, one elected lane issues (see gemm_tcgen05_kernel)
    {
      const bool leader = elect_one();
      const uint64_t adesc0 = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(0, BK * 128, 1024);
      RingPos r{0, 0u};
      int k = 0, it = 0;
      const uint32_t a_addr = smem_u32(a_smem);
      const uint32_t ring_addr = smem_u32(ring);
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++k) {
        mbar_wait(a_full, static_cast<uint32_t>(k) & 1u);
        tc_fence_after();
        for (int nt = 0; nt < p.tiles_n; ++nt, ++it) {
          const int as = it & 1;
          const uint32_t aphase = (it >> 1) & 1;
          mbar_wait(&tmem_empty[as], aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kAccStride;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&full_bar[r.pos], r.phase);
            tc_fence_after();
            const uint32_t sa = a_addr + kb * (BM * 128);
            const uint32_t sb = ring_addr + static_cast<uint32_t>(r.pos) * p.slot_bytes;
            const uint64_t adesc = adesc0 | static_cast<uint64_t>((sa & 0x3FFFFu) >> 4);
            const uint64_t bdesc = bdesc0 | static_cast<uint64_t>((sb & 0x3FFFFu) >> 4);
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk)
                umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 128 * kk, p.idesc, (kb | kk) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[r.pos]);
            }
            __syncwarp();
            r.advance(p.slots);
          }
          r.advance(p.slots);                       // the two aux slots belong to the epilogue
          r.advance(p.slots);
          if (leader) umma_commit(&tmem_full[as]);
          __syncwarp();
        }
        if (leader) umma_commit(a_empty);           // every MMA that read the resident operand has retired
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue (8 warps) ----------------------------------
    const int ew = warp - 2;
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = ew >> 2;                       // which of the quarter's two warps: column units half, half+2, ...
    const int row_in_tile = quarter * 32 + lane;
    const int aux_half = quarter >> 1;              // rows 0..63 travel in the first aux slot, 64..127 in the second
    const int row_in_slot = (quarter & 1) * 32 + lane;
    const GemmEpilogue& e = p.epi;
    RingPos r{0, 0u};
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int batch = item / per_batch;
      const int m0 = (item - batch * per_batch) * BM;
      const int b1 = batch % p.nb1, b2 = batch / p.nb1;
      const int m = m0 + row_in_tile;
      const bool row_ok = m < p.M;
      const int64_t c_off = static_cast<int64_t>(b2) * e.c_b2 + static_cast<int64_t>(b1) * e.c_b1 +
                            static_cast<int64_t>(m) * e.ldc;
      float* rr = e.rowred + static_cast<int64_t>(b2) * e.rr_b2 + static_cast<int64_t>(b1) * e.rr_b1 + m;
      float rsum = 0.f;
      for (int nt = 0; nt < p.tiles_n; ++nt, ++it) {
        const int n0 = nt * p.block_n;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        for (int kb = 0; kb < p.num_kb; ++kb) r.advance(p.slots);     // the B blocks of this column tile
        RingPos mine = r;
        if (aux_half == 1) mine.advance(p.slots);
        r.advance(p.slots);
        r.advance(p.slots);
        mbar_wait(&full_bar[mine.pos], mine.phase);
        uint8_t* aux_slot = ring + static_cast<size_t>(mine.pos) * p.slot_bytes;
        uint8_t* aux_rows = aux_slot + row_in_slot * 128;
        const bool tma_out = p.tma_out != 0 && e.C != nullptr;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kAccStride;
        for (int u0 = half * 64; u0 < p.block_n; u0 += 128) {         // 64-column units of this warp
          uint8_t* box = aux_rows + (u0 >> 6) * (64 * 128);
          uint32_t acc[16], nxt[16];
          tmem_ld16(t_row + u0, nxt);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = nxt[i];
            if (cc < 3) tmem_ld16(t_row + u0 + 16 * (cc + 1), nxt);
            const int c = u0 + 16 * cc;
            const int n = n0 + c;
            const int nvalid = min(16, p.N - n);
            // 16 aux values: two 16-byte chunks of the 128-byte swizzled row
            const int j = cc * 2;
            uint4* s0 = reinterpret_cast<uint4*>(box + ((j ^ (row_in_slot & 7)) << 4));
            uint4* s1 = reinterpret_cast<uint4*>(box + (((j + 1) ^ (row_in_slot & 7)) << 4));
            const uint4 q0 = *s0;
            const uint4 q1 = *s1;
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            float v[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
              const float d0 = 2 * i < nvalid ? f.x - __uint_as_float(acc[2 * i]) : 0.f;
              const float d1 = 2 * i + 1 < nvalid ? f.y - __uint_as_float(acc[2 * i + 1]) : 0.f;
              v[2 * i] = d0;
              v[2 * i + 1] = d1;
              rsum = fmaf(d0, d0, rsum);
              rsum = fmaf(d1, d1, rsum);
            }
            if (tma_out) {
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                o[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *s0 = make_uint4(o[0], o[1], o[2], o[3]);      // in place: exactly the 32 bytes just read
              *s1 = make_uint4(o[4], o[5], o[6], o[7]);
            } else if (e.C != nullptr && row_ok && nvalid > 0) {
              store16_bf16(static_cast<__nv_bfloat16*>(e.C) + c_off + n, v, nvalid, p.vec_ok);
            }
          }
          if (tma_out) {
            // this warp's 32 rows x 64 columns of D sit in the aux box: hand them to the TMA store
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tma_c, aux_slot + (u0 >> 6) * (64 * 128) + (quarter & 1) * 4096, n0 + u0,
                           m0 + quarter * 32, b1, b2);
              tma_store_commit();
            }
          }
        }
        if (tma_out && lane == 0) tma_store_wait_read();    // before the slot goes back to the producer
        // accumulator stage and aux slot are free again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
        // the four warps that read this aux slot meet, then one of them hands it back to the producer
        asm volatile("bar.sync %0, 128;" ::"r"(1 + aux_half) : "memory");
        if ((quarter & 1) == 0 && half == 0 && lane == 0) mbar_arrive(&empty_bar[mine.pos]);
      }
      if (row_ok) atomicAdd(rr, rsum);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// 4-D map (contiguous dim, pitched dim, b1, b2); box = [64, box_rows, 1, 1], 128B swizzle
int make_map(CUtensorMap* map, const GemmOperand& op, int64_t inner, int64_t outer, int nb1, int nb2,
             int box_outer, const char* name) {
  LMKD_CHECK(op.ld % 8 == 0, "gemm operand %s: pitch %lld not a multiple of 8 elements", name,
             (long long)op.ld);
  LMKD_CHECK(op.ld >= inner, "gemm operand %s: pitch %lld < extent %lld", name, (long long)op.ld,
             (long long)inner);
  LMKD_CHECK(nb1 == 1 || op.stride_b1 % 8 == 0, "gemm operand %s: b1 stride not a multiple of 8", name);
  LMKD_CHECK(nb2 == 1 || op.stride_b2 % 8 == 0, "gemm operand %s: b2 stride not a multiple of 8", name);
  TmapSpec t;
  t.base = op.ptr;
  t.dims[0] = (uint64_t)inner; t.dims[1] = (uint64_t)outer; t.dims[2] = (uint64_t)nb1; t.dims[3] = (uint64_t)nb2;
  const uint64_t span = (uint64_t)round_up(op.ld * outer * 2, 16);
  const uint64_t s1 = nb1 > 1 ? (uint64_t)op.stride_b1 * 2 : span;
  const uint64_t s2 = nb2 > 1 ? (uint64_t)op.stride_b2 * 2 : (nb1 > 1 ? s1 * nb1 : span);
  t.strides[0] = (uint64_t)op.ld * 2; t.strides[1] = s1; t.strides[2] = s2;
  t.box[0] = 64; t.box[1] = (uint32_t)box_outer;
  return encode_tmap(map, t, name);
}

// LMKD_GEMM_TMA_STORE=0 keeps the direct per-row epilogue stores (A/B measurements)
bool g_allow_tma_store = [] {
  const char* e = getenv("LMKD_GEMM_TMA_STORE");
  return !(e && e[0] == '0');
}();
// LMKD_GEMM_2CTA=0 forces the single-CTA kernel (A/B measurements)
bool g_allow_cta2 = [] {
  const char* e = getenv("LMKD_GEMM_2CTA");
  return !(e && e[0] == '0');
}();
// LMKD_GEMM_TMA_KINDS: bit k = epilogue kind k may store through TMA (default: all but DIFF_SQ and LNRED, which
// share shared memory with their double-buffered aux tile and were measured faster with direct stores, and
// COSDIST, whose 800-byte output pitch at 200 frames makes every 128-byte box row straddle lines: 1 ms of
// 7.3 at config 4)
int g_tma_kinds = [] {
  const char* e = getenv("LMKD_GEMM_TMA_KINDS");
  return e ? atoi(e) : (((1 << 9) - 1) | (1 << EPI_BIAS_BF16) | (1 << EPI_AXPY_B16)) & ~(1 << EPI_DIFF_SQ) & ~(1 << EPI_LNRED_F32) & ~(1 << EPI_COSDIST);
}();
// LMKD_GEMM_2CTA_MINK: smallest K for which CTAs are paired (default 2048)
int g_cta2_min_k = [] {
  const char* e = getenv("LMKD_GEMM_2CTA_MINK");
  return e ? atoi(e) : 2048;
}();
// LMKD_GEMM_2CTA=2: pair CTAs for every shape with <= 20 % row padding, whatever K (A/B measurements)
bool g_force_cta2 = [] {
  const char* e = getenv("LMKD_GEMM_2CTA");
  return e && e[0] == '2';
}();
// LMKD_GEMM_2CTA_AUX=1: pair CTAs for the aux-tile products whatever their K (measured neutral: 12.71 vs 12.72 ms)
bool g_cta2_aux = [] {
  const char* e = getenv("LMKD_GEMM_2CTA_AUX");
  return e && e[0] == '1';
}();
// LMKD_GEMM_AUX_SINGLE=1: one aux slot instead of two when that yields more operand stages.  Measured slower
// (config-2 step 12.7 -> 13.9 ms, OTAM config 4 14.5 -> 18.3 ms): the aux load latency lands on every tile.
bool g_aux_single = [] {
  const char* e = getenv("LMKD_GEMM_AUX_SINGLE");
  return e && e[0] == '1';
}();
// LMKD_GEMM_EPI8=0: four epilogue warps for every epilogue kind (A/B measurements)
bool g_epi8 = [] {
  const char* e = getenv("LMKD_GEMM_EPI8");
  return !(e && e[0] == '0');
}();
// LMKD_GEMM_AUX_BN: widest column tile of the products whose epilogue reads a bf16 aux tile (DIFF_SQ, LNRED)
int g_aux_bn = [] {
  const char* e = getenv("LMKD_GEMM_AUX_BN");
  const int v = e ? atoi(e) : 192;
  return v >= 128 && v <= 192 ? v / 16 * 16 : 192;
}();
// LMKD_GEMM_STAGE2=0: single store-staging slab per epilogue warp for every shape (A/B measurements)
bool g_stage_bufs2 = [] {
  const char* e = getenv("LMKD_GEMM_STAGE2");
  return !(e && e[0] == '0');
}();
// LMKD_GEMM_AXPY_TMA=0: AXPY reads its fp32 aux operand with per-thread loads (A/B measurements)
bool g_axpy_tma = [] {
  const char* e = getenv("LMKD_GEMM_AXPY_TMA");
  return !(e && e[0] == '0');
}();
// LMKD_GEMM_AXPY_BN: widest column tile of an AXPY product with a TMA-staged aux tile (<= 128)
int g_axpy_bn = [] {
  const char* e = getenv("LMKD_GEMM_AXPY_BN");
  const int v = e ? atoi(e) : 128;
  return v >= 64 && v <= 128 ? v / 16 * 16 : 128;
}();

// aux tile map: [n (contiguous), m, b1 or 1, b2], box = [64, 128, 1, 1], 128B swizzle
int make_aux_map(CUtensorMap* map, const GemmEpilogue& e, int M, int N, int nb1, int nb2, bool use_b1, bool f32,
                 int box_rows = 128) {
  const int esz = f32 ? 4 : 2;
  TmapSpec t;
  t.dtype = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  t.base = e.aux;
  t.dims[0] = (uint64_t)N; t.dims[1] = (uint64_t)M; t.dims[2] = (uint64_t)(use_b1 ? nb1 : 1); t.dims[3] = (uint64_t)nb2;
  const uint64_t span = (uint64_t)round_up(e.ldaux * (int64_t)M * esz, 16);
  const uint64_t s1 = use_b1 && nb1 > 1 ? (uint64_t)e.aux_b1 * esz : span;
  const uint64_t s2 = nb2 > 1 ? (uint64_t)e.aux_b2 * esz : (use_b1 && nb1 > 1 ? s1 * nb1 : span);
  t.strides[0] = (uint64_t)e.ldaux * esz; t.strides[1] = s1; t.strides[2] = s2;
  t.box[0] = (uint32_t)(128 / esz); t.box[1] = (uint32_t)box_rows;
  return encode_tmap(map, t, "aux");
}

// output map: [n (contiguous), m, b1, b2], box = [128 bytes of columns, 32 rows, 1, 1], 128B swizzle
int make_out_map(CUtensorMap* map, void* C, int64_t ldc, int64_t c_b1, int64_t c_b2, int M, int N, int nb1, int nb2,
                 bool bf16) {
  const int esz = bf16 ? 2 : 4;
  TmapSpec t;
  t.dtype = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  t.base = C;
  t.dims[0] = (uint64_t)N; t.dims[1] = (uint64_t)M; t.dims[2] = (uint64_t)nb1; t.dims[3] = (uint64_t)nb2;
  const uint64_t span = (uint64_t)round_up(ldc * (int64_t)M * esz, 16);
  const uint64_t s1 = nb1 > 1 ? (uint64_t)c_b1 * esz : span;
  const uint64_t s2 = nb2 > 1 ? (uint64_t)c_b2 * esz : (nb1 > 1 ? s1 * nb1 : span);
  t.strides[0] = (uint64_t)ldc * esz; t.strides[1] = s1; t.strides[2] = s2;
  t.box[0] = (uint32_t)(128 / esz); t.box[1] = 32;
  t.promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  return encode_tmap(map, t, "out");
}
int make_out_map(CUtensorMap* map, const GemmEpilogue& e, int M, int N, int nb1, int nb2, bool bf16) {
  return make_out_map(map, e.C, e.ldc, e.c_b1, e.c_b2, M, N, nb1, nb2, bf16);
}

inline bool is_lnred(int kind) { return kind == EPI_LNRED_F32 || kind == EPI_LNRED_BF16; }

int pick_block_n(int N) {
  if (N >= 256) {
    // prefer an exact divisor in [128, 256] (multiple of 16) to avoid a ragged last tile
    for (int bn = 256; bn >= 128; bn -= 16)
      if (N % bn == 0) return bn;
    return 256;
  }
  return (int)round_up(N, 16);
}

// LMKD_GEMM_RESIDENT_TMA=0: the resident-A kernel writes D with per-lane stores instead of TMA stores from the aux slot
bool g_resident_tma = [] {
  const char* e = getenv("LMKD_GEMM_RESIDENT_TMA");
  return !(e && e[0] == '0');
}();
// LMKD_GEMM_RESIDENT_BN: preferred column tile of the resident-A kernel (64, 128 or 192)
int g_resident_bn = [] {
  const char* e = getenv("LMKD_GEMM_RESIDENT_BN");
  const int v = e ? atoi(e) : 128;
  return (v == 64 || v == 128 || v == 192) ? v : 128;
}();
// LMKD_GEMM_RESIDENT_A=0: P.V products go through the generic kernel (A/B measurements).  GEMM time of the
// config-2 step, event-timed: generic kernel with 8 epilogue warps 7.89 ms; resident-A with per-lane stores 7.83 ms
// (192-column tiles: 8.05 ms); resident-A with the TMA-stored epilogue 7.61 ms -- see profiles/r01_notes.md.
bool g_resident_a = [] {
  const char* e = getenv("LMKD_GEMM_RESIDENT_A");
  return !(e && e[0] == '0');
}();

// true if the product qualifies for (and was launched through) the resident-A kernel
int launch_resident_a(const GemmDesc& g, cudaStream_t stream, bool* taken) {
  *taken = false;
  const GemmEpilogue& e = g.epi;
  if (!g_resident_a || e.kind != EPI_DIFF_SQ || g.A.mn_major || !g.B.mn_major || g.block_n > 0) return 0;
  const int num_kb = (int)ceil_div(g.K, BK);
  const bool aux_ok = e.aux != nullptr && (reinterpret_cast<uintptr_t>(e.aux) % 16 == 0) && e.ldaux % 8 == 0 &&
                      (g.nb1 == 1 || e.aux_b1 == 0 || e.aux_b1 % 8 == 0) && (g.nb2 == 1 || e.aux_b2 % 8 == 0);
  if (!aux_ok || g.N % 16 != 0 || g.N < 128 || num_kb > 6 || !e.rowred) return 0;
  RParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.nb1 = g.nb1;
  // 128 columns = two 64-column units, one per warp of a TMEM lane quarter (192 leaves one of them with twice the
  // work), and 16 KB slots: the ring then holds more slots than a column tile has blocks (K/64 B blocks + 2 aux
  // halves), so the aux halves of tile n never wait for the epilogue of tile n-1 to hand a slot back
  const int pref = g_resident_bn;
  p.block_n = pref;
  if (g.N % pref != 0)
    for (int bn = 192; bn >= 64; bn -= 64)
      if (g.N % bn == 0) { p.block_n = bn; break; }
  p.tiles_m = (int)ceil_div(g.M, BM);
  p.tiles_n = (int)ceil_div(g.N, p.block_n);
  if (p.tiles_n < 2) return 0;                         // nothing to reuse the resident operand for
  const int64_t items = (int64_t)p.tiles_m * g.nb1 * g.nb2;
  LMKD_CHECK(items < (1ll << 31), "gemm: too many tiles");
  p.num_items = (int)items;
  p.num_kb = num_kb;
  p.a_bytes = (uint32_t)num_kb * BM * 128;
  p.slot_bytes = (uint32_t)p.block_n * 128;            // 64 rows x block_n bf16
  const int tail = 1024 + (2 * kRMaxSlots + 8) * 8 + 16;
  int slots = (int)((227 * 1024 - tail - (int)p.a_bytes) / (int)p.slot_bytes);
  if (slots > kRMaxSlots) slots = kRMaxSlots;
  if (slots < 4) return 0;
  p.slots = slots;
  p.idesc = make_idesc_bf16(BM, p.block_n, 0, 1);
  p.aux_use_b1 = (g.nb1 > 1 && e.aux_b1 != 0) ? 1 : 0;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(e.C) % 16) == 0) && (e.ldc % 8 == 0) &&
             (g.nb1 == 1 || e.c_b1 % 8 == 0) && (g.nb2 == 1 || e.c_b2 % 8 == 0);
  p.epi = e;
  p.tma_out = (g_resident_tma && e.C != nullptr && p.vec_ok && (g.nb1 == 1 || e.c_b1 > 0) && (g.nb2 == 1 || e.c_b2 > 0)) ? 1 : 0;
  CUtensorMap ma, mb, maux, mc;
  memset(&mc, 0, sizeof(mc));
  if (p.tma_out)
    if (int rc = make_out_map(&mc, e, g.M, g.N, g.nb1, g.nb2, true)) return rc;
  if (int rc = make_map(&ma, g.A, g.K, g.M, g.nb1, g.nb2, BM, "A")) return rc;
  if (int rc = make_map(&mb, g.B, g.N, g.K, g.nb1, g.nb2, BK, "B(mn)")) return rc;
  if (int rc = make_aux_map(&maux, e, g.M, g.N, g.nb1, g.nb2, p.aux_use_b1 != 0, false, 64)) return rc;
  const size_t smem = (size_t)p.a_bytes + (size_t)slots * p.slot_bytes + tail;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(gemm_resident_a_kernel), 227 * 1024)) return rc;
  KernelTimingScope timing(TIME_TENSOR, stream, 2.0 * g.M * g.N * g.K * g.nb1 * g.nb2);
  if (int rc = timing.begin()) return rc;
  const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
  gemm_resident_a_kernel<<<grid, kRThreads, smem < 120 * 1024 ? 120 * 1024 : smem, stream>>>(ma, mb, maux,
                                                                                             p.tma_out ? mc : ma, p);
  LMKD_LAUNCH_CHECK("gemm_resident_a_kernel");
  if (int rc = timing.end()) return rc;
  *taken = true;
  return 0;
}

}  // namespace

int gemm_bf16(const GemmDesc& g, cudaStream_t stream) {
  LMKD_CHECK(g.M > 0 && g.N > 0 && g.K > 0 && g.nb1 > 0 && g.nb2 > 0, "gemm: empty problem %d %d %d",
             g.M, g.N, g.K);
  LMKD_CHECK(g.epi.C != nullptr || g.epi.kind == EPI_DIFF_SQ || g.epi.kind == EPI_MINDIST, "gemm: null output");
  LMKD_CHECK(g.epi.kind >= 0 && g.epi.kind <= EPI_AXPY_B16, "gemm: unknown epilogue kind %d", g.epi.kind);
  LMKD_CHECK(!g.epi.c_transposed || g.epi.kind == EPI_STORE_F32 || g.epi.kind == EPI_STORE_BF16,
             "gemm: transposed output needs a plain store epilogue");
  {
    bool taken = false;
    if (int rc = launch_resident_a(g, stream, &taken)) return rc;
    if (taken) return 0;
  }
  KParams p{};
  p.M = g.M;
  p.N = g.N;
  p.K = g.K;
  p.nb1 = g.nb1;
  p.block_n = g.block_n > 0 ? g.block_n : pick_block_n(g.N);
  if (g.block_n <= 0 && (g.epi.kind == EPI_AXPY_F32 || g.epi.kind == EPI_AXPY_B16) && g_axpy_tma && p.block_n > 128) {
    // fp32 aux tile: 2 x 128 x BN x 4 bytes next to the operand ring
    p.block_n = g_axpy_bn;
    for (int bn = g_axpy_bn; bn >= 64; bn -= 16)
      if (g.N % bn == 0) { p.block_n = bn; break; }
  }
  if (g.block_n <= 0 && (g.epi.kind == EPI_DIFF_SQ || is_lnred(g.epi.kind) || g.epi.kind == EPI_SMBWD_BF16) &&
      p.block_n > g_aux_bn) {
    // the TMA-staged aux tile (2 x 128 x BN bf16) shares shared memory with the operand ring
    p.block_n = 128;
    for (int bn = g_aux_bn; bn >= 128; bn -= 16)
      if (g.N % bn == 0) { p.block_n = bn; break; }
  }
  LMKD_CHECK(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "gemm: bad block_n %d", p.block_n);
  // CTA pairs (cta_group::2): 256-row tiles, each CTA loads its 128 A rows and HALF of the B tile, so the
  // L2 -> SM operand traffic per flop drops by up to 1.5x.  Used when the extra row padding is small.
  const int64_t rows1 = ceil_div(g.M, BM) * BM, rows2 = ceil_div(g.M, 2 * BM) * 2 * BM;
  // Measured on B200 (profiles/r01_gemm_1cta_vs_2cta.txt): +7..17 % for K >= 2048, neutral or slightly
  // negative for the short-K attention products, whose tiles are epilogue-bound.
  const bool aux_kind = g.epi.kind == EPI_DIFF_SQ || is_lnred(g.epi.kind);
  const bool cta2 = g_allow_cta2 && g.M > BM && p.block_n >= 32 && sm_count() >= 2 &&
                    (g_force_cta2 ? rows2 * 10 <= rows1 * 12
                                  : (rows2 * 100 <= rows1 * 110 && (g.K >= g_cta2_min_k || (aux_kind && g_cta2_aux))));
  p.bm = cta2 ? 2 * BM : BM;
  p.tiles_m = (int)ceil_div(g.M, p.bm);
  p.tiles_n = (int)ceil_div(g.N, p.block_n);
  const int64_t nt = (int64_t)p.tiles_m * p.tiles_n * g.nb1 * g.nb2;
  LMKD_CHECK(nt < (1ll << 31), "gemm: too many tiles");
  p.num_tiles = (int)nt;
  p.num_kb = (int)ceil_div(g.K, BK);
  p.a_mn = g.A.mn_major;
  p.b_mn = g.B.mn_major;
  const int bn_cta = cta2 ? p.block_n / 2 : p.block_n;     // B columns held by one CTA
  p.b_boxes = (int)ceil_div(bn_cta, 64);
  p.idesc = make_idesc_bf16(p.bm, p.block_n, p.a_mn, p.b_mn);
  p.a_tile_bytes = BM * 128;
  p.b_tile_bytes = (uint32_t)round_up(bn_cta, 64) * 128;
  p.tx_bytes = BM * 128 + (p.b_mn ? p.b_boxes * BK * 128 : bn_cta * 128);
  const uint32_t stage_bytes = p.a_tile_bytes + p.b_tile_bytes;
  const GemmEpilogue& e0 = g.epi;
  // DIFF_SQ: prefetch the aux tile with TMA when its layout allows (16-byte aligned strides)
  const bool aux_f32 = e0.kind == EPI_AXPY_F32;
  const int aux_al = aux_f32 ? 4 : 8;                       // elements per 16 bytes
  p.aux_tma = (e0.kind == EPI_DIFF_SQ || is_lnred(e0.kind) || e0.kind == EPI_SMBWD_BF16 ||
               (e0.kind == EPI_AXPY_B16 && g_axpy_tma && p.block_n <= 128) ||
               (aux_f32 && g_axpy_tma && p.block_n <= 128)) &&
              e0.aux != nullptr && (reinterpret_cast<uintptr_t>(e0.aux) % 16 == 0) &&
              e0.ldaux % aux_al == 0 && (g.nb1 == 1 || e0.aux_b1 == 0 || e0.aux_b1 % aux_al == 0) &&
              (g.nb2 == 1 || e0.aux_b2 % aux_al == 0);
  p.aux_box_cols = aux_f32 ? 32 : 64;
  p.aux_boxes = (int)ceil_div(p.block_n, p.aux_box_cols);
  p.aux_use_b1 = (g.nb1 > 1 && e0.aux_b1 != 0) ? 1 : 0;
  p.aux_tile_bytes = p.aux_tma ? (uint32_t)p.aux_boxes * 128 * 128 : 0;
  // epilogue stores through TMA when the output layout qualifies (16-byte aligned base and strides)
  p.out_bf16 = (e0.kind == EPI_STORE_BF16 || e0.kind == EPI_DIFF_SQ || e0.kind == EPI_SMBWD_BF16 ||
                e0.kind == EPI_BIAS_BF16 || e0.kind == EPI_LNRED_BF16) ? 1 : 0;
  {
    const int esz0 = p.out_bf16 ? 2 : 4;
    const int al = 16 / esz0;
    p.tma_store = g_allow_tma_store && ((g_tma_kinds >> e0.kind) & 1) && e0.C != nullptr && e0.kind != EPI_ACCUM_F32 &&
                  !e0.c_transposed &&
                  (reinterpret_cast<uintptr_t>(e0.C) % 16 == 0) && e0.ldc % al == 0 &&
                  (g.nb1 == 1 || (e0.c_b1 % al == 0 && e0.c_b1 > 0)) && (g.nb2 == 1 || (e0.c_b2 % al == 0 && e0.c_b2 > 0));
  }
  const bool inplace_kind = e0.kind == EPI_DIFF_SQ || e0.kind == EPI_AXPY_F32 || e0.kind == EPI_SMBWD_BF16;
  if (inplace_kind && !p.aux_tma) p.tma_store = 0;          // in-place kinds stage in the aux tile only
  // SMBWD: first output in place over the aux tile, second output through staging slabs
  const bool own_staging = p.tma_store && (!inplace_kind || e0.kind == EPI_SMBWD_BF16);
  p.own_staging = own_staging ? 1 : 0;
  const int epi_warps = (g_epi8 && (e0.kind == EPI_DIFF_SQ || is_lnred(e0.kind) || e0.kind == EPI_AXPY_F32 ||
                                    e0.kind == EPI_AXPY_B16)) ? 8 : 4;
  const int threads = 64 + epi_warps * 32;
  // short contractions are epilogue-bound: a second staging slab per warp keeps the stores flowing
  // (taken only when it does not cost an operand stage the contraction could use)
  auto plan = [&](int bufs, int* tail_out) {
    const int t = 1024 /*align slack*/ + (2 * kMaxStages + 8) * 8 + 16 + p.aux_bufs * (int)p.aux_tile_bytes +
                  (own_staging ? bufs * epi_warps * 4096 : 0);
    *tail_out = t;
    const int st = (int)((220 * 1024 - t) / stage_bytes);
    return st > kMaxStages ? kMaxStages : st;
  };
  int tail = 0, tail2 = 0;
  p.aux_bufs = 2;
  int stages = plan(1, &tail);
  // SMBWD rides on a long contraction (K = d): its aux tile has a whole tile's main loop to arrive, so one
  // slot is enough and the other 48 KB buy an operand stage
  if ((g_aux_single || e0.kind == EPI_SMBWD_BF16) && p.aux_tma && stages < p.num_kb) {
    // a single aux slot exposes its load latency once per tile but buys operand stages for short contractions
    p.aux_bufs = 1;
    const int st1 = plan(1, &tail2);
    if (st1 > stages) { stages = st1; tail = tail2; } else p.aux_bufs = 2;
  }
  p.stage_bufs = 1;
  if (own_staging && g_stage_bufs2 && p.num_kb <= 10) {
    const int st2 = plan(2, &tail2);
    if (st2 >= 2 && (st2 == stages || st2 >= p.num_kb)) {
      p.stage_bufs = 2;
      stages = st2;
      tail = tail2;
    }
  }
  LMKD_CHECK(stages >= 2, "gemm: not enough shared memory for 2 stages");
  p.stages = stages;
  p.epi = g.epi;
  // vector stores need 16-byte alignment of every row start
  const GemmEpilogue& e = g.epi;
  const bool bf16_out = (e.kind == EPI_STORE_BF16 || e.kind == EPI_DIFF_SQ || e.kind == EPI_SMBWD_BF16 ||
                         e.kind == EPI_BIAS_BF16 || e.kind == EPI_LNRED_BF16);
  const int esz = bf16_out ? 2 : 4;
  const int q = 16 / esz * (bf16_out ? 2 : 1);  // bf16 path writes 2 x 16B per chunk -> 16 elems
  p.vec_ok = ((reinterpret_cast<uintptr_t>(e.C) % 16) == 0) && (e.ldc % (16 / esz) == 0) &&
             (g.nb1 == 1 || e.c_b1 % (16 / esz) == 0) && (g.nb2 == 1 || e.c_b2 % (16 / esz) == 0);
  (void)q;
  if (e.kind == EPI_DIFF_SQ)
    p.vec_ok = p.vec_ok && (reinterpret_cast<uintptr_t>(e.aux) % 16 == 0) && e.ldaux % 8 == 0 &&
               (g.nb1 == 1 || e.aux_b1 % 8 == 0) && (g.nb2 == 1 || e.aux_b2 % 8 == 0);
  if (e.kind == EPI_AXPY_B16)
    p.vec_ok = p.vec_ok && (reinterpret_cast<uintptr_t>(e.aux) % 16 == 0) && e.ldaux % 8 == 0 &&
               (g.nb1 == 1 || e.aux_b1 % 8 == 0) && (g.nb2 == 1 || e.aux_b2 % 8 == 0);
  if (e.kind == EPI_AXPY_F32)
    p.vec_ok = p.vec_ok && (reinterpret_cast<uintptr_t>(e.aux) % 16 == 0) && e.ldaux % 4 == 0 &&
               (g.nb1 == 1 || e.aux_b1 % 4 == 0) && (g.nb2 == 1 || e.aux_b2 % 4 == 0);
  if (e.kind == EPI_COSDIST) LMKD_CHECK(e.rowv && e.colv, "gemm: COSDIST needs rowv and colv");
  if (e.kind == EPI_DIFF_SQ) LMKD_CHECK(e.aux && e.rowred, "gemm: DIFF_SQ needs aux and rowred");
  if (is_lnred(e.kind)) {
    LMKD_CHECK(e.aux && e.rowred && e.colv && e.colv2, "gemm: LNRED needs aux, rowred, colv and colv2");
    LMKD_CHECK(p.aux_tma, "gemm: LNRED needs a TMA-compatible aux layout");
  }
  if (e.kind == EPI_AXPY_F32 || e.kind == EPI_AXPY_B16) LMKD_CHECK(e.aux && e.rowv, "gemm: AXPY needs aux and rowv");
  if (e.kind == EPI_BIAS_F32 || e.kind == EPI_BIAS_BF16) LMKD_CHECK(e.colv, "gemm: BIAS needs colv");
  if (e.kind == EPI_MINDIST) LMKD_CHECK(e.rowv && e.colv && e.rowred, "gemm: MINDIST needs rowv, colv and rowred");
  if (e.kind == EPI_DXSCATTER) {
    LMKD_CHECK(e.C2 && e.group_rows > 0 && e.split_rows >= 0 && e.split_rows <= e.group_rows && g.M % e.group_rows == 0 &&
                   g.nb1 == 1 && g.nb2 == 1 && g.N % 4 == 0 && (e.drop_p <= 0.f || e.seed != nullptr),
               "gemm: DXSCATTER needs two outputs, a row grouping that divides M, N %% 4 == 0 and a seed when dropping");
    p.vec_ok = (reinterpret_cast<uintptr_t>(e.C) % 16 == 0) && (reinterpret_cast<uintptr_t>(e.C2) % 16 == 0);
  }
  if (e.kind == EPI_SMBWD_BF16) {
    LMKD_CHECK(e.aux && e.rowv && e.rowv2 && e.C2, "gemm: SMBWD needs aux, rowv, rowv2 and C2");
    LMKD_CHECK(p.aux_tma, "gemm: SMBWD needs a TMA-compatible aux layout");
    p.vec_ok = p.vec_ok && (reinterpret_cast<uintptr_t>(e.C2) % 16 == 0);
  }

  CUtensorMap ma, mb, maux, mc, mc2;
  memset(&maux, 0, sizeof(maux));
  memset(&mc, 0, sizeof(mc));
  int rc;
  if (p.tma_store) {
    rc = make_out_map(&mc, g.epi, g.M, g.N, g.nb1, g.nb2, p.out_bf16 != 0);
    if (rc) return rc;
  }
  const bool second = p.tma_store && e.kind == EPI_SMBWD_BF16;
  if (second) {
    rc = make_out_map(&mc2, e.C2, e.ldc, e.c_b1, e.c_b2, g.M, g.N, g.nb1, g.nb2, true);
    if (rc) return rc;
  }
  if (p.aux_tma) {
    rc = make_aux_map(&maux, g.epi, g.M, g.N, g.nb1, g.nb2, p.aux_use_b1 != 0, aux_f32);
    if (rc) return rc;
  }
  p.a_b1 = (g.nb1 > 1 && g.A.stride_b1 == 0) ? 0 : 1;
  const int a_nb1 = p.a_b1 ? g.nb1 : 1;
  if (!p.a_mn) rc = make_map(&ma, g.A, g.K, g.M, a_nb1, g.nb2, BM, "A");
  else rc = make_map(&ma, g.A, g.M, g.K, a_nb1, g.nb2, BK, "A(mn)");
  if (rc) return rc;
  if (!p.b_mn) rc = make_map(&mb, g.B, g.K, g.N, g.nb1, g.nb2, bn_cta, "B");
  else rc = make_map(&mb, g.B, g.N, g.K, g.nb1, g.nb2, BK, "B(mn)");
  if (rc) return rc;

  const size_t smem = (size_t)stages * stage_bytes + tail;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<false, 4>), 227 * 1024)) return rc;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<true, 4>), 227 * 1024)) return rc;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<false, 8>), 227 * 1024)) return rc;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(gemm_tcgen05_kernel<true, 8>), 227 * 1024)) return rc;
  // >= 120 KB of dynamic smem keeps it at one CTA per SM (each CTA allocates all 512 TMEM columns)
  const size_t smem_launch = smem < 120 * 1024 ? 120 * 1024 : smem;
  KernelTimingScope timing(TIME_TENSOR, stream, 2.0 * g.M * g.N * g.K * g.nb1 * g.nb2);
  if (int rc = timing.begin()) return rc;
  if (!cta2) {
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    if (epi_warps == 8)
      gemm_tcgen05_kernel<false, 8><<<grid, threads, smem_launch, stream>>>(ma, mb, p.aux_tma ? maux : ma,
                                                                            p.tma_store ? mc : ma, second ? mc2 : ma, p);
    else
      gemm_tcgen05_kernel<false, 4><<<grid, threads, smem_launch, stream>>>(ma, mb, p.aux_tma ? maux : ma,
                                                                            p.tma_store ? mc : ma, second ? mc2 : ma, p);
  } else {
    const int pairs = p.num_tiles < sm_count() / 2 ? p.num_tiles : sm_count() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_launch;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (epi_warps == 8)
      LMKD_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<true, 8>, ma, mb, p.aux_tma ? maux : ma,
                                   p.tma_store ? mc : ma, second ? mc2 : ma, p));
    else
      LMKD_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<true, 4>, ma, mb, p.aux_tma ? maux : ma,
                                   p.tma_store ? mc : ma, second ? mc2 : ma, p));
  }
  LMKD_LAUNCH_CHECK("gemm_tcgen05_kernel");
  return timing.end();
}

void gemm_timing_enable(int on) { kernel_timing_enable(on); }

int gemm_timing_read(double* ms, double* flops, int* launches) { return kernel_timing_read(TIME_TENSOR, ms, flops, launches); }

}  // namespace lmkd
