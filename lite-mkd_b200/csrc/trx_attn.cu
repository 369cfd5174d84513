// Fused class-grouped tuple attention of the TRX head, forward (sm_100a).
//
// One work item = (episode b, class c, 128 query-tuple rows).  For it the CTA computes
//     S  = Kq_rows . Ks_c^T                       tcgen05.mma, fp32 accumulator in tensor memory
//     P~ = exp((S - rowmax) / sqrt(d))            softmax warps: TMEM -> registers -> TMEM (bf16, two per column)
//     O  = P~ . Vs_c                              tcgen05.mma with the A operand READ FROM TENSOR MEMORY,
//                                                 64 output columns at a time into a ring of accumulator stages
//     diff = v_q - O / rowsum ;  rowred += |diff|^2 ;  rowdot += <diff, O / rowsum>
// so neither the scores nor the probabilities touch shared memory or HBM (TRX.py:125-141 materialises both per
// class; round 1 of this repo wrote fp32 scores, re-read them in a softmax kernel and re-read bf16 P in the P.V
// product).  Training passes additionally write the un-normalised P~ (bf16) and 1/rowsum for the backward.
//
// Warp roles (12 warps): 0 = TMA producer of the "column" ring (Ks k-blocks, then Vs chunks), 1 = TMA producer of
// the "row" ring (Kq k-blocks, then the v_q tiles the epilogue combines with), 2 = TMEM allocator + MMA issuer,
// 3 = TMA store issuer (takes finished row-ring tiles from the compute warps, stores them, waits for the stores to
// have read shared memory and only then returns the slots -- so no compute warp ever waits on a store),
// 4..11 = compute: softmax (two warps per TMEM lane quarter, splitting the columns) then the P.V epilogue (two
// groups of four warps alternating over the 64-column chunks).  The two rings are separate because their slots
// live for different times: column slots are released by tcgen05.commit, v_q tiles only after the epilogue has
// written diff back in place and the TMA store has read it.
//
// Tensor memory (512 columns): [0, KTp) S, overwritten in place by P~ (each softmax half packs into the front
// of its own column range); [o_base, o_base + 64 * nacc) output stages.  KTp <= 384 keeps >= 2 stages.
#include "trx_attn.cuh"

#include <cmath>

#include "gemm.cuh"
#include "tmap.cuh"

namespace lmkd {

namespace {

constexpr int kThreads = 384;
constexpr int kComputeThreads = 256;
constexpr int kMaxRing = 8;
constexpr int kMaxAcc = 6;
constexpr uint32_t kASlot = 128 * 128;          // 128 rows x 64 bf16, 128-byte swizzled
constexpr uint32_t kTmemCols = 512;
constexpr int kTailBytes = 4096;                // barriers + TMEM pointer + softmax exchange

struct AttnParams {
  int way, NqT, KTp, T;
  int tiles_m, num_items;
  int nrow, nbox;                               // Ks / Vs TMA boxes: nbox boxes of nrow rows per slot
  int load_rows;                                // rows per box actually fetched (== nrow; less only in the traffic experiment)
  int n1, n2;                                   // S is issued as one or two MMAs along N (N <= 256 each)
  int nchunks;                                  // d / 64: k-blocks of S = K.K^T and output chunks of P.V
  int nk16;                                     // KTp / 16: MMAs per output chunk
  int nch, chs;                                 // 16-column chunks of S; the first softmax half owns [0, chs)
  int nunits;                                   // 64-column units of P~ written out (training)
  int na, nb, nacc, o_base;
  uint32_t slot_b;
  uint32_t idesc_qk1, idesc_qk2, idesc_pv;
  float scale_log2;                             // log2(e) / sqrt(d)
  int write_diff, write_p;
  const int* cnt;
  float* rowred;
  float* rowdot;
  float* linv;
};

struct Ring {
  int pos;
  uint32_t phase;
  __device__ __forceinline__ void advance(int n) {
    if (++pos == n) {
      pos = 0;
      phase ^= 1u;
    }
  }
};

struct Item {
  int b, c, m0;
};
__device__ __forceinline__ Item decode_item(const AttnParams& p, int item) {
  Item t;
  const int mt = item % p.tiles_m;
  const int bc = item / p.tiles_m;
  t.c = bc % p.way;
  t.b = bc / p.way;
  t.m0 = mt * 128;
  return t;
}

// TMEM column of the 16 probabilities of S chunk j (8 packed columns): each softmax half packs into the front
// of the column range it read its scores from
__device__ __forceinline__ uint32_t p_col(int chs, int j) {
  return j < chs ? 8u * j : 16u * chs + 8u * (j - chs);
}

// NK16 = KTp / 16 as a compile-time constant for the two config-2 shapes (9: pairs, 18: triples; 0 = generic): the
// single MMA-issuing warp then runs straight-line tcgen05.mma sequences -- with run-time loop bounds its per-chunk
// bookkeeping (~150 instructions) took twice as long as the nine 128x64x16 products of a pairs chunk.
template <int NK16>
__global__ void __launch_bounds__(kThreads, 1)
trx_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_kq, const __grid_constant__ CUtensorMap tm_ks,
                    const __grid_constant__ CUtensorMap tm_vq, const __grid_constant__ CUtensorMap tm_vs,
                    const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_p,
                    const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* ring_b = smem;
  uint8_t* ring_a = ring_b + static_cast<size_t>(p.nb) * p.slot_b;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_a + static_cast<size_t>(p.na) * kASlot);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + kMaxRing;
  uint64_t* full_b = bars + 2 * kMaxRing;
  uint64_t* empty_b = bars + 3 * kMaxRing;
  uint64_t* tile_ready = bars + 4 * kMaxRing;   // row-ring tile finished by its epilogue group -> store warp
  uint64_t* acc_full = bars + 5 * kMaxRing;
  uint64_t* acc_empty = acc_full + kMaxAcc;
  uint64_t* s_full = acc_empty + kMaxAcc;
  uint64_t* p_ready = s_full + 1;
  uint64_t* p_drained = s_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_full + 3);
  float* xmax = reinterpret_cast<float*>(s_full + 4);     // [2][128]
  float* xsum = xmax + 256;                                // [2][128]

  // warp index through a shuffle: provably warp-uniform, so the role branches below are uniform control flow
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nk16 = NK16 > 0 ? NK16 : p.nk16;           // = nch: 16-column chunks of S, MMAs per output chunk
  const int chs = NK16 > 0 ? (NK16 + 1) / 2 : p.chs;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_kq);
    tma_prefetch_desc(&tm_ks);
    tma_prefetch_desc(&tm_vq);
    tma_prefetch_desc(&tm_vs);
    if (p.write_diff) tma_prefetch_desc(&tm_dq);
    if (p.write_p) tma_prefetch_desc(&tm_p);
    for (int i = 0; i < kMaxRing; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);     // tcgen05.commit (Kq k-blocks) or the store warp (staging / v_q tiles)
      mbar_init(&tile_ready[i], 4);  // the four warps of an epilogue group
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 1);
    }
    for (int i = 0; i < kMaxAcc; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 8);
    mbar_init(p_drained, 8);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------ column ring: Ks k-blocks, then Vs chunks ------------------------------
    if (lane == 0) {
      Ring r{0, 0u};
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const Item t = decode_item(p, item);
        for (int pass = 0; pass < 2; ++pass) {
          const CUtensorMap* tm = pass == 0 ? &tm_ks : &tm_vs;
          for (int kb = 0; kb < p.nchunks; ++kb) {
            mbar_wait(&empty_b[r.pos], r.phase ^ 1u);
            mbar_expect_tx(&full_b[r.pos], static_cast<uint32_t>(p.nbox) * p.load_rows * 128);
            uint8_t* dst = ring_b + static_cast<size_t>(r.pos) * p.slot_b;
            for (int h = 0; h < p.nbox; ++h)
              tma_load_4d(dst + h * p.nrow * 128, tm, &full_b[r.pos], 64 * kb, h * p.nrow, t.c, t.b);
            r.advance(p.nb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ row ring: Kq k-blocks, (P~ staging), v_q tiles ------------------------
    if (lane == 0) {
      Ring r{0, 0u};
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const Item t = decode_item(p, item);
        for (int kb = 0; kb < p.nchunks; ++kb) {
          mbar_wait(&empty_a[r.pos], r.phase ^ 1u);
          mbar_expect_tx(&full_a[r.pos], kASlot);
          tma_load_4d(ring_a + static_cast<size_t>(r.pos) * kASlot, &tm_kq, &full_a[r.pos], 64 * kb, t.m0, t.b, 0);
          r.advance(p.na);
        }
        if (p.write_p) {
          // slots lent to the compute warps as staging for the P~ tiles they write out: nothing to load
          for (int u = 0; u < p.nunits; ++u) {
            mbar_wait(&empty_a[r.pos], r.phase ^ 1u);
            mbar_arrive(&full_a[r.pos]);
            r.advance(p.na);
          }
        }
        for (int n = 0; n < p.nchunks; ++n) {
          mbar_wait(&empty_a[r.pos], r.phase ^ 1u);
          mbar_expect_tx(&full_a[r.pos], kASlot);
          tma_load_4d(ring_a + static_cast<size_t>(r.pos) * kASlot, &tm_vq, &full_a[r.pos], 64 * n, t.m0, t.b, 0);
          r.advance(p.na);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------ MMA issuer ------------------------------------------------------------
    // The whole warp walks the loop in uniform control flow (so descriptors and addresses live in uniform
    // registers) and one elected lane issues.  With the loop inside an `if (lane == 0)` the compiler had to move
    // every operand through R2UR broadcast loops: ~140 cycles per tcgen05.mma, 4x the 32 cycles a 128x64x16
    // product occupies the tensor pipe (ncu source view, profiles/r02_attn_notes.md).
    const bool leader = elect_one();
    const uint64_t desc_k = make_smem_desc_sw128(0, 16, 1024);          // K-major operand, address added below
    const uint64_t desc_mn = make_smem_desc_sw128(0, 64 * 128, 1024);   // MN-major operand (V chunk)
    const uint32_t ring_a_addr = smem_u32(ring_a), ring_b_addr = smem_u32(ring_b);
    Ring ra{0, 0u}, rb{0, 0u};
    uint32_t acc_s = 0, acc_use = 0;   // accumulator stage of the next output chunk and how often it has been used
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      // the S columns still hold the previous item's P~ until the compute warps have copied it out
      if (p.write_p && it > 0) mbar_wait(p_drained, static_cast<uint32_t>(it - 1) & 1u);
      for (int kb = 0; kb < p.nchunks; ++kb) {
        mbar_wait(&full_a[ra.pos], ra.phase);
        mbar_wait(&full_b[rb.pos], rb.phase);
        tc_fence_after();
        const uint32_t sa = ring_a_addr + static_cast<uint32_t>(ra.pos) * kASlot;
        const uint32_t sb = ring_b_addr + static_cast<uint32_t>(rb.pos) * p.slot_b;
        const uint64_t adesc = desc_k | static_cast<uint64_t>((sa & 0x3FFFFu) >> 4);
        const uint64_t bdesc1 = desc_k | static_cast<uint64_t>((sb & 0x3FFFFu) >> 4);
        const uint64_t bdesc2 = desc_k | static_cast<uint64_t>(((sb + p.n1 * 128) & 0x3FFFFu) >> 4);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {           // 16 bf16 = 32 bytes = 2 descriptor units per step
            const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
            umma_bf16(tmem_base, adesc + 2 * kk, bdesc1 + 2 * kk, p.idesc_qk1, acc);
            if (NK16 > 0 ? NK16 > 16 : p.n2 > 0)
              umma_bf16(tmem_base + p.n1, adesc + 2 * kk, bdesc2 + 2 * kk, p.idesc_qk2, acc);
          }
          umma_commit(&empty_b[rb.pos]);
          umma_commit(&empty_a[ra.pos]);
        }
        __syncwarp();
        ra.advance(p.na);
        rb.advance(p.nb);
      }
      if (leader) umma_commit(s_full);
      __syncwarp();
      if (p.write_p)
        for (int u = 0; u < p.nunits; ++u) ra.advance(p.na);
      mbar_wait(p_ready, static_cast<uint32_t>(it) & 1u);
      tc_fence_after();
      for (int n = 0; n < p.nchunks; ++n) {
        const uint32_t s = acc_s, use = acc_use;
        if (++acc_s == static_cast<uint32_t>(p.nacc)) {
          acc_s = 0;
          ++acc_use;
        }
        mbar_wait(&acc_empty[s], (use & 1u) ^ 1u);
        mbar_wait(&full_b[rb.pos], rb.phase);
        tc_fence_after();
        const uint32_t sb = ring_b_addr + static_cast<uint32_t>(rb.pos) * p.slot_b;
        const uint32_t d_tmem = tmem_base + p.o_base + 64 * s;
        if (leader) {
          // P~ chunk j sits at column p_col(j): two runs of 8-column steps; 16 k-rows of V = 2048 bytes = 128 units
          const uint64_t bdesc = desc_mn | static_cast<uint64_t>((sb & 0x3FFFFu) >> 4);
          if constexpr (NK16 > 0) {
#pragma unroll
            for (int j = 0; j < NK16; ++j) {
              const uint32_t col = j < (NK16 + 1) / 2 ? 8u * j : 16u * ((NK16 + 1) / 2) + 8u * (j - (NK16 + 1) / 2);
              umma_bf16_ts(d_tmem, tmem_base + col, bdesc + 128 * j, p.idesc_pv, j != 0 ? 1u : 0u);
            }
          } else {
            uint64_t bd = bdesc;
            uint32_t acol = tmem_base;
            int j = 0;
#pragma unroll 3
            for (; j < chs; ++j, acol += 8, bd += 128) umma_bf16_ts(d_tmem, acol, bd, p.idesc_pv, j != 0 ? 1u : 0u);
            acol = tmem_base + 16 * chs;
#pragma unroll 3
            for (; j < nk16; ++j, acol += 8, bd += 128) umma_bf16_ts(d_tmem, acol, bd, p.idesc_pv, 1u);
          }
          umma_commit(&empty_b[rb.pos]);
          umma_commit(&acc_full[s]);
        }
        __syncwarp();
        rb.advance(p.nb);
        ra.advance(p.na);                        // the v_q tile of this chunk belongs to the epilogue
      }
    }
  } else if (warp == 3) {
    // ------------------------------ TMA store issuer --------------------------------------------------------
    if (lane == 0) {
      Ring r{0, 0u};
      uint32_t ready_parity = 0;          // per slot: parity of the next tile_ready completion
      int pending = -1;                   // slot whose store may still be reading shared memory
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const Item t = decode_item(p, item);
        for (int kb = 0; kb < p.nchunks; ++kb) r.advance(p.na);      // Kq k-blocks: released by tcgen05.commit
        const int ntiles = (p.write_p ? p.nunits : 0) + p.nchunks;
        for (int i = 0; i < ntiles; ++i) {
          const bool is_p = p.write_p && i < p.nunits;
          const int col = 64 * (is_p ? i : i - (p.write_p ? p.nunits : 0));
          mbar_wait(&tile_ready[r.pos], (ready_parity >> r.pos) & 1u);
          ready_parity ^= 1u << r.pos;
          if (is_p || p.write_diff) {
            tma_store_4d(is_p ? &tm_p : &tm_dq, ring_a + static_cast<size_t>(r.pos) * kASlot, col, t.m0, t.c, t.b);
            tma_store_commit();
            if (pending >= 0) {
              tma_store_wait_read_but_one();         // the previous store has read its slot
              mbar_arrive(&empty_a[pending]);
            }
            pending = r.pos;
          } else {
            mbar_arrive(&empty_a[r.pos]);
          }
          r.advance(p.na);
        }
        if (pending >= 0) {                           // do not hold a slot across the next item's K.K^T phase
          tma_store_wait_read();
          mbar_arrive(&empty_a[pending]);
          pending = -1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ softmax, then P.V epilogue ---------------------------------------------
    const int cw = warp - 4;
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int grp = cw >> 2;                       // softmax: column half; epilogue: chunk parity
    const int row = quarter * 32 + lane;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);   // 128-byte swizzle phase of this row
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t ring_a_addr = smem_u32(ring_a);
    const int j0 = grp == 0 ? 0 : chs, j1 = grp == 0 ? chs : nk16;
    Ring ra{0, 0u};
    uint32_t g = 0, acc_s = 0, acc_use = 0;        // output chunk counter; its accumulator stage and use count
    int it = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
      const Item t = decode_item(p, item);
      const int m = t.m0 + row;
      const bool row_ok = m < p.NqT;
      const int valid = __ldg(p.cnt + t.b * p.way + t.c) * p.T;
      for (int kb = 0; kb < p.nchunks; ++kb) ra.advance(p.na);
      mbar_wait(s_full, static_cast<uint32_t>(it) & 1u);
      tc_fence_after();
      // ---- pass 1: row maximum over the valid columns ----
      float mx = -INFINITY;
      {
        uint32_t r[16], rn[16];
        if (j0 < j1) tmem_ld16(t_row + 16 * j0, rn);
        for (int j = j0; j < j1; ++j) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = rn[i];
          if (j + 1 < j1) tmem_ld16(t_row + 16 * (j + 1), rn);
          const int c0 = 16 * j;
          if (c0 + 16 <= valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < valid) mx = fmaxf(mx, __uint_as_float(r[i]));
          }
        }
      }
      xmax[grp * 128 + row] = mx;
      bar_sync(1, kComputeThreads);
      mx = fmaxf(xmax[row], xmax[128 + row]);
      // ---- pass 2: P~ = exp2((s - max) * log2(e)/sqrt(d)) -> bf16 pairs -> TMEM, in place over S ----
      const float mk = valid > 0 ? mx * p.scale_log2 : 0.f;
      float sum = 0.f;
      {
        uint32_t r[16], rn[16];
        if (j0 < j1) tmem_ld16(t_row + 16 * j0, rn);
        for (int j = j0; j < j1; ++j) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = rn[i];
          if (j + 1 < j1) tmem_ld16(t_row + 16 * (j + 1), rn);
          const int c0 = 16 * j;
          uint32_t w[8];
          if (c0 + 16 <= valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float e0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -mk));
              const float e1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -mk));
              sum += e0 + e1;
              __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float e0 = c0 + 2 * i < valid ? ex2_approx(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -mk)) : 0.f;
              const float e1 =
                  c0 + 2 * i + 1 < valid ? ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -mk)) : 0.f;
              sum += e0 + e1;
              __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
          tmem_st8(t_row + p_col(chs, j), w);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      xsum[grp * 128 + row] = sum;
      bar_sync(1, kComputeThreads);
      tc_fence_after();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      const float l = xsum[row] + xsum[128 + row];
      const float inv_l = (valid > 0 && l > 0.f) ? 1.f / l : 0.f;
      // ---- training: copy P~ out (the tiles are staged in row-ring slots lent by the producer; warp 3 stores) ----
      if (p.write_p) {
        for (int u = 0; u < p.nunits; ++u) {
          if ((u & 1) == grp) {
            mbar_wait(&full_a[ra.pos], ra.phase);
            const uint32_t rowa = ring_a_addr + static_cast<uint32_t>(ra.pos) * kASlot + row * 128;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = 4 * u + jj;
              if (j < nk16) {
                uint32_t w[8];
                tmem_ld8(t_row + p_col(chs, j), w);
                tmem_ld_wait();
                sts128(rowa + (((2 * jj) ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
                sts128(rowa + (((2 * jj + 1) ^ sw) << 4), make_uint4(w[4], w[5], w[6], w[7]));
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tile_ready[ra.pos]);
          }
          ra.advance(p.na);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_drained);
      }
      // ---- epilogue of the output chunks with this group's parity ----
      float rsum = 0.f, rdot = 0.f, rsum_b = 0.f, rdot_b = 0.f;
      for (int n = 0; n < p.nchunks; ++n, ++g) {
        if (static_cast<int>(g & 1u) == grp) {
          mbar_wait(&full_a[ra.pos], ra.phase);
          mbar_wait(&acc_full[acc_s], acc_use & 1u);
          tc_fence_after();
          const uint32_t rowa = ring_a_addr + static_cast<uint32_t>(ra.pos) * kASlot + row * 128;
          const uint32_t t_acc = t_row + p.o_base + 64 * acc_s;
          // everything this chunk needs is requested up front: 4 TMEM loads and the 8 v_q pieces of this row
          uint32_t acc[4][16];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) tmem_ld16(t_acc + 16 * cc, acc[cc]);
          uint4 q[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) q[i] = lds128(rowa + ((static_cast<uint32_t>(i) ^ sw) << 4));
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc_s]);           // the accumulator stage is free again
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint4& qq = q[2 * cc + hh];
              uint32_t aw[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
                const float o0 = __uint_as_float(acc[cc][8 * hh + 2 * i]) * inv_l;
                const float o1 = __uint_as_float(acc[cc][8 * hh + 2 * i + 1]) * inv_l;
                const float d0 = a.x - o0, d1 = a.y - o1;
                // two independent partial sums per moment: 64-long dependent FMA chains were the longest in the chunk
                rsum = fmaf(d0, d0, rsum);
                rsum_b = fmaf(d1, d1, rsum_b);
                rdot = fmaf(d0, o0, rdot);
                rdot_b = fmaf(d1, o1, rdot_b);
                __nv_bfloat162 h = __floats2bfloat162_rn(d0, d1);
                aw[i] = *reinterpret_cast<uint32_t*>(&h);
              }
              if (p.write_diff)                                     // in place: exactly the 16 bytes read above
                sts128(rowa + ((static_cast<uint32_t>(2 * cc + hh) ^ sw) << 4), make_uint4(aw[0], aw[1], aw[2], aw[3]));
            }
          }
          if (p.write_diff) fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tile_ready[ra.pos]);          // warp 3 stores the tile and frees the slot
        }
        ra.advance(p.na);
        if (++acc_s == static_cast<uint32_t>(p.nacc)) {
          acc_s = 0;
          ++acc_use;
        }
      }
      if (row_ok) {
        const int64_t o = (static_cast<int64_t>(t.b) * p.way + t.c) * p.NqT + m;
        atomicAdd(p.rowred + o, rsum + rsum_b);
        if (p.rowdot != nullptr) atomicAdd(p.rowdot + o, rdot + rdot_b);
        if (p.linv != nullptr && grp == 0) p.linv[o] = inv_l;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// LMKD_TRX_FUSED_ATTN=0 keeps the round-1 pipeline (scores GEMM, softmax kernel, P.V GEMM) for A/B measurements
static const bool g_fused_attn = [] {
  const char* e = getenv("LMKD_TRX_FUSED_ATTN");
  return !(e && e[0] == '0');
}();

bool trx_attn_fused_fits(const TrxDims& s) {
  return g_fused_attn && s.d % 64 == 0 && s.KTp % 16 == 0 && s.KTp >= 16 && s.KTp <= 384;
}

int trx_attn_fwd(const TrxAttnFwd& a, const TrxDims& s, cudaStream_t st) {
  LMKD_CHECK(trx_attn_fused_fits(s), "trx_attn_fwd: shape does not fit the fused kernel (KTp %d, d %d)", s.KTp, s.d);
  LMKD_CHECK(a.kq && a.vq && a.ks && a.vs && a.cnt && a.rowred, "trx_attn_fwd: null pointer");
  AttnParams p{};
  p.way = s.way; p.NqT = s.NqT; p.KTp = s.KTp; p.T = s.T;
  p.tiles_m = static_cast<int>(ceil_div(s.NqT, 128));
  const int64_t items = static_cast<int64_t>(s.B) * s.way * p.tiles_m;
  LMKD_CHECK(items < (1ll << 31), "trx_attn_fwd: too many tiles");
  p.num_items = static_cast<int>(items);
  if (s.KTp <= 256) {
    p.nbox = 1; p.nrow = s.KTp; p.n1 = s.KTp; p.n2 = 0;
  } else {
    p.nbox = 2; p.nrow = static_cast<int>(round_up(s.KTp / 2, 16)); p.n1 = p.nrow; p.n2 = s.KTp - p.n1;
  }
  p.nchunks = s.d / 64;
  p.nk16 = s.KTp / 16;
  p.nch = s.KTp / 16;
  p.chs = (p.nch + 1) / 2;
  p.nunits = static_cast<int>(ceil_div(s.KTp, 64));
  p.o_base = static_cast<int>(round_up(s.KTp, 64));
  p.nacc = (static_cast<int>(kTmemCols) - p.o_base) / 64;
  if (p.nacc > kMaxAcc) p.nacc = kMaxAcc;
  p.slot_b = static_cast<uint32_t>(p.nbox) * p.nrow * 128;
  // LMKD_ATTN_EXPERIMENT_HALFB=1 (measurement only, results are WRONG): fetch half of the Ks / Vs rows of every slot,
  // to see what the kernel would gain if a CTA pair shared those loads
  static const bool half_b = [] {
    const char* e = getenv("LMKD_ATTN_EXPERIMENT_HALFB");
    return e && e[0] == '1';
  }();
  p.load_rows = half_b ? p.nrow / 2 : p.nrow;
  const int avail = 227 * 1024 - 1024 - kTailBytes;
  int depth = avail / static_cast<int>(p.slot_b + kASlot);
  if (depth > kMaxRing) depth = kMaxRing;
  LMKD_CHECK(depth >= 2 && p.nacc >= 2, "trx_attn_fwd: not enough shared / tensor memory (KTp %d)", s.KTp);
  p.nb = depth;
  // row tiles live longer than column slots (epilogue + store): they get whatever is left
  p.na = (avail - depth * static_cast<int>(p.slot_b)) / static_cast<int>(kASlot);
  if (p.na > kMaxRing) p.na = kMaxRing;
  p.idesc_qk1 = make_idesc_bf16(128, p.n1, 0, 0);
  p.idesc_qk2 = p.n2 > 0 ? make_idesc_bf16(128, p.n2, 0, 0) : 0u;
  p.idesc_pv = make_idesc_bf16(128, 64, 0, 1);
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(s.d));
  p.write_diff = a.dq != nullptr;
  p.write_p = a.patt != nullptr;
  LMKD_CHECK(!p.write_p || a.linv != nullptr, "trx_attn_fwd: patt needs linv");
  p.cnt = a.cnt; p.rowred = a.rowred; p.rowdot = a.rowdot; p.linv = a.linv;

  const uint64_t d2 = static_cast<uint64_t>(s.d) * 2;
  const int64_t pitch = static_cast<int64_t>(s.way) * s.KTp;
  CUtensorMap m_kq, m_ks, m_vq, m_vs, m_dq, m_p;
  auto rows_map = [&](CUtensorMap* out, const void* base, const char* what) {   // [B][NqT][d]
    TmapSpec t;
    t.base = base;
    t.dims[0] = s.d; t.dims[1] = s.NqT; t.dims[2] = s.B; t.dims[3] = 1;
    t.strides[0] = d2; t.strides[1] = d2 * s.NqT; t.strides[2] = d2 * s.NqT * s.B;
    t.box[0] = 64; t.box[1] = 128;
    return encode_tmap(out, t, what);
  };
  auto cls_map = [&](CUtensorMap* out, const void* base, const char* what) {    // [B][way][KTp][d]
    TmapSpec t;
    t.base = base;
    t.dims[0] = s.d; t.dims[1] = s.KTp; t.dims[2] = s.way; t.dims[3] = s.B;
    t.strides[0] = d2; t.strides[1] = d2 * s.KTp; t.strides[2] = d2 * s.KTp * s.way;
    t.box[0] = 64; t.box[1] = static_cast<uint32_t>(p.load_rows);
    return encode_tmap(out, t, what);
  };
  if (int rc = rows_map(&m_kq, a.kq, "attn kq")) return rc;
  if (int rc = rows_map(&m_vq, a.vq, "attn vq")) return rc;
  if (int rc = cls_map(&m_ks, a.ks, "attn ks")) return rc;
  if (int rc = cls_map(&m_vs, a.vs, "attn vs")) return rc;
  m_dq = m_kq;
  m_p = m_kq;
  if (p.write_diff) {                                                            // [B][way][NqT][d]
    TmapSpec t;
    t.base = a.dq;
    t.dims[0] = s.d; t.dims[1] = s.NqT; t.dims[2] = s.way; t.dims[3] = s.B;
    t.strides[0] = d2; t.strides[1] = d2 * s.NqT; t.strides[2] = d2 * s.NqT * s.way;
    t.box[0] = 64; t.box[1] = 128;
    t.promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (int rc = encode_tmap(&m_dq, t, "attn dq")) return rc;
  }
  if (p.write_p) {                                                               // [B][NqT][way][KTp]
    TmapSpec t;
    t.base = a.patt;
    t.dims[0] = s.KTp; t.dims[1] = s.NqT; t.dims[2] = s.way; t.dims[3] = s.B;
    t.strides[0] = static_cast<uint64_t>(pitch) * 2; t.strides[1] = static_cast<uint64_t>(s.KTp) * 2;
    t.strides[2] = static_cast<uint64_t>(pitch) * 2 * s.NqT;
    t.box[0] = 64; t.box[1] = 128;
    t.promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (int rc = encode_tmap(&m_p, t, "attn patt")) return rc;
  }
  const size_t smem = 1024 + static_cast<size_t>(p.nb) * p.slot_b + static_cast<size_t>(p.na) * kASlot + kTailBytes;
  auto kern = p.nk16 == 9 ? trx_attn_fwd_kernel<9> : p.nk16 == 18 ? trx_attn_fwd_kernel<18> : trx_attn_fwd_kernel<0>;
  if (int rc = ensure_max_dynamic_smem(reinterpret_cast<const void*>(kern), 227 * 1024)) return rc;
  KernelTimingScope timing(TIME_TENSOR, st, 4.0 * s.B * s.way * static_cast<double>(s.NqT) * s.KTp * s.d);
  if (int rc = timing.begin()) return rc;
  const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
  // >= 120 KB of dynamic shared memory keeps one CTA per SM (each CTA allocates all 512 TMEM columns)
  kern<<<grid, kThreads, smem < 120 * 1024 ? 120 * 1024 : smem, st>>>(m_kq, m_ks, m_vq, m_vs, m_dq, m_p, p);
  LMKD_LAUNCH_CHECK("trx_attn_fwd_kernel");
  return timing.end();
}

}  // namespace lmkd
