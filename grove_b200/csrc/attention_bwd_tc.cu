// Key side of the GLOBAL attention backward on tcgen05 / TMEM / TMA (image_encoder.py:301-326, 420-458 under autograd; the blocks are
// frozen, so only d(qkv) is produced).  Replaces the warp-level mma.sync kernel attn_bwd_kv_kernel for head dim 64 on 32x32 / 64x64 grids.
//
//   S[q,k] = scale q.k + rel_h[q, kh] + rel_w[q, kw],  P = exp(S - lse_q),  dP = dO V^T,  dS = P (dP - D_q),  dV = P^T dO,  dK = scale dS^T Q
//
// One CTA owns 128 keys of one (frame, head) and sweeps the queries in blocks of 128, everything in the TRANSPOSED orientation (TMEM lane =
// key) so that P^T and dS^T are directly the A operands of the two accumulating products:
//   MMA1   S^T = K_own Q_i^T,  dP^T = V_own dO_i^T          (SS-mode UMMA 128x128x64, both into tensor memory)
//   warps  p = exp2(c S^T + (rel_w[q,kw] + rel_h[q,kh]) log2e - lse_q),  ds = p (dP^T - D_q)   -> bf16 P^T, dS^T written back to TMEM
//   MMA2   dV += P^T dO_i,  dK += dS^T Q_i                  (TS-mode UMMA 128x64x128, B = the same Q_i / dO_i tiles read MN-major)
// The per-query quantities come from tables the query-side pass already built (attention_bwd.cu): rel[token, head, 2G] (fp32 bias rows,
// relpos_kernel), the forward's log-sum-exp and D = rowsum(dO * O), packed into aux[token, head, 4]; TMA boxes deliver the 128 x G rel_w
// slab, the 128 x 4 rel_h slab of this key block's grid rows and the aux rows of every query block.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-9 elementwise (two threads per key row, 64 query columns each).
// TMEM: S^T | dP^T | P^T | dS^T | dV | dK = 128 + 128 + 64 + 64 + 64 + 64 columns.  S^T/dP^T of block i+1 are issued as soon as block i's
// values are in registers, so MMA1 and MMA2 run under the elementwise phase of the neighbouring blocks.
#include <cuda.h>

#include "common.cuh"
#include "grove_b200.h"
#include "tmem_ldst.cuh"

namespace grove {

// qkv / dO: 128-row boxes of the 64-wide main part; *_x: the 16-wide tail of an 80-wide head (SWIZZLE_32B); *_b: the same two with
// BQ-row boxes (the key-side kernel's query blocks)
struct BwdKvTmaps { CUtensorMap qkv, dO, relw, relh, aux, qkv_x, dO_x, qkv_b, dO_b, qkv_bx, dO_bx, rh, rw, rh_x, rw_x; };   // rh / rw: the rel-pos tables (query side)

constexpr int kBwdThreads = 320;

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int G, int HD>
struct BwdKvCfg {
  static constexpr bool kX = HD > 64;
  // 64-wide heads sweep the queries in blocks of 128; 80-wide heads in blocks of 64, because S^T | dP^T | P^T | dS^T of a 128-query block
  // (384 columns) plus two 80-column accumulators would not fit the 512 columns of tensor memory
  static constexpr int BQ = kX ? 64 : 128;
  static constexpr int TS = 16384 + (kX ? 4096 : 0);     // own K / V: [128 x 64] bf16 (SWIZZLE_128B) [+ 128 x 16 tail, SWIZZLE_32B]
  static constexpr int TQM = BQ * 128, TQ = TQM + (kX ? BQ * 32 : 0);   // Q_i / dO_i of a query block
  static constexpr int kRelW = BQ * G * 4;               // rel_w slab of a query block: [BQ q][G] fp32
  static constexpr int kSmall = BQ * 4 * 4;              // rel_h slab / aux rows / combined rows: [BQ q][4] fp32
  static constexpr int kStage = 2 * TQ + kRelW + 4 * kSmall;   // Q_i | dO_i | rel_w | rel_h | aux | comb_h | D
  static constexpr int kTx = 2 * TQ + kRelW + 2 * kSmall;       // bytes the TMA delivers per stage
  static constexpr int kSmem = 2 * TS + 2 * kStage + 1024 /*align*/ + 256 /*barriers*/;
};

template <int G, int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kv_tc_kernel(const __grid_constant__ BwdKvTmaps tm, __nv_bfloat16* __restrict__ dqkv, int heads) {
  using C = BwdKvCfg<G, HD>;
  constexpr bool kX = C::kX;
  constexpr int TS = C::TS, TQ = C::TQ, TQM = C::TQM, BQ = C::BQ, N = G * G, NB = N / BQ, RPT = 128 / G;   // RPT: grid rows covered by one 128-key block
  constexpr int NCC = BQ / 64;                           // 32-column chunks per elementwise thread and block
  constexpr float kScale = HD == 64 ? 0.125f : 0.11180339887498949f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = s0, sVo = s0 + TS;
  auto sStage = [&](int s) { return s0 + 2 * TS + (uint32_t)s * C::kStage; };
  const uint32_t bar0 = s0 + 2 * TS + 2 * C::kStage;
  uint8_t* smem_al = smem_raw + (s0 - smem_u32(smem_raw));
  enum { OWN_FULL = 0, STAGE_FULL, STAGE_EMPTY = STAGE_FULL + 2, SD_FULL = STAGE_EMPTY + 2, SD_EMPTY, PD_FULL, PD_EMPTY, ACC_FULL, NUM_BARS };
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_slot = bar0 + 8u * NUM_BARS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = blockIdx.x, h = blockIdx.y, f = blockIdx.z;
  const int D = heads * HD;
  const int tok0 = f * N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.qkv);
    mbar_init(bar(OWN_FULL), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(STAGE_FULL + i), 1); mbar_init(bar(STAGE_EMPTY + i), 9); }
    mbar_init(bar(SD_FULL), 1); mbar_init(bar(SD_EMPTY), 8); mbar_init(bar(PD_FULL), 8); mbar_init(bar(PD_EMPTY), 1); mbar_init(bar(ACC_FULL), 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tDP = tmem_base + BQ, tP = tmem_base + 2 * BQ, tDS = tP + BQ / 2, tDV = tmem_base + 3 * BQ, tDK = tDV + HD;
  static_assert(3 * BQ + 2 * HD <= 512, "tensor memory budget");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(bar(OWN_FULL), 2 * TS);
      tma_load_2d(sK, &tm.qkv, bar(OWN_FULL), D + h * HD, tok0 + jb * 128);
      tma_load_2d(sVo, &tm.qkv, bar(OWN_FULL), 2 * D + h * HD, tok0 + jb * 128);
      if (kX) {
        tma_load_2d(sK + 16384, &tm.qkv_x, bar(OWN_FULL), D + h * HD + 64, tok0 + jb * 128);
        tma_load_2d(sVo + 16384, &tm.qkv_x, bar(OWN_FULL), 2 * D + h * HD + 64, tok0 + jb * 128);
      }
      for (int i = 0; i < NB; ++i) {
        const int s = i & 1;
        mbar_wait(bar(STAGE_EMPTY + s), ((i >> 1) & 1u) ^ 1u);
        const uint32_t st = sStage(s);
        mbar_expect_tx(bar(STAGE_FULL + s), C::kTx);
        const int tq = tok0 + i * BQ;
        tma_load_2d(st, &tm.qkv_b, bar(STAGE_FULL + s), h * HD, tq);                     // Q_i
        tma_load_2d(st + TQ, &tm.dO_b, bar(STAGE_FULL + s), h * HD, tq);                 // dO_i
        if (kX) {
          tma_load_2d(st + TQM, &tm.qkv_bx, bar(STAGE_FULL + s), h * HD + 64, tq);
          tma_load_2d(st + TQ + TQM, &tm.dO_bx, bar(STAGE_FULL + s), h * HD + 64, tq);
        }
        tma_load_3d(st + 2 * TQ, &tm.relw, bar(STAGE_FULL + s), G, h, tq);               // rel_w[q, 0:G]   (columns G..2G of the bias row)
        // rel_h[q, kh of this key block]: 4 floats from a 16-byte aligned column (TMA box origins must be; jb * RPT is odd pairs for G = 64)
        tma_load_3d(st + 2 * TQ + C::kRelW, &tm.relh, bar(STAGE_FULL + s), (jb * RPT) & ~3, h, tq);
        tma_load_3d(st + 2 * TQ + C::kRelW + C::kSmall, &tm.aux, bar(STAGE_FULL + s), 0, h, tq);       // (lse, D, -, -)
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, BQ);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);   // B is MN-major (the [q][d] tile read along q)
    constexpr uint32_t idesc_ox = umma_idesc_bf16(128, 16) | (1u << 16);
    auto mma2 = [&](int i) {           // dV += P^T dO_i, dK += dS^T Q_i
      const uint32_t st = sStage(i & 1);
      mbar_wait(bar(PD_FULL), i & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BQ / 16; ++kk) {
          tc_mma_f16_ts(tDV, tP + kk * 8, umma_desc_sw128(st + TQ + kk * 2048), idesc_o, (i | kk) != 0);
          if (kX) tc_mma_f16_ts(tDV + 64, tP + kk * 8, umma_desc_sw32(st + TQ + TQM + kk * 512), idesc_ox, (i | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < BQ / 16; ++kk) {
          tc_mma_f16_ts(tDK, tDS + kk * 8, umma_desc_sw128(st + kk * 2048), idesc_o, (i | kk) != 0);
          if (kX) tc_mma_f16_ts(tDK + 64, tDS + kk * 8, umma_desc_sw32(st + TQM + kk * 512), idesc_ox, (i | kk) != 0);
        }
        tc_commit(bar(PD_EMPTY));
        tc_commit(bar(STAGE_EMPTY + (i & 1)));
        if (i == NB - 1) tc_commit(bar(ACC_FULL));
      }
      __syncwarp();
    };
    mbar_wait(bar(OWN_FULL), 0);
#pragma unroll 1
    for (int i = 0; i < NB; ++i) {
      const int s = i & 1;
      const uint32_t st = sStage(s);
      mbar_wait(bar(STAGE_FULL + s), (i >> 1) & 1u);
      if (i > 0) mbar_wait(bar(SD_EMPTY), (i - 1) & 1u);       // block i-1's S^T / dP^T are in registers
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tS, umma_desc_sw128(sK + k * 32), umma_desc_sw128(st + k * 32), idesc_s, k != 0);
        if (kX) tc_mma_f16(tS, umma_desc_sw32(sK + 16384), umma_desc_sw32(st + TQM), idesc_s, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tDP, umma_desc_sw128(sVo + k * 32), umma_desc_sw128(st + TQ + k * 32), idesc_s, k != 0);
        if (kX) tc_mma_f16(tDP, umma_desc_sw32(sVo + 16384), umma_desc_sw32(st + TQ + TQM), idesc_s, 1);
        tc_commit(bar(SD_FULL));
      }
      __syncwarp();
      if (i > 0) mma2(i - 1);
    }
    mma2(NB - 1);
  } else {
    // ===================== elementwise warps: two threads per key row =====================
    const int quad = warp & 3;
    const int hs = (warp - 2) >> 2;                      // this thread owns query chunks {hs, hs + 2} (32 columns each) of every block
    const int row = quad * 32 + lane;                    // key row inside the block == TMEM lane
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    const int et = (warp - 2) * 32 + lane;               // 0..255
    const int kw = row % G, khl = row / G + ((jb * RPT) & 3);   // key column; key row relative to the aligned rel_h box
    constexpr float kL2e = 1.4426950408889634f;
    const float c_l2 = kScale * kL2e;                    // hd^-0.5 * log2(e)
#pragma unroll 1
    for (int i = 0; i < NB; ++i) {
      const int s = i & 1;
      float* stg = reinterpret_cast<float*>(smem_al + (sStage(s) - s0) + 2 * TQ);
      const float* relw_s = stg;                                         // [BQ][G]
      const float* relh_s = stg + BQ * G;                                // [BQ][4]
      const float* aux_s = relh_s + BQ * 4;                              // [BQ][4] = lse, D, -, -
      float* comb_s = stg + BQ * G + 2 * BQ * 4;                         // [BQ][4] = rel_h * log2e - lse for the block's grid rows
      float* dsum_s = comb_s + BQ * 4;                                   // [BQ]
      mbar_wait(bar(STAGE_FULL + s), (i >> 1) & 1u);
      if (et < BQ) {
        const float4 rh = *reinterpret_cast<const float4*>(relh_s + et * 4);
        const float2 ax = *reinterpret_cast<const float2*>(aux_s + et * 4);
        *reinterpret_cast<float4*>(comb_s + et * 4) = make_float4(fmaf(rh.x, kL2e, -ax.x), fmaf(rh.y, kL2e, -ax.x), fmaf(rh.z, kL2e, -ax.x), fmaf(rh.w, kL2e, -ax.x));
        dsum_s[et] = ax.y;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(bar(SD_FULL), i & 1u);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) {
        const int c = 2 * cc + hs;
        uint32_t rs[32], rd[32];
        tmem_ld_x32(tS + c * 32 + tlane, rs);
        tmem_ld_x32(tDP + c * 32 + tlane, rd);
        tmem_ld_wait();
        if (cc == NCC - 1) {                             // all chunks of S^T / dP^T are in registers: MMA1 of the next block may overwrite them
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(SD_EMPTY));
        }
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int q = c * 32 + j;
          const float v0 = fmaf(__uint_as_float(rs[j]), c_l2, fmaf(relw_s[q * G + kw], kL2e, comb_s[q * 4 + khl]));
          const float v1 = fmaf(__uint_as_float(rs[j + 1]), c_l2, fmaf(relw_s[(q + 1) * G + kw], kL2e, comb_s[(q + 1) * 4 + khl]));
          float p0, p1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(v0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(v1));
          pp[j >> 1] = pack_bf16(p0, p1);
          pd[j >> 1] = pack_bf16(p0 * (__uint_as_float(rd[j]) - dsum_s[q]), p1 * (__uint_as_float(rd[j + 1]) - dsum_s[q + 1]));
        }
        if (cc == 0 && i > 0) { mbar_wait(bar(PD_EMPTY), (i - 1) & 1u); tc_fence_after(); }   // MMA2 of block i-1 has read P^T / dS^T
        tmem_st_x16(tP + c * 16 + tlane, pp);
        tmem_st_x16(tDS + c * 16 + tlane, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(PD_FULL)); mbar_arrive(bar(STAGE_EMPTY + s)); }
    }
    // ---- epilogue: dK (x scale), dV -> bf16 into the k / v slots of dqkv; each thread of a row stores 32 of the 64 head dims
    mbar_wait(bar(ACC_FULL), 0);
    tc_fence_after();
    __nv_bfloat16* orow = dqkv + ((size_t)tok0 + jb * 128 + row) * (3 * D) + h * HD;
#pragma unroll
    for (int which = 0; which < 2; ++which) {            // 0: dK, 1: dV
      uint32_t r[32], rx[8];
      tmem_ld_x32((which == 0 ? tDK : tDV) + hs * 32 + tlane, r);
      if (kX) tmem_ld_x8((which == 0 ? tDK : tDV) + 64 + hs * 8 + tlane, rx);
      tmem_ld_wait();
      const float sc = which == 0 ? kScale : 1.0f;
      __nv_bfloat16* o = orow + (which == 0 ? D : 2 * D);
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(o + hs * 32 + j) =
            make_uint4(pack_bf16(__uint_as_float(r[j]) * sc, __uint_as_float(r[j + 1]) * sc), pack_bf16(__uint_as_float(r[j + 2]) * sc, __uint_as_float(r[j + 3]) * sc),
                       pack_bf16(__uint_as_float(r[j + 4]) * sc, __uint_as_float(r[j + 5]) * sc), pack_bf16(__uint_as_float(r[j + 6]) * sc, __uint_as_float(r[j + 7]) * sc));
      if (kX)
        *reinterpret_cast<uint4*>(o + 64 + hs * 8) =
            make_uint4(pack_bf16(__uint_as_float(rx[0]) * sc, __uint_as_float(rx[1]) * sc), pack_bf16(__uint_as_float(rx[2]) * sc, __uint_as_float(rx[3]) * sc),
                       pack_bf16(__uint_as_float(rx[4]) * sc, __uint_as_float(rx[5]) * sc), pack_bf16(__uint_as_float(rx[6]) * sc, __uint_as_float(rx[7]) * sc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// =====================================================================================================================
// Query side of the global attention backward on tcgen05: one CTA owns 128 queries of one (frame, head) and sweeps the keys in blocks of 128
//   prologue  T_h = Q_own Rh^T, T_w = Q_own Rw^T (UMMA 128 x 2G x HD, the whole rel-pos tables as B) -> the bias rows of the 128 queries:
//             rel_w[q, kw] = T_w[q, qw-kw+G-1] into registers, rel_h into shared memory (x log2 e), and both, coalesced, into the global
//             rel [M, heads, 2G] table the key-side kernel reads
//   MMA1   S = Q_own K_j^T,  dP = dO_own V_j^T               (SS-mode UMMA 128x128xHD)
//   warps  ds = exp2(c S + (rel_w + rel_h) log2e - lse) (dP - D)   -> scale*ds (bf16) into tensor memory; the rel-pos bias cotangents
//          A_w[q, kw] += sum_kh ds (registers), A_h[q, kh] = sum_kw ds (one value per thread and chunk, combined through shared memory)
//   MMA2   dQ += (scale dS) K_j                               (TS-mode UMMA 128xHDx128, K_j read MN-major)
//   epilogue  dQ += dT_h Rh + dT_w Rw (SS-mode UMMA 128 x HD x 2G): dT[q, qh-kh+G-1] = A_h[q, kh] is the cotangent of T_h (likewise T_w),
//             scattered as bf16 into K-major tiles over the retired K / V stages; the tables are re-fetched over the Q / dO tiles.
//             d q leaves as bf16 straight into the q slot of dqkv.
// (Round-2 history: the bias rows and the table back-projection were two CUDA-core kernels, relpos_kernel<0/1>, together 1.5 ms per ViT-B
// layer next to 2.7 ms of attention kernels.)
// Two threads per query row: thread hs owns the key chunks {hs, hs+2} of every block, i.e. always the same 32 key columns kw.
// =====================================================================================================================
template <int G, int HD>
struct BwdQCfg {
  static constexpr bool kX = HD > 64;
  static constexpr int TS = 16384 + (kX ? 4096 : 0);      // [128 x 64] bf16 (SWIZZLE_128B) [+ 128 x 16 tail, SWIZZLE_32B]
  static constexpr int kTabMain = 2 * G * 128;            // a rel-pos table as an operand tile: 2G rows (2G-1 real) x HD
  static constexpr int kTab = kTabMain + (kX ? 2 * G * 32 : 0);
  static constexpr int kRelH = G * 128 * 4;               // rel_h [G][128] fp32
  static constexpr int kAh = 2 * G * 128 * 4;             // A_h partial sums [2][G][128] fp32; the prologue's [G][129] transpose buffer
  static constexpr int kDT = (2 * G / 64) * 16384;        // one cotangent tile dT [128][2G] bf16, K-major in 64-column sub-tiles
  static constexpr int kSmem = 2 * TS + 4 * TS + kRelH + kAh + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kTab <= TS && 2 * kDT + (G == 32 ? 32768 : 0) <= 4 * TS && G * 129 * 4 <= kAh, "regions reused by the prologue / epilogue");
};

template <int G, int HD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_q_tc_kernel(const __grid_constant__ BwdKvTmaps tm, float* __restrict__ rel_out, const float* __restrict__ lse, const float* __restrict__ dsum,
                     __nv_bfloat16* __restrict__ dqkv, int heads) {
  using C = BwdQCfg<G, HD>;
  constexpr bool kX = C::kX;
  constexpr int TS = C::TS, N = G * G, NB = N / 128;
  constexpr float kScale = HD == 64 ? 0.125f : 0.11180339887498949f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = s0, sDO = s0 + TS;
  auto sStage = [&](int s) { return s0 + 2 * TS + (uint32_t)s * (2 * TS); };     // K_j | V_j
  const uint32_t sRelH = s0 + 6 * TS, sAh = sRelH + C::kRelH, bar0 = sAh + C::kAh;
  const uint32_t sDTh = s0 + 2 * TS, sDTw = sDTh + C::kDT;                         // epilogue: over the K / V stages
  uint8_t* smem_al = smem_raw + (s0 - smem_u32(smem_raw));
  float* relh_s = reinterpret_cast<float*>(smem_al + (sRelH - s0));       // [G][128], x log2 e
  float* ah_s = reinterpret_cast<float*>(smem_al + (sAh - s0));           // [2][G][128]
  float* stg_s = ah_s;                                                    // prologue: [G][129] transpose buffer
  uint8_t* dt_b = smem_al + 2 * TS;                                       // dT_h | dT_w
  float* xch_s = reinterpret_cast<float*>(smem_al + 2 * TS + 2 * C::kDT); // G = 32 epilogue exchange: [2][32][128]
  enum { OWN_FULL = 0, STAGE_FULL, STAGE_EMPTY = STAGE_FULL + 2, SD_FULL = STAGE_EMPTY + 2, SD_EMPTY, PD_FULL, PD_EMPTY, ACC_FULL, TAB_FULL, TAB_FREE,
         T_FULL, T_READ, DT_FULL, ACC2_FULL, NUM_BARS };
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_slot = bar0 + 8u * NUM_BARS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ib = blockIdx.x, h = blockIdx.y, f = blockIdx.z;
  const int D = heads * HD;
  const int tok0 = f * N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.qkv);
    mbar_init(bar(OWN_FULL), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar(STAGE_FULL + i), 1); mbar_init(bar(STAGE_EMPTY + i), 1); }
    mbar_init(bar(SD_FULL), 1); mbar_init(bar(SD_EMPTY), 8); mbar_init(bar(PD_FULL), 8); mbar_init(bar(PD_EMPTY), 1); mbar_init(bar(ACC_FULL), 1);
    mbar_init(bar(TAB_FULL), 1); mbar_init(bar(TAB_FREE), 1); mbar_init(bar(T_FULL), 1); mbar_init(bar(T_READ), 8); mbar_init(bar(DT_FULL), 8);
    mbar_init(bar(ACC2_FULL), 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDS = tmem_base + 256, tDQ = tmem_base + 320;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      auto load_tables = [&](uint32_t dst_h, uint32_t dst_w) {           // rows 2G-1.. of the boxes are out of bounds: zero-filled
        mbar_expect_tx(bar(TAB_FULL), 2 * C::kTab);
        tma_load_2d(dst_h, &tm.rh, bar(TAB_FULL), 0, 0);
        tma_load_2d(dst_w, &tm.rw, bar(TAB_FULL), 0, 0);
        if (kX) {
          tma_load_2d(dst_h + C::kTabMain, &tm.rh_x, bar(TAB_FULL), 64, 0);
          tma_load_2d(dst_w + C::kTabMain, &tm.rw_x, bar(TAB_FULL), 64, 0);
        }
      };
      mbar_expect_tx(bar(OWN_FULL), 2 * TS);
      tma_load_2d(sQ, &tm.qkv, bar(OWN_FULL), h * HD, tok0 + ib * 128);
      tma_load_2d(sDO, &tm.dO, bar(OWN_FULL), h * HD, tok0 + ib * 128);
      if (kX) {
        tma_load_2d(sQ + 16384, &tm.qkv_x, bar(OWN_FULL), h * HD + 64, tok0 + ib * 128);
        tma_load_2d(sDO + 16384, &tm.dO_x, bar(OWN_FULL), h * HD + 64, tok0 + ib * 128);
      }
      load_tables(sStage(1), sStage(1) + TS);                            // the prologue's tables sit in stage 1 until T_h / T_w have retired
      for (int j = 0; j < NB; ++j) {
        const int s = j & 1;
        if (j == 1) mbar_wait(bar(TAB_FREE), 0);
        mbar_wait(bar(STAGE_EMPTY + s), ((j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(bar(STAGE_FULL + s), 2 * TS);
        tma_load_2d(sStage(s), &tm.qkv, bar(STAGE_FULL + s), D + h * HD, tok0 + j * 128);           // K_j
        tma_load_2d(sStage(s) + TS, &tm.qkv, bar(STAGE_FULL + s), 2 * D + h * HD, tok0 + j * 128);  // V_j
        if (kX) {
          tma_load_2d(sStage(s) + 16384, &tm.qkv_x, bar(STAGE_FULL + s), D + h * HD + 64, tok0 + j * 128);
          tma_load_2d(sStage(s) + TS + 16384, &tm.qkv_x, bar(STAGE_FULL + s), 2 * D + h * HD + 64, tok0 + j * 128);
        }
      }
      mbar_wait(bar(ACC_FULL), 0);                                       // Q_own / dO_own have served their last product
      load_tables(sQ, sDO);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128), idesc_t = umma_idesc_bf16(128, 2 * G);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);   // B read MN-major
    constexpr uint32_t idesc_ox = umma_idesc_bf16(128, 16) | (1u << 16);
    auto mma2 = [&](int j) {           // dQ += (scale dS) K_j
      const uint32_t st = sStage(j & 1);
      mbar_wait(bar(PD_FULL), j & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          tc_mma_f16_ts(tDQ, tDS + kk * 8, umma_desc_sw128(st + kk * 2048), idesc_o, (j | kk) != 0);
          if (kX) tc_mma_f16_ts(tDQ + 64, tDS + kk * 8, umma_desc_sw32(st + 16384 + kk * 512), idesc_ox, (j | kk) != 0);
        }
        tc_commit(bar(PD_EMPTY));
        tc_commit(bar(STAGE_EMPTY + (j & 1)));
        if (j == NB - 1) tc_commit(bar(ACC_FULL));
      }
      __syncwarp();
    };
    mbar_wait(bar(OWN_FULL), 0);
    mbar_wait(bar(TAB_FULL), 0);
    tc_fence_after();
    if (elect_one()) {                 // T_h -> the S columns, T_w -> the dP columns
      const uint32_t th = sStage(1), tw = sStage(1) + TS;
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16(tS, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(th + k * 32), idesc_t, k != 0);
      if (kX) tc_mma_f16(tS, umma_desc_sw32(sQ + 16384), umma_desc_sw32(th + C::kTabMain), idesc_t, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16(tDP, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(tw + k * 32), idesc_t, k != 0);
      if (kX) tc_mma_f16(tDP, umma_desc_sw32(sQ + 16384), umma_desc_sw32(tw + C::kTabMain), idesc_t, 1);
      tc_commit(bar(T_FULL));
      tc_commit(bar(TAB_FREE));
    }
    __syncwarp();
    mbar_wait(bar(T_READ), 0);
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
      const int s = j & 1;
      const uint32_t st = sStage(s);
      mbar_wait(bar(STAGE_FULL + s), (j >> 1) & 1u);
      if (j > 0) mbar_wait(bar(SD_EMPTY), (j - 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tS, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(st + k * 32), idesc_s, k != 0);
        if (kX) tc_mma_f16(tS, umma_desc_sw32(sQ + 16384), umma_desc_sw32(st + 16384), idesc_s, 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16(tDP, umma_desc_sw128(sDO + k * 32), umma_desc_sw128(st + TS + k * 32), idesc_s, k != 0);
        if (kX) tc_mma_f16(tDP, umma_desc_sw32(sDO + 16384), umma_desc_sw32(st + TS + 16384), idesc_s, 1);
        tc_commit(bar(SD_FULL));
      }
      __syncwarp();
      if (j > 0) mma2(j - 1);
    }
    mma2(NB - 1);
    // ---- table back-projection: dQ += dT_h Rh + dT_w Rw
    mbar_wait(bar(DT_FULL), 0);
    mbar_wait(bar(TAB_FULL), 1);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const uint32_t dt = w == 0 ? sDTh : sDTw, tab = w == 0 ? sQ : sDO;
#pragma unroll
        for (int kk = 0; kk < 2 * G / 16; ++kk) {
          const uint64_t a = umma_desc_sw128(dt + (kk >> 2) * 16384 + (kk & 3) * 32);
          tc_mma_f16(tDQ, a, umma_desc_sw128(tab + kk * 2048), idesc_o, 1);
          if (kX) tc_mma_f16(tDQ + 64, a, umma_desc_sw32(tab + C::kTabMain + kk * 512), idesc_ox, 1);
        }
      }
      tc_commit(bar(ACC2_FULL));
    }
    __syncwarp();
  } else {
    // ===================== elementwise warps: two threads per query row =====================
    const int quad = warp & 3;
    const int hs = (warp - 2) >> 2;
    const int wi = warp - 2;
    const int row = quad * 32 + lane;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    constexpr float kL2e = 1.4426950408889634f;
    const float c_l2 = kScale * kL2e;
    const int kw_base = ((hs & 1) * 32) % G;             // the 32 key columns this thread sees in every chunk
    const int qn = ib * 128 + row, qh = qn / G, qw = qn % G;
    const size_t rh = ((size_t)tok0 + qn) * heads + h;   // (token, head) row of rel / lse / D
    const float my_lse = lse[rh], my_d = dsum[rh];
    auto sync256 = []() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    float relw[32], aw[32];
    // ---- prologue: the bias rows of the 128 queries from T_w (phase 0) and T_h (phase 1)
    mbar_wait(bar(T_FULL), 0);
    tc_fence_after();
#pragma unroll 1
    for (int ph = 0; ph < 2; ++ph) {
      const int q0 = ph == 0 ? qw : qh;
#pragma unroll 1
      for (int c = hs; c < 2 * G / 32; c += 2) {         // T[q, a] with a = q0 - k + G - 1: scatter to [k][row]
        uint32_t r[32];
        tmem_ld_x32((ph == 0 ? tDP : tS) + c * 32 + tlane, r);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int k = q0 + G - 1 - (c * 32 + jj);
          if ((unsigned)k < (unsigned)G) stg_s[k * 129 + row] = __uint_as_float(r[jj]);
        }
      }
      if (ph == 1) {                                     // T_h / T_w are out of tensor memory: the first S / dP may overwrite them
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(T_READ));
      }
      sync256();
      if (ph == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { relw[j] = stg_s[(kw_base + j) * 129 + row] * kL2e; aw[j] = 0.f; }
      } else {
        for (int kh = hs * (G / 2); kh < (hs + 1) * (G / 2); ++kh) relh_s[kh * 128 + row] = stg_s[kh * 129 + row] * kL2e;
      }
      // the key-side kernel reads these rows from the global table: 16 rows per warp, one coalesced row segment per instruction
      for (int rr = 0; rr < 16; ++rr) {
        const int r = wi * 16 + rr;
        float* dst = rel_out + (((size_t)tok0 + ib * 128 + r) * heads + h) * (2 * G) + (ph == 0 ? G : 0);
        if (G == 64) *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(stg_s[(2 * lane) * 129 + r], stg_s[(2 * lane + 1) * 129 + r]);
        else dst[lane] = stg_s[lane * 129 + r];
      }
      sync256();                                         // the transpose buffer is rewritten by the next phase / zeroed below
    }
    for (int kh = 0; kh < G; ++kh) ah_s[(hs * G + kh) * 128 + row] = 0.f;
    sync256();
#pragma unroll 1
    for (int j = 0; j < NB; ++j) {
      mbar_wait(bar(SD_FULL), j & 1u);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * cc + hs;
        const int kh = (j * 128 + c * 32) / G;
        uint32_t rs[32], rd[32];
        tmem_ld_x32(tS + c * 32 + tlane, rs);
        tmem_ld_x32(tDP + c * 32 + tlane, rd);
        tmem_ld_wait();
        if (cc == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(SD_EMPTY));
        }
        const float off = relh_s[kh * 128 + row] - my_lse;
        float asum = 0.f;
        uint32_t pd[16];
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
          float p0, p1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(rs[jj]), c_l2, relw[jj]) + off));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(rs[jj + 1]), c_l2, relw[jj + 1]) + off));
          const float d0 = p0 * (__uint_as_float(rd[jj]) - my_d), d1 = p1 * (__uint_as_float(rd[jj + 1]) - my_d);
          aw[jj] += d0; aw[jj + 1] += d1;
          asum += d0 + d1;
          pd[jj >> 1] = pack_bf16(d0 * kScale, d1 * kScale);
        }
        ah_s[(hs * G + kh) * 128 + row] += asum;         // G = 64: the only contribution of this thread to (row, kh); G = 32: likewise
        if (cc == 0 && j > 0) { mbar_wait(bar(PD_EMPTY), (j - 1) & 1u); tc_fence_after(); }
        tmem_st_x16(tDS + c * 16 + tlane, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(PD_FULL));
    }
    // ---- epilogue: the cotangents of T_h / T_w as bf16 operand tiles over the retired K / V stages
    mbar_wait(bar(ACC_FULL), 0);
    tc_fence_after();
    {
      const int et = wi * 32 + lane;
      constexpr int kZ = 2 * C::kDT / 256;                // bytes per thread
#pragma unroll
      for (int i = 0; i < kZ / 16; ++i) *reinterpret_cast<uint4*>(dt_b + et * kZ + i * 16) = make_uint4(0, 0, 0, 0);
    }
    if (G != 64) {                                       // G = 32: both threads saw all 32 columns (different key rows)
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) xch_s[(hs * 32 + jj) * 128 + row] = aw[jj];
    }
    sync256();
    auto dt_put = [&](int tile, int col, float v) {      // dT[row, col] in 64-column sub-tiles of 128-byte rows, SWIZZLE_128B
      *reinterpret_cast<__nv_bfloat16*>(dt_b + tile * C::kDT + (col >> 6) * 16384 + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + (col & 7) * 2) =
          __float2bfloat16(v);
    };
    for (int kh = hs * (G / 2); kh < (hs + 1) * (G / 2); ++kh) dt_put(0, qh + G - 1 - kh, ah_s[kh * 128 + row] + ah_s[(G + kh) * 128 + row]);
    if (G == 64) {                                       // the two threads of a row own disjoint halves of the key columns
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) dt_put(1, qw + G - 1 - (kw_base + jj), aw[jj]);
    } else {
      for (int jj = hs * 16; jj < hs * 16 + 16; ++jj) dt_put(1, qw + G - 1 - jj, xch_s[jj * 128 + row] + xch_s[(32 + jj) * 128 + row]);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(DT_FULL));
    mbar_wait(bar(ACC2_FULL), 0);
    tc_fence_after();
    {
      uint32_t r[32], rx[8];
      tmem_ld_x32(tDQ + hs * 32 + tlane, r);
      if (kX) tmem_ld_x8(tDQ + 64 + hs * 8 + tlane, rx);
      tmem_ld_wait();
      __nv_bfloat16* o = dqkv + ((size_t)tok0 + qn) * (3 * D) + h * HD;
#pragma unroll
      for (int jj = 0; jj < 32; jj += 8)
        *reinterpret_cast<uint4*>(o + hs * 32 + jj) =
            make_uint4(pack_bf16(__uint_as_float(r[jj]), __uint_as_float(r[jj + 1])), pack_bf16(__uint_as_float(r[jj + 2]), __uint_as_float(r[jj + 3])),
                       pack_bf16(__uint_as_float(r[jj + 4]), __uint_as_float(r[jj + 5])), pack_bf16(__uint_as_float(r[jj + 6]), __uint_as_float(r[jj + 7])));
      if (kX)
        *reinterpret_cast<uint4*>(o + 64 + hs * 8) =
            make_uint4(pack_bf16(__uint_as_float(rx[0]), __uint_as_float(rx[1])), pack_bf16(__uint_as_float(rx[2]), __uint_as_float(rx[3])),
                       pack_bf16(__uint_as_float(rx[4]), __uint_as_float(rx[5])), pack_bf16(__uint_as_float(rx[6]), __uint_as_float(rx[7])));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// aux[i] = (lse[i], D[i], 0, 0) for i over (token, head)
__global__ void pack_lse_dsum_kernel(const float* __restrict__ lse, const float* __restrict__ dsum, float4* __restrict__ aux, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) aux[i] = make_float4(lse[i], dsum[i], 0.f, 0.f);
}

int make_tmap_bf16_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows);

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 tensor [d2][d1][d0] (d0 contiguous), box (b0, 1, b2), no swizzle
static int make_tmap_f32_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b2) {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      grove_set_error("cuTensorMapEncodeTiled entry point not available");
      return GROVE_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  cuuint64_t gdim[3] = {d0, d1, d2}, gstr[2] = {d0 * 4, d0 * d1 * 4};
  cuuint32_t bdim[3] = {b0, 1, b2}, estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { grove_set_error("cuTensorMapEncodeTiled (fp32, 3-D) failed with CUresult %d", (int)r); return GROVE_ERR_CUDA; }
  return GROVE_OK;
}

static int make_bwd_tmaps(BwdKvTmaps& tm, const void* qkv, const void* dO, int hd, int heads, long long M, int bq) {
  const int D = heads * hd;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tm.qkv, qkv, (uint64_t)3 * D, (uint64_t)M, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.dO, dO, (uint64_t)D, (uint64_t)M, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.qkv_b, qkv, (uint64_t)3 * D, (uint64_t)M, 64, bq))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.dO_b, dO, (uint64_t)D, (uint64_t)M, 64, bq))) return rc;
  if (hd > 64) {
    if ((rc = make_tmap_bf16_2d(&tm.qkv_x, qkv, (uint64_t)3 * D, (uint64_t)M, 16, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.dO_x, dO, (uint64_t)D, (uint64_t)M, 16, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.qkv_bx, qkv, (uint64_t)3 * D, (uint64_t)M, 16, bq))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.dO_bx, dO, (uint64_t)D, (uint64_t)M, 16, bq))) return rc;
  } else {
    tm.qkv_x = tm.qkv; tm.dO_x = tm.dO; tm.qkv_bx = tm.qkv; tm.dO_bx = tm.dO;
  }
  return GROVE_OK;
}

template <int G, int HD>
static int launch_kv(const BwdKvTmaps& tm, void* dqkv, int F, int heads, cudaStream_t st) {
  constexpr int smem = BwdKvCfg<G, HD>::kSmem;
  static_assert(smem <= 232448, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_kv_tc_kernel<G, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  attn_bwd_kv_tc_kernel<G, HD><<<dim3(G * G / 128, heads, F), kBwdThreads, smem, st>>>(tm, (__nv_bfloat16*)dqkv, heads);
  return GROVE_OK;
}

template <int G, int HD>
static int launch_q(const BwdKvTmaps& tm, float* rel, const float* lse, const float* dsum, void* dqkv, int F, int heads, cudaStream_t st) {
  constexpr int smem = BwdQCfg<G, HD>::kSmem;
  static_assert(smem <= 232448, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_q_tc_kernel<G, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  attn_bwd_q_tc_kernel<G, HD><<<dim3(G * G / 128, heads, F), kBwdThreads, smem, st>>>(tm, rel, lse, dsum, (__nv_bfloat16*)dqkv, heads);
  return GROVE_OK;
}

// called from run_attn_bwd (attention_bwd.cu) for global layers with head dim 64 / 80 on 32x32 / 64x64 grids.
// rel: fp32 [M, heads, 2G] bias rows; lse, dsum: fp32 [M, heads]; aux: fp32 scratch [M, heads, 4]
int launch_attn_bwd_kv_tc(const void* qkv, const void* dO, const float* rel, const float* lse, const float* dsum, float* aux, void* dqkv, int F,
                          int G, int heads, int hd, cudaStream_t st) {
  const int N = G * G;
  const long long M = (long long)F * N;
  pack_lse_dsum_kernel<<<(unsigned)((M * heads + 255) / 256), 256, 0, st>>>(lse, dsum, reinterpret_cast<float4*>(aux), M * heads);
  grove_count_launch();
  const int bq = hd > 64 ? 64 : 128;
  BwdKvTmaps tm;
  int rc;
  if ((rc = make_bwd_tmaps(tm, qkv, dO, hd, heads, M, bq))) return rc;
  if ((rc = make_tmap_f32_3d(&tm.relw, rel, (uint64_t)2 * G, heads, (uint64_t)M, (uint32_t)G, bq))) return rc;
  if ((rc = make_tmap_f32_3d(&tm.relh, rel, (uint64_t)2 * G, heads, (uint64_t)M, 4, bq))) return rc;
  if ((rc = make_tmap_f32_3d(&tm.aux, aux, 4, heads, (uint64_t)M, 4, bq))) return rc;
  tm.rh = tm.qkv; tm.rw = tm.qkv; tm.rh_x = tm.qkv; tm.rw_x = tm.qkv;      // unused by this kernel
  if (G == 64) rc = hd == 64 ? launch_kv<64, 64>(tm, dqkv, F, heads, st) : launch_kv<64, 80>(tm, dqkv, F, heads, st);
  else rc = hd == 64 ? launch_kv<32, 64>(tm, dqkv, F, heads, st) : launch_kv<32, 80>(tm, dqkv, F, heads, st);
  if (rc) return rc;
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

// query side (see attn_bwd_q_tc_kernel); same preconditions as launch_attn_bwd_kv_tc.  Rh, Rw: bf16 [2G-1, hd].  Writes the bias rows
// rel [M, heads, 2G] (fp32, read by the key side) and the q slot of dqkv.
int launch_attn_bwd_q_tc(const void* qkv, const void* dO, const void* Rh, const void* Rw, float* rel, const float* lse, const float* dsum, void* dqkv,
                         int F, int G, int heads, int hd, cudaStream_t st) {
  const long long M = (long long)F * G * G;
  BwdKvTmaps tm;
  int rc;
  if ((rc = make_bwd_tmaps(tm, qkv, dO, hd, heads, M, 128))) return rc;
  tm.relw = tm.qkv; tm.relh = tm.qkv; tm.aux = tm.qkv;      // unused by this kernel
  if ((rc = make_tmap_bf16_2d(&tm.rh, Rh, (uint64_t)hd, (uint64_t)2 * G - 1, 64, 2 * G))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.rw, Rw, (uint64_t)hd, (uint64_t)2 * G - 1, 64, 2 * G))) return rc;
  if (hd > 64) {
    if ((rc = make_tmap_bf16_2d(&tm.rh_x, Rh, (uint64_t)hd, (uint64_t)2 * G - 1, 16, 2 * G))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.rw_x, Rw, (uint64_t)hd, (uint64_t)2 * G - 1, 16, 2 * G))) return rc;
  } else {
    tm.rh_x = tm.rh; tm.rw_x = tm.rw;
  }
  if (G == 64) rc = hd == 64 ? launch_q<64, 64>(tm, rel, lse, dsum, dqkv, F, heads, st) : launch_q<64, 80>(tm, rel, lse, dsum, dqkv, F, heads, st);
  else rc = hd == 64 ? launch_q<32, 64>(tm, rel, lse, dsum, dqkv, F, heads, st) : launch_q<32, 80>(tm, rel, lse, dsum, dqkv, F, heads, st);
  if (rc) return rc;
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

}  // namespace grove
