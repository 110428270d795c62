// Global attention with decomposed rel-pos bias on tcgen05 / TMEM / TMA (image_encoder.py:301-326, 420-458).
//
// One CTA = 128 queries of one (frame, head).  Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 softmax:
// two threads per query row (= TMEM lane), each owning the keys with kw in one half of the grid row, so the row softmax
// needs only one max / one sum exchange per tile and every SM sub-partition has two softmax warps to hide latency.
//   prologue  T_w = Q.Rw^T and T_h = Q.Rh^T as two UMMAs (128x128x64) -> each thread gathers its own
//             rel_w[kw] = T_w[qw-kw+G-1] into registers and rel_h[kh] into shared memory (pre-scaled by log2 e)
//   main loop  one pass over the keys, 128 per block: S = Q.K^T (UMMA 128x128x64, Q read from tensor memory, S double-buffered in
//             TMEM) -> block max -> p = exp2(scale*S + bias - m) -> bf16 P written with tcgen05.st into tensor memory -> O += P.V
//             (TS-mode UMMA 128x64x128, V consumed MN-major exactly as TMA lands it).  m is a LAZY running maximum: it is raised
//             (and the 32 rows x HD accumulator slice of that warp rescaled in TMEM) only when a row's block max exceeds it by
//             more than 8 (log2 domain), so p <= 256 and the correction runs a handful of times per tile, not once per block.
// Scores never leave the SM.  (Round-1 history: a two-phase exact-max variant spent 29 % of its time recomputing Q.K^T.)
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "grove_b200.h"
#include "tmem_ldst.cuh"

namespace grove {

// In-kernel timeline probes (compile with -DGROVE_ATT_PROBE; read with grove_att_probe_read, see profiles/att_probe.py): clock64 of
// one CTA's MMA warp and softmax warps at every hand-off.  Compiled out by default.
#ifdef GROVE_ATT_PROBE
__device__ long long g_att_probe[8192];
#define PROBE(slot) do { if (probe_on) g_att_probe[(slot)] = clock64(); } while (0)
#else
#define PROBE(slot) do { } while (0)
#endif


// warp 0 TMA, warp 1 MMA, then 4 * SPLIT softmax warps: SPLIT threads per query row (= TMEM lane), each owning 128 / SPLIT keys of a block
template <int SPLIT> constexpr int att_threads() { return 64 + 128 * SPLIT; }
constexpr int kVStages = 2;
template <int HD> constexpr int k_stages() { return HD == 64 ? 4 : 3; }

// Head dim 80 (ViT-H) = a 64-wide part (128-byte rows, SWIZZLE_128B) + a 16-wide tail (32-byte rows, SWIZZLE_32B): every operand
// tile is [128 x 64 | 128 x 16], Q.K^T takes a 5th k-step from the tails, P.V issues a second N=16 MMA into O columns 64-79.
struct AttTmaps { CUtensorMap qkv, rh, rw, qkv_x, rh_x, rw_x; };

template <int G, int HD, int SPLIT>
__global__ void __launch_bounds__(att_threads<SPLIT>(), 1)
attn_global_tc_kernel(const __grid_constant__ AttTmaps tm, __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, int heads) {
  constexpr int CPT = 4 / SPLIT;                    // 32-key chunks of a 128-key block per softmax thread
  constexpr bool kX = HD > 64;                      // has the 16-wide tail
  constexpr int TS = 16384 + (kX ? 4096 : 0);       // bytes of one [128 x HD] operand tile
  constexpr int kKStages = k_stages<HD>();
  const CUtensorMap& tmap_qkv = tm.qkv; const CUtensorMap& tmap_rh = tm.rh; const CUtensorMap& tmap_rw = tm.rw;
  constexpr int N = G * G;
  constexpr int NB = N / 128;          // key blocks
  constexpr int NKW = G;               // rel_w entries per row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = s0;                         // Q tile [128][HD] bf16
  const uint32_t sK = sQ + TS;                    // kKStages K tiles
  const uint32_t sV = sK + kKStages * TS;         // 2 V tiles; during the prologue: Rw | Rh tables
  const uint32_t sP = sV + kVStages * TS;      // 2 x 32 KB P buffers (two 64-key slabs each); prologue: fp32 staging [128][128]
  const uint32_t sRelH = sP + 65536;              // [G][128] fp32
  const uint32_t sXch = sRelH + G * 128 * 4;      // 3 x [SPLIT][128] fp32: block max (double-buffered) and final sum exchange between the threads of a row
  const uint32_t bar0 = sXch + 3 * SPLIT * 128 * 4;
  uint8_t* smem_al = smem_raw + (s0 - smem_u32(smem_raw));
  float* stage_f = reinterpret_cast<float*>(smem_al + (sP - s0));
  float* relh_f = reinterpret_cast<float*>(smem_al + (sRelH - s0));
  float* xch_f = reinterpret_cast<float*>(smem_al + (sXch - s0));
  enum { Q_FULL = 0, TAB_FREE, O_FULL, Q_TMEM, K_FULL, K_EMPTY = K_FULL + kKStages, V_FULL = K_EMPTY + kKStages, V_EMPTY = V_FULL + kVStages,
         S_FULL = V_EMPTY + kVStages, S_EMPTY = S_FULL + 2, P_FULL = S_EMPTY + 2, P_EMPTY = P_FULL + 2, NUM_BARS = P_EMPTY + 2 };
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_slot = bar0 + 8u * NUM_BARS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, f = blockIdx.z;
#ifdef GROVE_ATT_PROBE
  const bool probe_on = blockIdx.x == 3 && blockIdx.y == 1 && blockIdx.z == 0 && lane == 0;
#endif
  const int D = heads * HD;
  const int tok0 = f * N;                // first token row of this frame in the [F*N, 3D] qkv matrix

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(bar(Q_FULL), 1); mbar_init(bar(TAB_FREE), 1); mbar_init(bar(O_FULL), 1); mbar_init(bar(Q_TMEM), 4 * SPLIT);
    for (int i = 0; i < kKStages; ++i) { mbar_init(bar(K_FULL + i), 1); mbar_init(bar(K_EMPTY + i), 1); }
    for (int i = 0; i < kVStages; ++i) { mbar_init(bar(V_FULL + i), 1); mbar_init(bar(V_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(S_FULL + i), 1); mbar_init(bar(S_EMPTY + i), 4 * SPLIT);
      mbar_init(bar(P_FULL + i), 4 * SPLIT); mbar_init(bar(P_EMPTY + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                          // barrier init / TMEM allocation above overlapped the predecessor's last wave
  pdl_launch_dependents();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // S0|S1 (2x128), O (<=80), Q (HD/2 <= 40: bf16 pairs, the A operand of every Q.K^T), P0|P1 (2x64: bf16 pairs)
  const uint32_t tS0 = tmem_base, tO = tmem_base + 256, tQ = tmem_base + 336, tP0 = tmem_base + 384;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      auto load_tile = [&](uint32_t dst, const CUtensorMap* main_map, const CUtensorMap* tail_map, uint32_t b, int col, int rowc) {
        tma_load_2d(dst, main_map, b, col, rowc);
        if (kX) tma_load_2d(dst + 16384, tail_map, b, col + 64, rowc);
      };
      mbar_expect_tx(bar(Q_FULL), 3 * TS);
      load_tile(sQ, &tmap_qkv, &tm.qkv_x, bar(Q_FULL), h * HD, tok0 + q0);
      load_tile(sV, &tmap_rw, &tm.rw_x, bar(Q_FULL), 0, 0);          // Rw table -> first V slot (rows >= 2G-1 zero-filled)
      load_tile(sV + TS, &tmap_rh, &tm.rh_x, bar(Q_FULL), 0, 0);     // Rh table -> second V slot
      uint32_t kit = 0, vit = 0;
      for (int b = 0; b < NB; ++b, ++kit, ++vit) {             // K and V of every 128-key block
        const int s = kit % kKStages;
        mbar_wait(bar(K_EMPTY + s), ((kit / kKStages) & 1u) ^ 1u);
        mbar_expect_tx(bar(K_FULL + s), TS);
        load_tile(sK + s * TS, &tmap_qkv, &tm.qkv_x, bar(K_FULL + s), D + h * HD, tok0 + b * 128);
        const int v = vit % kVStages;
        if (b == 0) mbar_wait(bar(TAB_FREE), 0);               // the prologue MMAs have finished reading the tables (they sit in the V slots)
        mbar_wait(bar(V_EMPTY + v), ((vit / kVStages) & 1u) ^ 1u);
        mbar_expect_tx(bar(V_FULL + v), TS);
        load_tile(sV + v * TS, &tmap_qkv, &tm.qkv_x, bar(V_FULL + v), 2 * D + h * HD, tok0 + b * 128);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) is MN-major
    constexpr uint32_t idesc_ox = umma_idesc_bf16(128, 16) | (1u << 16);  // the 16-wide tail of V -> O columns 64..79
    uint32_t sit = 0, kit = 0, vit = 0;
    // S[sit&1] = Q . B^T  (B: [128 rows][HD] K-major in shared memory).  The two rel-pos tables take Q from shared memory (SS); every
    // key block takes Q from TENSOR memory (TS): an SS-mode 128x128x16 UMMA reads 8 KB of operands per 64-clk instruction, which is the
    // whole 128 B/clk shared-memory bandwidth of the SM and made Q.K^T run at half rate next to the TMA writes and the softmax's traffic.
    auto issue_s = [&](uint32_t b_smem, bool q_tmem) {
      const uint32_t sb = sit & 1u;
      tc_fence_after();
      if (elect_one()) {
        if (q_tmem) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16_ts(tS0 + sb * 128, tQ + k * 8, umma_desc_sw128(b_smem + k * 32), idesc_s, k != 0);
          if (kX) tc_mma_f16_ts(tS0 + sb * 128, tQ + 32, umma_desc_sw32(b_smem + 16384), idesc_s, 1);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tS0 + sb * 128, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(b_smem + k * 32), idesc_s, k != 0);
          if (kX) tc_mma_f16(tS0 + sb * 128, umma_desc_sw32(sQ + 16384), umma_desc_sw32(b_smem + 16384), idesc_s, 1);
        }
      }
      __syncwarp();
    };
    // S buffer `sit & 1` is free for the first two uses; afterwards wait for the softmax warps to have drained it
    auto s_empty_par = [&]() { return ((sit >> 1) & 1u) ^ 1u; };
    mbar_wait(bar(Q_FULL), 0);
    issue_s(sV, false);                                // T_w
    if (elect_one()) tc_commit(bar(S_FULL + 0));
    __syncwarp();
    ++sit;
    issue_s(sV + TS, false);                           // T_h
    if (elect_one()) { tc_commit(bar(S_FULL + 1)); tc_commit(bar(TAB_FREE)); }
    __syncwarp();
    ++sit;
    mbar_wait(bar(Q_TMEM), 0);                         // the softmax warps have copied Q into tensor memory
    // one key block: wait for (K landed, S buffer drained) with both polls in flight, then 4-5 UMMAs and two commits
    auto issue_qk = [&]() {
      const int s = kit % kKStages;
      PROBE(1000 + 3 * kit);
      mbar_wait2(bar(K_FULL + s), (kit / kKStages) & 1u, bar(S_EMPTY + (sit & 1u)), s_empty_par());
      PROBE(1001 + 3 * kit);
      issue_s(sK + s * TS, true);
      PROBE(4000 + 3 * kit);
      if (elect_one()) { tc_commit(bar(K_EMPTY + s)); PROBE(4001 + 3 * kit); tc_commit(bar(S_FULL + (sit & 1u))); PROBE(4002 + 3 * kit); }
      __syncwarp();
      PROBE(1002 + 3 * kit);
      ++kit; ++sit;
    };
    // Issue schedule: S(b+1) is issued before P(b).V(b) so the softmax of block b+1 overlaps the P.V of block b.
    // (Round-2 experiment, -DGROVE_ATT_INTERLEAVE: back-to-back UMMAs into the same accumulator run at about twice their nominal duration,
    // so Q.K^T(b+2) and P(b).V(b) -- independent accumulators -- were issued interleaved, S two blocks ahead.  Bit-identical results,
    // but 929 us instead of 704 us per layer: alternating instruction shapes costs more than the accumulator dependency.)
#ifndef GROVE_ATT_INTERLEAVE
    issue_qk();
    for (int b = 0; b < NB; ++b, ++vit) {
      if (b + 1 < NB) issue_qk();
      const uint32_t pb = b & 1u;
      const int v = vit % kVStages;
      mbar_wait2(bar(V_FULL + v), (vit / kVStages) & 1u, bar(P_FULL + pb), (b >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t ta = tP0 + pb * 64 + kk * 8;
          const uint64_t db = umma_desc_sw128(sV + v * TS + kk * 2048);
          tc_mma_f16_ts(tO, ta, db, idesc_o, (b | kk) != 0);
          if (kX) tc_mma_f16_ts(tO + 64, ta, umma_desc_sw32(sV + v * TS + 16384 + kk * 512), idesc_ox, (b | kk) != 0);
        }
        tc_commit(bar(V_EMPTY + v));
        tc_commit(bar(P_EMPTY + pb));
        if (b == NB - 1) tc_commit(bar(O_FULL));
      }
      __syncwarp();
    }
#else
    issue_qk();
    if (NB > 1) issue_qk();
    for (int b = 0; b < NB; ++b, ++vit) {
      const uint32_t pb = b & 1u;
      const int v = vit % kVStages;
      const bool more = b + 2 < NB;
      const int s = kit % kKStages;
      const uint32_t sb = sit & 1u;
      PROBE(1300 + 3 * b);
      mbar_wait2(bar(V_FULL + v), (vit / kVStages) & 1u, bar(P_FULL + pb), (b >> 1) & 1u);
      if (more) mbar_wait2(bar(K_FULL + s), (kit / kKStages) & 1u, bar(S_EMPTY + sb), s_empty_par());
      PROBE(1301 + 3 * b);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_smem = sK + s * TS;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (more) tc_mma_f16_ts(tS0 + sb * 128, tQ + k * 8, umma_desc_sw128(b_smem + k * 32), idesc_s, k != 0);
#pragma unroll
          for (int kk = 2 * k; kk < 2 * k + 2; ++kk) {
            const uint32_t ta = tP0 + pb * 64 + kk * 8;        // P[:, 16 keys] = 8 packed columns of tensor memory
            tc_mma_f16_ts(tO, ta, umma_desc_sw128(sV + v * TS + kk * 2048), idesc_o, (b | kk) != 0);
            if (kX) tc_mma_f16_ts(tO + 64, ta, umma_desc_sw32(sV + v * TS + 16384 + kk * 512), idesc_ox, (b | kk) != 0);
          }
        }
        if (kX && more) tc_mma_f16_ts(tS0 + sb * 128, tQ + 32, umma_desc_sw32(b_smem + 16384), idesc_s, 1);
        tc_commit(bar(V_EMPTY + v));
        tc_commit(bar(P_EMPTY + pb));
        if (b == NB - 1) tc_commit(bar(O_FULL));
        if (more) { tc_commit(bar(K_EMPTY + s)); tc_commit(bar(S_FULL + sb)); }
      }
      __syncwarp();
      if (more) { ++kit; ++sit; }
      PROBE(1302 + 3 * b);
    }
#endif
  } else {
    // ===================== softmax warps: two threads per query row =====================
    const int quad = warp & 3;
    const int hs = (warp - 2) >> 2;                      // this thread owns key chunks {hs, hs + SPLIT, ..} of every block
    const int row = quad * 32 + lane;                    // row inside the tile == TMEM lane
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    const int q = q0 + row;
    const int qh = q / G, qw = q % G;
    constexpr float kL2e = 1.4426950408889634f;
    auto softmax_sync = []() { asm volatile("bar.sync 1, %0;" ::"n"(128 * SPLIT) : "memory"); };
    // staging row with the float4 slots XOR-swizzled by the row (conflict-free 128-bit stores of 32 different rows)
    auto stage_at = [&](int e) { return stage_f[row * 128 + ((((e >> 2) ^ (row & 7)) << 2) | (e & 3))]; };
    uint32_t sit = 0;
    // ---- Q tile: shared memory (128B-swizzled rows as TMA landed them) -> tensor memory, row = lane, bf16 pairs along the head dim
    {
      constexpr int CH = 8 / SPLIT;                      // 16-byte chunks (8 head dims) of the 64-wide part per thread
      mbar_wait(bar(Q_FULL), 0);
      uint32_t qr[4 * CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int ch = hs * CH + c;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(qr[4 * c]), "=r"(qr[4 * c + 1]), "=r"(qr[4 * c + 2]), "=r"(qr[4 * c + 3])
                     : "r"(sQ + row * 128 + ((ch ^ (row & 7)) << 4)));
      }
      if constexpr (CH == 4) tmem_st_32x32b_x16(tQ + hs * 16 + tlane, qr);
      else tmem_st_x8(tQ + hs * 8 + tlane, qr);
      if (kX && hs == 0) {                               // head dims 64..79: 32-byte rows, SWIZZLE_32B
        uint32_t qx[8];
#pragma unroll
        for (int c = 0; c < 2; ++c)
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(qx[4 * c]), "=r"(qx[4 * c + 1]), "=r"(qx[4 * c + 2]), "=r"(qx[4 * c + 3])
                       : "r"(sQ + 16384 + row * 32 + ((c ^ ((row >> 2) & 1)) << 4)));
        tmem_st_x8(tQ + 32 + tlane, qx);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_TMEM));
    }
    float2 relw[16];                                     // this thread's kw half (G=64) / the whole grid row (G=32), as FFMA2 operand pairs
    const int kw0 = (G == 64) ? (hs & 1) * 32 : 0;       // chunk c covers kw = (c % 2) * 32 .. + 32 of a 64-wide grid row; SPLIT is even
    // ---- prologue: rel_w -> registers, rel_h -> smem (both x log2 e)
#pragma unroll
    for (int which = 0; which < 2; ++which, ++sit) {
      mbar_wait(bar(S_FULL + (sit & 1u)), (sit >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {                 // the threads of a row share the 128 table columns
        const int c = SPLIT * cc + hs;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tS0 + (sit & 1u) * 128 + c * 32 + tlane, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage_f + row * 128 + (((c * 8 + (j >> 2)) ^ (row & 7)) << 2)) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(S_EMPTY + (sit & 1u)));
      softmax_sync();                                    // both halves of every staging row are written
      if (which == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          relw[j >> 1] = make_float2(stage_at(qw + (G - 1) - (kw0 + j)) * kL2e, stage_at(qw + (G - 1) - (kw0 + j + 1)) * kL2e);
      } else {
        for (int kh = hs * (G / SPLIT); kh < (hs + 1) * (G / SPLIT); ++kh) relh_f[kh * 128 + row] = stage_at(qh + (G - 1) - kh) * kL2e;
      }
      softmax_sync();                                    // staging is rewritten by the next table / rel_h complete
    }
    const float c_scale = (HD == 64 ? 0.125f : 0.11180339887498949f) * kL2e;   // hd^-0.5 * log2(e)
    const float2 c_scale2 = make_float2(c_scale, c_scale);
    // ---- main loop: lazy running max, probabilities, P.V
    auto quad_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(2 + quad), "n"(32 * SPLIT) : "memory"); };   // the SPLIT warps that share 32 rows
    float m = -INFINITY;
    float2 lsum2 = make_float2(0.f, 0.f);
    for (int b = 0; b < NB; ++b, ++sit) {
      const uint32_t sb = sit & 1u, pb = b & 1u;
      // S(b) complete, and P.V of block b-2 has finished reading this P buffer (both polls in flight together)
      PROBE(2100 + (warp - 2) * 400 + 4 * b);
      mbar_wait2(bar(S_FULL + sb), (sit >> 1) & 1u, bar(P_EMPTY + pb), ((b >> 1) & 1u) ^ 1u);
      tc_fence_after();
      uint32_t rr[CPT][32];
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) tmem_ld_32x32b_x32(tS0 + sb * 128 + (SPLIT * cc + hs) * 32 + tlane, rr[cc]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(S_EMPTY + sb));     // S is in registers: the MMA warp may overwrite this buffer
      PROBE(2101 + (warp - 2) * 400 + 4 * b);
      // block maximum of scale*S + rel_h + rel_w over this thread's keys, then over the SPLIT threads of the row
      float mb = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {
        const int kh = (b * 128 + (SPLIT * cc + hs) * 32) / G;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {                 // scale*S + rel_w is kept in place: the exponential pass adds only a per-chunk offset
          const float2 e0 = __ffma2_rn(make_float2(__uint_as_float(rr[cc][j]), __uint_as_float(rr[cc][j + 1])), c_scale2, relw[j >> 1]);
          const float2 e1 = __ffma2_rn(make_float2(__uint_as_float(rr[cc][j + 2]), __uint_as_float(rr[cc][j + 3])), c_scale2, relw[(j >> 1) + 1]);
          rr[cc][j] = __float_as_uint(e0.x); rr[cc][j + 1] = __float_as_uint(e0.y);
          rr[cc][j + 2] = __float_as_uint(e1.x); rr[cc][j + 3] = __float_as_uint(e1.y);
          mx0 = fmaxf(mx0, fmaxf(e0.x, e0.y));            // FMNMX3
          mx1 = fmaxf(mx1, fmaxf(e1.x, e1.y));
        }
        mb = fmaxf(mb, fmaxf(mx0, mx1) + relh_f[kh * 128 + row]);
      }
      float* xb = xch_f + (b & 1) * SPLIT * 128;         // double-buffered: the write of block b+2 is behind the barrier of block b+1
      xb[hs * 128 + row] = mb;
      quad_sync();
#pragma unroll
      for (int o = 1; o < SPLIT; ++o) mb = fmaxf(mb, xb[((hs + o) % SPLIT) * 128 + row]);
      // lazy rescale: raise m only when some row of this warp outgrew it by more than 2^8 (always on the first block).  Every
      // thread of a row sees the same mb and m, so the SPLIT warps of the quad take this branch together.
      const bool grow = mb > m + 8.0f;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mb : m;
        float fac;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(fac) : "f"(m - m_new));   // 1 for rows that keep m, 0 on the first block
        if (b > 0) {
          // every P.V issued so far (the last one is block b-1; P.V of block b waits for our P_FULL arrive) has retired
          mbar_wait(bar(P_EMPTY + (pb ^ 1u)), ((b - 1) >> 1) & 1u);
          tc_fence_after();
          constexpr int DPT = 64 / SPLIT;                // accumulator columns of this thread
          uint32_t o[DPT];
          if constexpr (DPT == 32) tmem_ld_32x32b_x32(tO + hs * DPT + tlane, o);
          else tmem_ld_x16(tO + hs * DPT + tlane, o);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < DPT; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * fac);
          if constexpr (DPT == 32) tmem_st_x32(tO + hs * DPT + tlane, o);
          else tmem_st_32x32b_x16(tO + hs * DPT + tlane, o);
          if (kX && hs < 2) {
            uint32_t ox[8];
            tmem_ld_32x32b_x8(tO + 64 + hs * 8 + tlane, ox);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) ox[j] = __float_as_uint(__uint_as_float(ox[j]) * fac);
            tmem_st_x8(tO + 64 + hs * 8 + tlane, ox);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        lsum2 = __fmul2_rn(lsum2, make_float2(fac, fac));
        m = m_new;
      }
#pragma unroll
      for (int cc = 0; cc < CPT; ++cc) {
        const int c = SPLIT * cc + hs;
        const int kh = (b * 128 + c * 32) / G;
        const float off = relh_f[kh * 128 + row] - m;
        const float2 off2 = make_float2(off, off);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 e = __fadd2_rn(make_float2(__uint_as_float(rr[cc][j]), __uint_as_float(rr[cc][j + 1])), off2);
          float2 p;
          // 3 pairs in 8 take the packed FMA-pipe polynomial (10 issue slots per pair), the rest the MUFU (2 slots, 16 XU clocks per pair):
          // balances the sub-partition's issue port against its 4-lane-per-clock exponential unit
#ifndef GROVE_ATT_POLY
#define GROVE_ATT_POLY 0x52                               /* bit i: pair i of every 8 takes the polynomial (pairs 1, 4, 6).  Measured, 8 x 12 x 4096^2: */
                                                          /* 0x00 (all MUFU) 779 us, 0x12 745, 0x52 719, 0x5A 727-733; four row-sum chains: no change  */
#endif
          if ((GROVE_ATT_POLY >> ((j >> 1) & 7)) & 1) {
            p = ex2_fma2(e);
          } else {
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p.x) : "f"(e.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p.y) : "f"(e.y));
          }
          lsum2 = __fadd2_rn(lsum2, p);
          pk[j >> 1] = pack_bf16(p.x, p.y);
        }
        // keys [c*32, c*32+32) of the block -> 16 packed bf16x2 columns of the P buffer in tensor memory (the A operand of P.V)
        tmem_st_32x32b_x16(tP0 + pb * 64 + c * 16 + tlane, pk);
      }
      PROBE(2102 + (warp - 2) * 400 + 4 * b);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(P_FULL + pb));
      PROBE(2103 + (warp - 2) * 400 + 4 * b);
    }
    // ---- epilogue: O / l -> bf16 -> global (each thread of a row stores 64 / SPLIT of the first 64 head dims)
    float lsum = lsum2.x + lsum2.y;
    xch_f[2 * SPLIT * 128 + hs * 128 + row] = lsum;
    softmax_sync();
#pragma unroll
    for (int o = 1; o < SPLIT; ++o) lsum += xch_f[2 * SPLIT * 128 + ((hs + o) % SPLIT) * 128 + row];
    const float inv = 1.f / lsum;
    if (lse_out != nullptr && hs == 0) lse_out[((size_t)tok0 + q) * heads + h] = m + log2f(lsum);   // log2 domain, for the backward pass
    mbar_wait(bar(O_FULL), 0);
    tc_fence_after();
    constexpr int DPT = 64 / SPLIT;                      // head dims per thread
    __nv_bfloat16* orow = out + ((size_t)tok0 + q) * D + h * HD + hs * DPT;
    {
      uint32_t r[DPT];
      if constexpr (DPT == 32) tmem_ld_32x32b_x32(tO + hs * DPT + tlane, r);
      else tmem_ld_x16(tO + hs * DPT + tlane, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < DPT; j += 8)
        *reinterpret_cast<uint4*>(orow + j) =
            make_uint4(pack_bf16(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv), pack_bf16(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv),
                       pack_bf16(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv), pack_bf16(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv));
    }
    if (kX && hs < 2) {   // head dims 64..79: 8 per thread for two threads of the row
      uint32_t r[8];
      tmem_ld_32x32b_x8(tO + 64 + hs * 8 + tlane, r);
      tmem_ld_wait();
      *reinterpret_cast<uint4*>(out + ((size_t)tok0 + q) * D + h * HD + 64 + hs * 8) =
          make_uint4(pack_bf16(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv), pack_bf16(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv),
                     pack_bf16(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv), pack_bf16(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int G, int HD>
constexpr int att_tc_smem() {
  return (1 + k_stages<HD>() + kVStages) * (16384 + (HD > 64 ? 4096 : 0)) + 65536 + G * 128 * 4 + 6144 /*xch*/ + 1024 /*align*/ + 512 /*barriers*/;
}

int make_tmap_bf16_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows);

}  // namespace grove
using namespace grove;

static int att_split() {   // softmax threads per query row: 2 by default; GROVE_ATT_SPLIT=4 selects four (measured equal: the kernel is not
  static int v = 0;        // bound by softmax-warp latency)
  if (!v) { const char* e = getenv("GROVE_ATT_SPLIT"); v = (e && e[0] == '4') ? 4 : 2; }
  return v;
}

template <int G, int HD, int SPLIT>
static int launch_att_tc(const void* qkv, const void* rh, const void* rw, void* out, float* lse, int F, int heads, cudaStream_t stream) {
  const int N = G * G, D = heads * HD;
  AttTmaps tm;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tm.qkv, qkv, (uint64_t)3 * D, (uint64_t)F * N, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.rh, rh, HD, 2 * G - 1, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.rw, rw, HD, 2 * G - 1, 64, 128))) return rc;
  if (HD > 64) {
    if ((rc = make_tmap_bf16_2d(&tm.qkv_x, qkv, (uint64_t)3 * D, (uint64_t)F * N, 16, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.rh_x, rh, HD, 2 * G - 1, 16, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.rw_x, rw, HD, 2 * G - 1, 16, 128))) return rc;
  } else {
    tm.qkv_x = tm.qkv; tm.rh_x = tm.rh; tm.rw_x = tm.rw;
  }
  constexpr int smem = att_tc_smem<G, HD>();
  cudaError_t e = cudaFuncSetAttribute(attn_global_tc_kernel<G, HD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  grove_launch_pdl(attn_global_tc_kernel<G, HD, SPLIT>, dim3(N / 128, heads, F), dim3(att_threads<SPLIT>()), smem, stream, tm, (__nv_bfloat16*)out, lse, heads);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

template <int G, int HD>
static int launch_att_tc_split(const void* qkv, const void* rh, const void* rw, void* out, float* lse, int F, int heads, cudaStream_t stream) {
  return att_split() == 2 ? launch_att_tc<G, HD, 2>(qkv, rh, rw, out, lse, F, heads, stream)
                          : launch_att_tc<G, HD, 4>(qkv, rh, rw, out, lse, F, heads, stream);
}

extern "C" int grove_attn_global_relpos_fwd_lse(const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out, float* lse, int F, int G,
                                                int heads, int hd, cudaStream_t stream);

#ifdef GROVE_ATT_PROBE
extern "C" int grove_att_probe_read(long long* host, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host, g_att_probe, sizeof(long long) * n) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int grove_attn_global_relpos_fwd(const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out, int F, int G, int heads,
                                            int hd, cudaStream_t stream) {
  return grove_attn_global_relpos_fwd_lse(qkv, rel_pos_h, rel_pos_w, out, nullptr, F, G, heads, hd, stream);
}

extern "C" int grove_attn_global_relpos_fwd_lse(const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out, float* lse, int F, int G,
                                                int heads, int hd, cudaStream_t stream) {
  GROVE_CHECK_ARG(qkv && rel_pos_h && rel_pos_w && out && F > 0 && heads > 0);
  if ((hd != 64 && hd != 80) || (G != 64 && G != 32)) {
    grove_set_error("grove_attn_global_relpos_fwd: head dim 64 / 80 and G in {32,64} are built (got hd=%d G=%d)", hd, G);
    return GROVE_ERR_UNSUPPORTED;
  }
  GROVE_CHECK_ARG(F <= 65535 && heads <= 65535);
  GROVE_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)rel_pos_h & 15) == 0 && ((uintptr_t)rel_pos_w & 15) == 0);
  if (hd == 64) return G == 64 ? launch_att_tc_split<64, 64>(qkv, rel_pos_h, rel_pos_w, out, lse, F, heads, stream)
                               : launch_att_tc_split<32, 64>(qkv, rel_pos_h, rel_pos_w, out, lse, F, heads, stream);
  return G == 64 ? launch_att_tc_split<64, 80>(qkv, rel_pos_h, rel_pos_w, out, lse, F, heads, stream)
                 : launch_att_tc_split<32, 80>(qkv, rel_pos_h, rel_pos_w, out, lse, F, heads, stream);
}
