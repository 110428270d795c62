// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (UMMA 128 x BN x 16, fp32 accumulators double-buffered in TMEM) -> tcgen05.ld epilogue with fused
// bias / GELU / ReLU / tanh-gate / fp32 residual.  One kernel serves
//   * every Linear of the SAM ViT blocks (K4, K9, K10 of SURVEY.md §2.3), the patch-embed GEMM (K1),
//     the neck 1x1 conv (K12), text_hidden_fcs (K13) and the box decoder's image-side projections (K17),
//   * the Conv3d spatio-temporal adapter (K11) and the neck 3x3 conv (K12) as *implicit* GEMMs: the A
//     operand is fetched with a 5-D TMA box at shifted (w,h,t) coordinates, and TMA's out-of-bounds
//     zero fill implements the 'same' zero padding in t, h and w.
//
// out[m, n] = resid[m (mod resid_mod), n] + gate * act( sum_k A[m, k] * W[n, k] + bias[n] )
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "grove_b200.h"

namespace grove {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 320;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int kEpiWarps = 8;

struct GemmParams {
  int M, N;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int conv;            // 0: A is a plain [M,K] matrix; 1: implicit-GEMM taps over [V,T,G,G,C] (A shifted);
                       // 2: conv weight gradient — K runs over tokens, B is X^T [C,V,T,G,G] fetched at shifted (w,h,t)
  int G, rows_per_tile, tiles_per_frame, T, kt, kc_blocks;
  int cblocks_per_tap, kblocks_per_frame, rows_per_kblock, wg_C;   // conv == 2
  int splits, kb_per_split;                                   // split-K: partial sums to out + split*plane_rows*N (fp32)
  int m_blk0;          // first m-block (of 128*CTAS rows) of this launch: a launch may cover a row range of the problem
  int plane_rows;      // split-K: rows of one partial plane (the launch's own row range), partial row = row - m_blk0*128*CTAS
  const __nv_bfloat16* dact_pre;  // bf16 [M,N] or null: multiply by act'(dact_pre) (backward of GELU / ReLU)
  int dact;            // 1 exact-GELU derivative, 2 ReLU mask
  const float* bias;   // [N] or null
  const float* resid;  // fp32 [*,N] or null
  const __nv_bfloat16* resid16;  // bf16 [M,N] or null: bf16 residual stream (EPI = 4 / general bf16-output epilogue)
  // LayerNorm folded into the GEMM that consumes the normalised rows (EPI = 5 / 6): A is the raw bf16 stream, W carries gamma, and
  //   out[m, n] = rstd_m * (acc[m, n] - mu_m * ln_csum[n]) + bias[n]        (bias = b + W.beta, ln_csum[n] = sum_k (gamma*W)[n, k])
  // with mu_m / rstd_m from the per-row partial sums (sum x, sum x^2) that the PRODUCING epilogue (EPI = 4, ln_stats_out) wrote.
  const float* ln_stats;   // [M, ln_parts, 2] fp32 partial (sum, sum of squares) of row m of A, or null
  const float* ln_csum;    // [N]
  int ln_parts;
  float ln_eps, ln_inv_k;  // 1 / (width of the normalised row)
  float* ln_stats_out;     // EPI = 4: [M, 2 * num_n_blocks, 2] partial sums of the bf16 values written to `out`, or null
  int resid_mod;       // >0: residual row = m % resid_mod (abs-pos embedding broadcast over frames)
  const float* gate_alpha;  // non-null: gate = tanh(*gate_alpha)   (adapter, image_encoder.py:54)
  int act;             // 0 none, 1 exact GELU, 2 ReLU
  void* out;           // [M,N] fp32 or bf16
  int out_f32;
  __nv_bfloat16* out2; // optional extra bf16 copy of the output (feeds the next tensor-core op)
  int out2_pre;        // 1: out2 receives the value BEFORE the activation (saved for the backward pass)
};

// CTAS = 1: one CTA owns a 128 x BN tile.  CTAS = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) owns a 256 x BN tile —
// each CTA stages its own 128 A rows and HALF of the B tile (BN/2 rows), the leader issues UMMA 256 x BN x 16 that reads both
// CTAs' shared memory, and each CTA keeps its 128 accumulator rows in its own TMEM.  Per CTA and k-block that is 32 KB of
// L2->smem traffic instead of 48 KB for the same MMA work: these GEMMs are L2-bandwidth bound, not MMA bound.
template <int BN, int CTAS>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBRows = BN / CTAS;           // B rows staged by one CTA
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStageRowF = 36;              // epilogue staging: 32 rows x 32 fp32 columns per warp, padded row (144 B)
  static constexpr int kEpiBytes = kEpiWarps * 32 * kStageRowF * 4;
  static constexpr int kStages = (224 * 1024 - kEpiBytes) / kStageBytes > 6 ? 6 : (224 * 1024 - kEpiBytes) / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;           // two accumulator stages (power of two: 256 or 512)
};

// EPI selects the epilogue: 0 = general (every fused option, run-time flags); 1 / 2 = bf16 output with bias and no / GELU
// activation and nothing else (the qkv and fc1 GEMMs of every encoder block) as straight-line code: with K = 768 these tiles
// are epilogue-bound (the MMA warp was measured waiting ~4k clk per tile for a drained accumulator), and the general epilogue's
// run-time branches, 5k-instruction body (instruction-cache misses) and fp32 staging cost 13.7k clk per tile against an 8.5k clk
// main loop.
template <int BN, int CTAS, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b) {
  using Cfg = GemmCfg<BN, CTAS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int tiles_mn = p.num_m_blocks * p.num_n_blocks;    // m-blocks of 128*CTAS rows
  const int num_tiles = tiles_mn * p.splits;
  const int tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;
  // Conv3d taps that look one frame before the first / after the last frame of an 8-frame group multiply by TMA's zero fill only:
  // those k-blocks (a third of K for two frames in eight, 8 % of the adapter's MMA work and operand traffic) are skipped by the
  // producer and the MMA issuer alike.  The first k-block of a range is always executed so that every accumulator is written.
  // live k-blocks of a tile of frame-in-group ft: [lo, hi) (whole range unless ft is the first / last frame of its group)
  auto conv_live_lo = [&](int ft) { return (p.conv == 1 && p.kt == 3 && ft == 0) ? 9 * p.kc_blocks : 0; };
  auto conv_live_hi = [&](int ft) { return (p.conv == 1 && p.kt == 3 && ft == p.T - 1) ? 18 * p.kc_blocks : p.num_k_blocks; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps * CTAS);   // one arrival per epilogue warp of every CTA in the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_cg2(tmem_slot, Cfg::kTmemCols); tmem_relinquish_cg2(); }
    else           { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();   // peer barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  pdl_wait();                          // everything above overlapped the predecessor's last wave; no global memory has been touched yet
  pdl_launch_dependents();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the schedule, one elected lane issues) =====================
    {
      uint32_t it = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int sp = tile / tiles_mn, tmn = tile % tiles_mn;
        const int m_blk = (p.m_blk0 + tmn / p.num_n_blocks) * CTAS + (int)cta_rank;   // this CTA's 128-row block
        const int n_blk = tmn % p.num_n_blocks;
        const int kb_begin = sp * p.kb_per_split, kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
        int f = 0, h0 = 0;
        if (p.conv == 1) {
          f = m_blk / p.tiles_per_frame;
          h0 = (m_blk % p.tiles_per_frame) * p.rows_per_tile;
        }
        int wtap = 0, wc0 = 0;
        if (p.conv == 2) {
          wtap = n_blk / p.cblocks_per_tap;
          wc0 = (n_blk % p.cblocks_per_tap) * BN + (int)cta_rank * Cfg::kBRows;
        }
        const int live_lo = conv_live_lo(p.conv == 1 ? f % p.T : 1), live_hi = conv_live_hi(p.conv == 1 ? f % p.T : 1);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (kb != kb_begin && (kb < live_lo || kb >= live_hi)) continue;
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1u;
          ++it;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const int brow = n_blk * BN + (int)cta_rank * Cfg::kBRows;
          if (!elect_one()) continue;     // uniform-datapath instructions below: see elect_one() in common.cuh
          if (CTAS == 1) {
            mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
            if (p.conv != 1) {
              tma_load_2d(sa, &tmap_a, full_bar(s), kb * BK, m_blk * BM);
            } else {
              const int tap = kb / p.kc_blocks, c0 = (kb % p.kc_blocks) * BK;
              int dt = 0, dh, dw;
              if (p.kt == 3) { dt = tap / 9 - 1; dh = (tap / 3) % 3 - 1; dw = tap % 3 - 1; }
              else           { dh = tap / 3 - 1; dw = tap % 3 - 1; }
              tma_load_5d(sa, &tmap_a, full_bar(s), c0, dw, h0 + dh, (f % p.T) + dt, f / p.T);
            }
            if (p.conv != 2) {
              tma_load_2d(sb, &tmap_b, full_bar(s), kb * BK, brow);
            } else {
              int dt = 0, dh, dw;
              if (p.kt == 3) { dt = wtap / 9 - 1; dh = (wtap / 3) % 3 - 1; dw = wtap % 3 - 1; }
              else           { dh = wtap / 3 - 1; dw = wtap % 3 - 1; }
              const int fr = kb / p.kblocks_per_frame, s0 = (kb % p.kblocks_per_frame) * BK;
              // the w shift selects one of the three pre-shifted planes (outermost coordinate); the h shift moves the origin of the
              // flattened (h,w) axis by whole grid rows (16-byte aligned, zero fill above / below the frame); t shifts its own axis
              tma_load_4d_cta(sb, &tmap_b, full_bar(s), s0 + dh * p.G, (fr % p.T) + dt, fr / p.T, (dw + 1) * p.wg_C + wc0);
            }
          } else {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            if (leader) mbar_expect_tx(full_bar(s), 2 * Cfg::kStageBytes);
            if (p.conv != 1) {
              tma_load_2d_cg2(sa, &tmap_a, full_bar(s), kb * BK, m_blk * BM);
            } else {
              const int tap = kb / p.kc_blocks, c0 = (kb % p.kc_blocks) * BK;
              int dt = 0, dh, dw;
              if (p.kt == 3) { dt = tap / 9 - 1; dh = (tap / 3) % 3 - 1; dw = tap % 3 - 1; }
              else           { dh = tap / 3 - 1; dw = tap % 3 - 1; }
              tma_load_5d_cg2(sa, &tmap_a, full_bar(s), c0, dw, h0 + dh, (f % p.T) + dt, f / p.T);
            }
            if (p.conv != 2) {
              tma_load_2d_cg2(sb, &tmap_b, full_bar(s), kb * BK, brow);
            } else {
              int dt = 0, dh, dw;
              if (p.kt == 3) { dt = wtap / 9 - 1; dh = (wtap / 3) % 3 - 1; dw = wtap % 3 - 1; }
              else           { dh = wtap / 3 - 1; dw = wtap % 3 - 1; }
              const int fr = kb / p.kblocks_per_frame, s0 = (kb % p.kblocks_per_frame) * BK;
              // the w shift selects one of the three pre-shifted planes (outermost coordinate); the h shift moves the origin of the
              // flattened (h,w) axis by whole grid rows (16-byte aligned, zero fill above / below the frame); t shifts its own axis
              tma_load_4d_cg2(sb, &tmap_b, full_bar(s), s0 + dh * p.G, (fr % p.T) + dt, fr / p.T, (dw + 1) * p.wg_C + wc0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CTAS, BN);
      uint32_t it = 0, tile_it = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tile_it) {
        const uint32_t acc = tile_it & 1u, acc_ph = (tile_it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1u);  // epilogues (of both CTAs) have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int sp = tile / tiles_mn;
        const int kb_begin = sp * p.kb_per_split, kb_end = min(p.num_k_blocks, kb_begin + p.kb_per_split);
        const int ft = p.conv == 1 ? (((p.m_blk0 + (tile % tiles_mn) / p.num_n_blocks) * CTAS) / p.tiles_per_frame) % p.T : 1;
        const int live_lo = conv_live_lo(ft), live_hi = conv_live_hi(ft);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (kb != kb_begin && (kb < live_lo || kb >= live_hi)) continue;
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1u;
          ++it;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + s * Cfg::kStageBytes;
            const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = umma_desc_sw128(sa + k * 32);
              const uint64_t db = umma_desc_sw128(sb + k * 32);
              if (CTAS == 2) tc_mma_f16_cg2(d_tmem, da, db, idesc, ((kb - kb_begin) | k) != 0);
              else           tc_mma_f16(d_tmem, da, db, idesc, ((kb - kb_begin) | k) != 0);
            }
            if (CTAS == 2) tc_commit_cg2(empty_bar(s), 3);                      // frees the slot in both CTAs
            else tc_commit(empty_bar(s));
          }
          __syncwarp();
        }
        if (elect_one()) {                                                      // accumulators complete (in both CTAs)
          if (CTAS == 2) tc_commit_cg2(tfull_bar(acc), 3);
          else tc_commit(tfull_bar(acc));
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> smem transpose -> coalesced global =====================
    // 8 warps: warp%4 selects the TMEM lane quadrant (32 accumulator rows), (warp-2)/4 the column half of the tile.
    // Each pass moves a 32-row x 32-column block: one tcgen05.ld, a padded smem transpose, then row-contiguous
    // 128-byte (fp32) / 64-byte (bf16) global segments.  Residual rows are fetched BEFORE the TMEM wait so their
    // latency overlaps the accumulator read.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const float gate = p.gate_alpha ? tanhf(__ldg(p.gate_alpha)) : 1.0f;
    const uint32_t stg = epi_base + (warp - 2) * (32 * Cfg::kStageRowF * 4);
    constexpr int kPasses = BN / 2 / 32;
    uint32_t tile_it = 0;
    if constexpr (EPI == 1 || EPI == 2 || EPI == 5 || EPI == 6 || EPI == 7 || EPI == 8) {
      constexpr bool kFold = EPI == 5 || EPI == 6;         // LayerNorm of the A rows folded in (see GemmParams::ln_stats)
      constexpr bool kGelu = EPI == 2 || EPI == 6;
      constexpr bool kDact = EPI == 7;                     // out = (acc + bias) * gelu'(pre): the fc2 input gradient (backward of the MLP's GELU)
      // EPI = 8: out = bf16(resid_bf16 + gate * relu?(acc + bias)) -- the bf16 residual-stream form of EPI = 4 in the lane = row layout of
      // this block (no fp32 staging, no shuffles for the row statistics: a thread owns its row's 128 columns of the tile)
      constexpr bool kRes = EPI == 8;
      constexpr bool kSide = kDact || kRes;                // a bf16 [M, N] side input fetched coalesced, one pass ahead
      const __nv_bfloat16* const side = kRes ? p.resid16 : p.dact_pre;
      const float2 gate2 = make_float2(gate, gate);
      const float relu_floor = (kRes && p.act == 2) ? 0.f : -INFINITY;
      // ---- bf16 output, bias, optional GELU.  Per warp and pass: one tcgen05.ld of 32 rows x 32 columns (lane = row), bias +
      // activation in registers, pack to bf16, stage 32 rows x 64 B through an XOR-swizzled (conflict-free both ways) private
      // buffer, then 8 rows x 64 B per store instruction.  The accumulator is released right after the last tcgen05.ld.
      const uint32_t bias_s = stg + 2048;                  // this warp's 128 bias values (fp32), 512 B behind the 2 KB staging
      const uint32_t csum_s = stg + 2560;                  // kFold: this warp's 128 column sums of gamma*W
      const uint32_t pre_s = stg + 2560;                   // kDact: this pass' pre-activation block, 32 rows x 64 B, swizzled like the output staging
      __nv_bfloat16* const outp = reinterpret_cast<__nv_bfloat16*>(p.out);
      // kDact: the pre-activation rows arrive like the output leaves (8 rows x 64 B per instruction), one pass ahead of their use
      auto pre_ptr = [&](int tile_, int ps_, int i_) {
        const int m_blk_ = (p.m_blk0 + tile_ / p.num_n_blocks) * CTAS + (int)cta_rank, n_blk_ = tile_ % p.num_n_blocks;
        const int row_ = min(m_blk_ * BM + quad * 32 + 8 * i_ + (lane >> 2), p.M - 1);
        return side + (size_t)row_ * p.N + n_blk_ * BN + half * (BN / 2) + ps_ * 32 + (lane & 3) * 8;
      };
      // fetched TWO passes ahead (a pass is ~1k clk, an HBM access under load ~2k): preg = the coming pass, preg2 = the one after
      uint4 preg[4], preg2[4];
      if (kSide && tile0 < num_tiles) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { preg[i] = *reinterpret_cast<const uint4*>(pre_ptr(tile0, 0, i)); preg2[i] = *reinterpret_cast<const uint4*>(pre_ptr(tile0, 1, i)); }
      }
      static_assert(!kSide || kPasses >= 2, "two-pass look-ahead");
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tile_it) {
        const int m_blk = (p.m_blk0 + tile / p.num_n_blocks) * CTAS + (int)cta_rank, n_blk = tile % p.num_n_blocks;
        const uint32_t acc = tile_it & 1u, acc_ph = (tile_it >> 1) & 1u;
        const int row_base = m_blk * BM + quad * 32;
        const int colw = n_blk * BN + half * (BN / 2);     // first output column of this warp
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + colw) + lane);
        float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kFold) cv = __ldg(reinterpret_cast<const float4*>(p.ln_csum + colw) + lane);
        __syncwarp();                                      // the previous tile's bias reads are done
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(bias_s + lane * 16), "f"(bv.x), "f"(bv.y), "f"(bv.z), "f"(bv.w) : "memory");
        if (kFold) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(csum_s + lane * 16), "f"(cv.x), "f"(cv.y), "f"(cv.z), "f"(cv.w) : "memory");
        float rstd = 1.f, nmr = 0.f;                       // row statistics of this thread's row (TMEM lane = row): rstd and -mu * rstd
        if (kFold) {
          const int row = min(row_base + lane, p.M - 1);
          const float2* sp2 = reinterpret_cast<const float2*>(p.ln_stats) + (size_t)row * p.ln_parts;
          float sx = 0.f, sq = 0.f;
          for (int i = 0; i < p.ln_parts; ++i) { const float2 t2 = sp2[i]; sx += t2.x; sq += t2.y; }
          const float mu = sx * p.ln_inv_k;
          rstd = rsqrtf(fmaxf(sq * p.ln_inv_k - mu * mu, 0.f) + p.ln_eps);
          nmr = -mu * rstd;
        }
        __syncwarp();
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        float st_s = 0.f, st_q = 0.f;                      // kRes + ln_stats_out: sum / sum of squares of this thread's row (rounded values)
#pragma unroll 1
        for (int ps = 0; ps < kPasses; ++ps) {
          if (kSide) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = 8 * i + (lane >> 2);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(pre_s + rl * 64 + (((lane & 3) ^ ((rl >> 1) & 3)) << 4)), "r"(preg[i].x), "r"(preg[i].y),
                           "r"(preg[i].z), "r"(preg[i].w) : "memory");
            }
            const bool wrap = ps + 2 >= kPasses;           // pass ps + 2 belongs to the next tile
            const int ntile = wrap ? tile + tile_step : tile;
#pragma unroll
            for (int i = 0; i < 4; ++i) preg[i] = preg2[i];
            if (ntile < num_tiles) {
#pragma unroll
              for (int i = 0; i < 4; ++i) preg2[i] = *reinterpret_cast<const uint4*>(pre_ptr(ntile, wrap ? ps + 2 - kPasses : ps + 2, i));
            }
          }
          uint32_t r0[32];
          tmem_ld_32x32b_x32(tmem_base + acc * BN + half * (BN / 2) + ps * 32 + ((uint32_t)(quad * 32) << 16), r0);
          tmem_ld_wait();
          uint32_t pu[16];
          if (kSide) {                                     // lane = row: its 32 pre-activations / residual values
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c)
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(pu[4 * c]), "=r"(pu[4 * c + 1]), "=r"(pu[4 * c + 2]), "=r"(pu[4 * c + 3])
                           : "r"(pre_s + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)));
          }
          if (ps == kPasses - 1) {                         // accumulator drained: the MMA warp may start the tile after next
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CTAS == 2) mbar_arrive_cluster_relaxed(tempty_bar(acc), 0);
              else mbar_arrive(tempty_bar(acc));
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(bias_s + (ps * 32 + j) * 4));
            // packed fp32 (FFMA2 / FADD2): the epilogue sits on the accumulator-release path of these short-K tiles
            float2 v01 = make_float2(__uint_as_float(r0[j]), __uint_as_float(r0[j + 1]));
            float2 v23 = make_float2(__uint_as_float(r0[j + 2]), __uint_as_float(r0[j + 3]));
            if (kFold) {
              float4 c4;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c4.x), "=f"(c4.y), "=f"(c4.z), "=f"(c4.w) : "r"(csum_s + (ps * 32 + j) * 4));
              const float2 nm2 = make_float2(nmr, nmr), rs2 = make_float2(rstd, rstd);
              v01 = __ffma2_rn(v01, rs2, __ffma2_rn(nm2, make_float2(c4.x, c4.y), make_float2(b4.x, b4.y)));
              v23 = __ffma2_rn(v23, rs2, __ffma2_rn(nm2, make_float2(c4.z, c4.w), make_float2(b4.z, b4.w)));
            } else {
              v01 = __fadd2_rn(v01, make_float2(b4.x, b4.y));
              v23 = __fadd2_rn(v23, make_float2(b4.z, b4.w));
            }
            if (kGelu) { v01 = gelu_fast2(v01); v23 = gelu_fast2(v23); }
            if (kDact) {
              v01 = __fmul2_rn(v01, gelu_grad_fast2(unpack_bf16(pu[j >> 1])));
              v23 = __fmul2_rn(v23, gelu_grad_fast2(unpack_bf16(pu[(j >> 1) + 1])));
            }
            if (kRes) {
              v01 = make_float2(fmaxf(v01.x, relu_floor), fmaxf(v01.y, relu_floor));
              v23 = make_float2(fmaxf(v23.x, relu_floor), fmaxf(v23.y, relu_floor));
              v01 = __ffma2_rn(v01, gate2, unpack_bf16(pu[j >> 1]));
              v23 = __ffma2_rn(v23, gate2, unpack_bf16(pu[(j >> 1) + 1]));
            }
            const float v0 = v01.x, v1 = v01.y, v2 = v23.x, v3 = v23.y;
            pk[j >> 1] = pack_bf16(v0, v1);
            pk[(j >> 1) + 1] = pack_bf16(v2, v3);
            if (kRes) {                                    // statistics of the ROUNDED values: what the consuming GEMM reads
              const float2 q01 = unpack_bf16(pk[j >> 1]), q23 = unpack_bf16(pk[(j >> 1) + 1]);
              st_s += (q01.x + q01.y) + (q23.x + q23.y);
              st_q = fmaf(q01.x, q01.x, fmaf(q01.y, q01.y, fmaf(q23.x, q23.x, fmaf(q23.y, q23.y, st_q))));
            }
          }
          // lane = row: four 16-byte chunks (8 columns each), chunk c stored at position c ^ ((row >> 1) & 3) of the 64-byte row
          const uint32_t wrow = stg + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(wrow + ((c ^ sw) << 4)), "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]),
                         "r"(pk[4 * c + 3]) : "memory");
          __syncwarp();
          const int col0 = colw + ps * 32 + (lane & 3) * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = 8 * i + (lane >> 2);
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(stg + rl * 64 + (((lane & 3) ^ ((rl >> 1) & 3)) << 4)));
            if (row_base + rl < p.M) *reinterpret_cast<uint4*>(outp + (size_t)(row_base + rl) * p.N + col0) = o;
          }
          __syncwarp();                                    // staging is rewritten by the next pass
        }
        if (kRes && p.ln_stats_out && row_base + lane < p.M)
          reinterpret_cast<float2*>(p.ln_stats_out)[(size_t)(row_base + lane) * (2 * p.num_n_blocks) + n_blk * 2 + half] = make_float2(st_s, st_q);
      }
    } else if constexpr (EPI == 3) {
      // ---- fp32 output = resid + gate * act(acc + bias) (+ bf16 copy): the residual-stream GEMMs (proj, fc2, Conv3d adapter).
      // These tiles move 256 KB of fp32 per CTA through 8 warps; the residual rows are fetched one pass ahead (and across the
      // tile boundary) so their latency overlaps the previous pass.  (Measured: an extra L2 prefetch a tile ahead and
      // L1::no_allocate loads change nothing -- the path sits at ~3.3 TB/s of combined traffic per kernel.)
      const float relu_floor = p.act == 2 ? 0.f : -INFINITY;
      float* const outp = reinterpret_cast<float*>(p.out);
      const int c = (lane & 7) * 4;                        // lane -> (row 4i + lane/8, 4 columns at c)
      auto res_ptr = [&](int tile_, int ps_, int i_) {
        const int m_blk_ = (p.m_blk0 + tile_ / p.num_n_blocks) * CTAS + (int)cta_rank, n_blk_ = tile_ % p.num_n_blocks;
        const int row_ = m_blk_ * BM + quad * 32 + 4 * i_ + (lane >> 3);
        const int rr_ = min(row_, p.M - 1);
        return p.resid + (size_t)(p.resid_mod > 0 ? rr_ % p.resid_mod : rr_) * p.N + n_blk_ * BN + half * (BN / 2) + ps_ * 32 + c;
      };
      float4 res[8];
      if (tile0 < num_tiles) {
#pragma unroll
        for (int i = 0; i < 8; ++i) res[i] = *reinterpret_cast<const float4*>(res_ptr(tile0, 0, i));
      }
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tile_it) {
        const int m_blk = (p.m_blk0 + tile / p.num_n_blocks) * CTAS + (int)cta_rank, n_blk = tile % p.num_n_blocks;
        const uint32_t acc = tile_it & 1u, acc_ph = (tile_it >> 1) & 1u;
        const int row_base = m_blk * BM + quad * 32;
        const int colw = n_blk * BN + half * (BN / 2);
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
#pragma unroll 1
        for (int ps = 0; ps < kPasses; ++ps) {
          const int col0 = colw + ps * 32 + c;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0));
          {
            uint32_t r0[32];
            tmem_ld_32x32b_x32(tmem_base + acc * BN + half * (BN / 2) + ps * 32 + ((uint32_t)(quad * 32) << 16), r0);
            tmem_ld_wait();
            if (ps == kPasses - 1) {                       // accumulator drained
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CTAS == 2) mbar_arrive_cluster_relaxed(tempty_bar(acc), 0);
                else mbar_arrive(tempty_bar(acc));
              }
            }
            const uint32_t wrow = stg + lane * (Cfg::kStageRowF * 4);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(wrow + j * 4), "r"(r0[j]), "r"(r0[j + 1]), "r"(r0[j + 2]), "r"(r0[j + 3]) : "memory");
          }
          __syncwarp();
          // next pass's residual rows (next tile's first pass after the last one) go in flight before this pass is consumed
          float4 nres[8];
          const bool last = ps == kPasses - 1;
          const int ntile = last ? tile + tile_step : tile;
          const bool more = ntile < num_tiles;
          if (more) {
#pragma unroll
            for (int i = 0; i < 8; ++i) nres[i] = *reinterpret_cast<const float4*>(res_ptr(ntile, last ? 0 : ps + 1, i));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + (lane >> 3);
            const int row = row_base + rl;
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(stg + (rl * Cfg::kStageRowF + c) * 4));
            v.x = fmaxf(v.x + bias4.x, relu_floor); v.y = fmaxf(v.y + bias4.y, relu_floor);
            v.z = fmaxf(v.z + bias4.z, relu_floor); v.w = fmaxf(v.w + bias4.w, relu_floor);
            v.x = fmaf(v.x, gate, res[i].x); v.y = fmaf(v.y, gate, res[i].y); v.z = fmaf(v.z, gate, res[i].z); v.w = fmaf(v.w, gate, res[i].w);
            if (row < p.M) {
              const size_t o = (size_t)row * p.N + col0;
              *reinterpret_cast<float4*>(outp + o) = v;
              if (p.out2) *reinterpret_cast<uint2*>(p.out2 + o) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            }
          }
          if (more) {
#pragma unroll
            for (int i = 0; i < 8; ++i) res[i] = nres[i];
          }
          __syncwarp();                                    // staging is rewritten by the next pass
        }
      }
    } else if constexpr (EPI == 4) {
      // ---- bf16 output = bf16(resid_bf16 + gate * act(acc + bias)): the residual-stream GEMMs with the stream kept in bf16 (half the
      // epilogue traffic of EPI = 3 and no separate bf16 copy for the next tensor-core operand).  Same structure: fp32 staging through the
      // padded transpose buffer, residual rows fetched one pass ahead, sum formed in fp32 and rounded once.
      const float relu_floor = p.act == 2 ? 0.f : -INFINITY;
      __nv_bfloat16* const outp = reinterpret_cast<__nv_bfloat16*>(p.out);
      const int c = (lane & 7) * 4;                        // lane -> (row 4i + lane/8, 4 columns at c)
      auto res_ptr = [&](int tile_, int ps_, int i_) {
        const int m_blk_ = (p.m_blk0 + tile_ / p.num_n_blocks) * CTAS + (int)cta_rank, n_blk_ = tile_ % p.num_n_blocks;
        const int row_ = m_blk_ * BM + quad * 32 + 4 * i_ + (lane >> 3);
        const int rr_ = min(row_, p.M - 1);
        return p.resid16 + (size_t)rr_ * p.N + n_blk_ * BN + half * (BN / 2) + ps_ * 32 + c;
      };
      uint2 res[8];
      if (tile0 < num_tiles) {
#pragma unroll
        for (int i = 0; i < 8; ++i) res[i] = *reinterpret_cast<const uint2*>(res_ptr(tile0, 0, i));
      }
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tile_it) {
        const int m_blk = (p.m_blk0 + tile / p.num_n_blocks) * CTAS + (int)cta_rank, n_blk = tile % p.num_n_blocks;
        const uint32_t acc = tile_it & 1u, acc_ph = (tile_it >> 1) & 1u;
        const int row_base = m_blk * BM + quad * 32;
        const int colw = n_blk * BN + half * (BN / 2);
        mbar_wait(tfull_bar(acc), acc_ph);
        tc_fence_after();
        float rs[8], rq[8];                                // ln_stats_out: sum / sum of squares of this thread's 8 rows (4 columns per pass)
#pragma unroll
        for (int i = 0; i < 8; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
#pragma unroll 1
        for (int ps = 0; ps < kPasses; ++ps) {
          const int col0 = colw + ps * 32 + c;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0));
          {
            uint32_t r0[32];
            tmem_ld_32x32b_x32(tmem_base + acc * BN + half * (BN / 2) + ps * 32 + ((uint32_t)(quad * 32) << 16), r0);
            tmem_ld_wait();
            if (ps == kPasses - 1) {                       // accumulator drained
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CTAS == 2) mbar_arrive_cluster_relaxed(tempty_bar(acc), 0);
                else mbar_arrive(tempty_bar(acc));
              }
            }
            const uint32_t wrow = stg + lane * (Cfg::kStageRowF * 4);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(wrow + j * 4), "r"(r0[j]), "r"(r0[j + 1]), "r"(r0[j + 2]), "r"(r0[j + 3]) : "memory");
          }
          __syncwarp();
          // next pass's residual rows (next tile's first pass after the last one) go in flight before this pass is consumed
          uint2 nres[8];
          const bool last = ps == kPasses - 1;
          const int ntile = last ? tile + tile_step : tile;
          const bool more = ntile < num_tiles;
          if (more) {
#pragma unroll
            for (int i = 0; i < 8; ++i) nres[i] = *reinterpret_cast<const uint2*>(res_ptr(ntile, last ? 0 : ps + 1, i));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + (lane >> 3);
            const int row = row_base + rl;
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(stg + (rl * Cfg::kStageRowF + c) * 4));
            v.x = fmaxf(v.x + bias4.x, relu_floor); v.y = fmaxf(v.y + bias4.y, relu_floor);
            v.z = fmaxf(v.z + bias4.z, relu_floor); v.w = fmaxf(v.w + bias4.w, relu_floor);
            const float2 r01 = unpack_bf16(res[i].x), r23 = unpack_bf16(res[i].y);
            v.x = fmaf(v.x, gate, r01.x); v.y = fmaf(v.y, gate, r01.y); v.z = fmaf(v.z, gate, r23.x); v.w = fmaf(v.w, gate, r23.y);
            const uint2 pk2 = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            if (row < p.M) *reinterpret_cast<uint2*>(outp + (size_t)row * p.N + col0) = pk2;
            if (p.ln_stats_out) {                          // statistics of the ROUNDED values: what the consuming GEMM reads
              const float2 q01 = unpack_bf16(pk2.x), q23 = unpack_bf16(pk2.y);
              rs[i] += (q01.x + q01.y) + (q23.x + q23.y);
              rq[i] += (q01.x * q01.x + q01.y * q01.y) + (q23.x * q23.x + q23.y * q23.y);
            }
          }
          if (more) {
#pragma unroll
            for (int i = 0; i < 8; ++i) res[i] = nres[i];
          }
          __syncwarp();                                    // staging is rewritten by the next pass
        }
        if (p.ln_stats_out) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o); rq[i] += __shfl_xor_sync(0xffffffffu, rq[i], o); }
            const int row = row_base + 4 * i + (lane >> 3);
            if ((lane & 7) == 0 && row < p.M)
              reinterpret_cast<float2*>(p.ln_stats_out)[(size_t)row * (2 * p.num_n_blocks) + n_blk * 2 + half] = make_float2(rs[i], rq[i]);
          }
        }
      }
    } else
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tile_it) {
      const int sp = tile / tiles_mn, tmn = tile % tiles_mn;
      const int m_blk = (p.m_blk0 + tmn / p.num_n_blocks) * CTAS + (int)cta_rank, n_blk = tmn % p.num_n_blocks;
      const uint32_t acc = tile_it & 1u, acc_ph = (tile_it >> 1) & 1u;
      const int row_base = m_blk * BM + quad * 32;
      // split-K: partial plane `sp` of this launch's row range; otherwise the output itself
      float* const out_f = reinterpret_cast<float*>(p.out) + (p.splits > 1 ? ((long long)sp * p.plane_rows - (long long)p.m_blk0 * CTAS * BM) * p.N : 0ll);
      bool waited = false;
#pragma unroll 1
      for (int ps = 0; ps < kPasses; ++ps) {
        const int ccol = half * (BN / 2) + ps * 32;   // column offset inside the tile
        const int col0 = n_blk * BN + ccol;
        if (p.out_f32) {
          // lane -> (row 4i + lane/8, 4 columns at (lane%8)*4)
          const int c = (lane & 7) * 4;
          float4 res[8];
          if (p.resid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = row_base + 4 * i + (lane >> 3);
              res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < p.M) res[i] = *reinterpret_cast<const float4*>(p.resid + (size_t)(p.resid_mod > 0 ? row % p.resid_mod : row) * p.N + col0 + c);
            }
          }
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c));
          if (!waited) { mbar_wait(tfull_bar(acc), acc_ph); tc_fence_after(); waited = true; }
          {
            uint32_t r0[32];
            tmem_ld_32x32b_x32(tmem_base + acc * BN + ccol + ((uint32_t)(quad * 32) << 16), r0);
            tmem_ld_wait();
            const uint32_t wrow = stg + lane * (Cfg::kStageRowF * 4);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(wrow + j * 4), "r"(r0[j]), "r"(r0[j + 1]), "r"(r0[j + 2]), "r"(r0[j + 3]) : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + (lane >> 3);
            const int row = row_base + rl;
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(stg + (rl * Cfg::kStageRowF + c) * 4));
            v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
            if (p.act == 1) { v.x = gelu_fast(v.x); v.y = gelu_fast(v.y); v.z = gelu_fast(v.z); v.w = gelu_fast(v.w); }
            else if (p.act == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (p.out2_pre && p.out2 && row < p.M)   // value after the activation, before gate / residual (saved for the backward pass)
              *reinterpret_cast<uint2*>(p.out2 + (size_t)row * p.N + col0 + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            if (p.gate_alpha) { v.x *= gate; v.y *= gate; v.z *= gate; v.w *= gate; }
            if (p.resid) { v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w; }
            if (row < p.M) {
              const size_t o = (size_t)row * p.N + col0 + c;
              *reinterpret_cast<float4*>(out_f + o) = v;
              if (p.out2 && !p.out2_pre) *reinterpret_cast<uint2*>(p.out2 + o) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            }
          }
        } else {
          // lane -> (row 8i + lane/4, 8 columns at (lane%4)*8)
          const int c = (lane & 3) * 8;
          float4 res[8];
          if (p.resid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = row_base + 8 * i + (lane >> 2);
              res[2 * i] = res[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < p.M) {
                const float* rp = p.resid + (size_t)(p.resid_mod > 0 ? row % p.resid_mod : row) * p.N + col0 + c;
                res[2 * i] = *reinterpret_cast<const float4*>(rp);
                res[2 * i + 1] = *reinterpret_cast<const float4*>(rp + 4);
              }
            }
          } else if (p.resid16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = row_base + 8 * i + (lane >> 2);
              res[2 * i] = res[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < p.M) {
                const uint4 u = *reinterpret_cast<const uint4*>(p.resid16 + (size_t)row * p.N + col0 + c);
                const float2 a = unpack_bf16(u.x), b2 = unpack_bf16(u.y), c2 = unpack_bf16(u.z), d2 = unpack_bf16(u.w);
                res[2 * i] = make_float4(a.x, a.y, b2.x, b2.y);
                res[2 * i + 1] = make_float4(c2.x, c2.y, d2.x, d2.y);
              }
            }
          }
          uint4 pre[4];
          if (p.dact_pre) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = row_base + 8 * i + (lane >> 2);
              pre[i] = make_uint4(0, 0, 0, 0);
              if (row < p.M) pre[i] = *reinterpret_cast<const uint4*>(p.dact_pre + (size_t)row * p.N + col0 + c);
            }
          }
          float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
          if (p.bias) {
            ba = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c));
            bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c + 4));
          }
          if (!waited) { mbar_wait(tfull_bar(acc), acc_ph); tc_fence_after(); waited = true; }
          {
            uint32_t r0[32];
            tmem_ld_32x32b_x32(tmem_base + acc * BN + ccol + ((uint32_t)(quad * 32) << 16), r0);
            tmem_ld_wait();
            const uint32_t wrow = stg + lane * (Cfg::kStageRowF * 4);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(wrow + j * 4), "r"(r0[j]), "r"(r0[j + 1]), "r"(r0[j + 2]), "r"(r0[j + 3]) : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = 8 * i + (lane >> 2);
            const int row = row_base + rl;
            float4 u, w;
            const uint32_t a = stg + (rl * Cfg::kStageRowF + c) * 4;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(u.x), "=f"(u.y), "=f"(u.z), "=f"(u.w) : "r"(a));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(a + 16));
            float v[8] = {u.x + ba.x, u.y + ba.y, u.z + ba.z, u.w + ba.w, w.x + bb.x, w.y + bb.y, w.z + bb.z, w.w + bb.w};
            if (p.out2_pre && p.out2 && row < p.M)
              *reinterpret_cast<uint4*>(p.out2 + (size_t)row * p.N + col0 + c) =
                  make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = gelu_fast(v[j]);
            } else if (p.act == 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.gate_alpha) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= gate;
            }
            if (p.dact_pre) {
              const uint32_t pu[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 z = unpack_bf16(pu[j]);
                if (p.dact == 1) { const float2 g = gelu_grad_fast2(z); v[2 * j] *= g.x; v[2 * j + 1] *= g.y; }
                else             { v[2 * j] = z.x > 0.f ? v[2 * j] : 0.f; v[2 * j + 1] = z.y > 0.f ? v[2 * j + 1] : 0.f; }
              }
            }
            if (p.resid || p.resid16) {
              v[0] += res[2 * i].x; v[1] += res[2 * i].y; v[2] += res[2 * i].z; v[3] += res[2 * i].w;
              v[4] += res[2 * i + 1].x; v[5] += res[2 * i + 1].y; v[6] += res[2 * i + 1].z; v[7] += res[2 * i + 1].w;
            }
            if (row < p.M) {
              const size_t o = (size_t)row * p.N + col0 + c;
              const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = pk;
              if (p.out2 && !p.out2_pre) *reinterpret_cast<uint4*>(p.out2 + o) = pk;
            }
          }
        }
        __syncwarp();   // staging buffer is reused by the next pass
      }
      // accumulator drained (all tcgen05.ld of this warp completed before the smem staging): release it to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster_relaxed(tempty_bar(acc), 0);
        else mbar_arrive(tempty_bar(acc));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();   // the peer may still multicast into / read from this CTA's shared memory
  if (warp == 1) {
    if (CTAS == 2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 tensor, innermost dimension contiguous; dims/box listed innermost first.  128-byte swizzle, zero OOB fill.
static int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box,
                          CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { grove_set_error("cuTensorMapEncodeTiled entry point not available"); return GROVE_ERR_CUDA; }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { grove_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank); return GROVE_ERR_CUDA; }
  return GROVE_OK;
}

// box_inner * 2 bytes must equal the swizzle span (64 elements -> 128B swizzle, 16 elements -> 32B swizzle)
int make_tmap_bf16_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
  uint64_t d[2] = {inner, rows};
  uint32_t b[2] = {box_inner, box_rows};
  return make_tmap_bf16(m, base, 2, d, b, box_inner == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_bf16_nd(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
  return make_tmap_bf16(m, base, rank, dims, box, box[0] == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

static int g_num_sms[64] = {0};   // per device ordinal
static int num_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = g_num_sms[dev & 63];
  if (!n) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n = v > 0 ? v : kNumSMs;
  }
  return n;
}

template <int BN, int CTAS, int EPI = 0>
static int launch_gemm(const GemmParams& p, const CUtensorMap& ta, const CUtensorMap& tb, int max_ctas, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CTAS>;
  static GrovePerDeviceOnce attr_set;
  if (attr_set.first_time()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, CTAS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      attr_set.reset_current();
      grove_set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return GROVE_ERR_CUDA;
    }
  }
  int grid = p.num_m_blocks * p.num_n_blocks * p.splits * CTAS;
  int cap = max_ctas > 0 ? max_ctas : num_sms();
  cap -= cap % CTAS;
  if (cap < CTAS) cap = CTAS;
  if (grid > cap) grid = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = grove_pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, CTAS, EPI>, p, ta, tb);
  grove_count_launch();
  if (e != cudaSuccess) { grove_set_error("gemm launch failed: %s", cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

static bool epi4_staged() {   // GROVE_GEMM_EPI4=1: the fp32-staged bf16 residual epilogue (EPI = 4) instead of the lane = row form (EPI = 8); A/B measurements
  static const bool v = []() { const char* e = getenv("GROVE_GEMM_EPI4"); return e && e[0] == '1'; }();
  return v;
}
static bool tail_split_disabled() {   // GROVE_GEMM_NO_TAIL_SPLIT=1 (A/B measurements)
  static int v = -1;
  if (v < 0) v = getenv("GROVE_GEMM_NO_TAIL_SPLIT") != nullptr;
  return v != 0;
}

static bool generic_epilogue_only() {   // GROVE_GEMM_GENERIC_EPI=1: route everything through the general epilogue (A/B measurements)
  static int v = -1;
  if (v < 0) v = getenv("GROVE_GEMM_GENERIC_EPI") != nullptr;
  return v != 0;
}

// Tail fix-up of a wave-quantised GEMM: out[r, n] = resid[r, n] + gate * act(sum_s part[s, r - row0, n] + bias[n]) (+ bf16 copy)
// for the rows [row0, row1) whose tiles were computed as split-K partial planes (see dispatch_gemm).
__global__ void __launch_bounds__(256) gemm_tail_fixup_kernel(const float* __restrict__ part, int splits, int plane_rows, int row0, int row1, int N,
                                                              const float* __restrict__ bias, const float* resid, const float* __restrict__ gate_alpha,
                                                              int act, float* out, __nv_bfloat16* out2, const __nv_bfloat16* resid16 = nullptr) {
  const float gate = gate_alpha ? tanhf(__ldg(gate_alpha)) : 1.0f;
  const float floor_v = act == 2 ? 0.f : -INFINITY;
  const int n4 = N / 4;
  const long long total = (long long)(row1 - row0) * n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4), c = (int)(i % n4) * 4;
    float4 a = *reinterpret_cast<const float4*>(part + (size_t)r * N + c);
    for (int s = 1; s < splits; ++s) {
      const float4 b = *reinterpret_cast<const float4*>(part + ((size_t)s * plane_rows + r) * N + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + c));
    const size_t o = (size_t)(row0 + r) * N + c;
    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (resid) rv = *reinterpret_cast<const float4*>(resid + o);
    else if (resid16) {
      const uint2 u = *reinterpret_cast<const uint2*>(resid16 + o);
      const float2 r01 = unpack_bf16(u.x), r23 = unpack_bf16(u.y);
      rv = make_float4(r01.x, r01.y, r23.x, r23.y);
    }
    float4 v;
    v.x = fmaf(fmaxf(a.x + bv.x, floor_v), gate, rv.x); v.y = fmaf(fmaxf(a.y + bv.y, floor_v), gate, rv.y);
    v.z = fmaf(fmaxf(a.z + bv.z, floor_v), gate, rv.z); v.w = fmaf(fmaxf(a.w + bv.w, floor_v), gate, rv.w);
    if (out) *reinterpret_cast<float4*>(out + o) = v;
    if (out2) *reinterpret_cast<uint2*>(out2 + o) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// tile configuration: CTA pairs (256 x 256 tiles) whenever the problem has them, else single-CTA 128 x {256,128} tiles
static int dispatch_gemm(GemmParams& p, const void* A_or_X, const void* W, int K_total, int N, int M, bool conv, const uint64_t* adims,
                         const uint32_t* abox, int arank, int max_ctas, int force_ctas, cudaStream_t st, const uint64_t* bdims5 = nullptr,
                         void* workspace = nullptr, size_t workspace_bytes = 0) {
  const int BN = (N % 256 == 0) ? 256 : 128;
  int ctas = (BN == 256 && M >= 256) ? 2 : 1;
  if (force_ctas == 1 || force_ctas == 2) ctas = (force_ctas == 2 && BN == 256) ? 2 : 1;
  p.num_m_blocks = (M + BM * ctas - 1) / (BM * ctas);
  p.num_n_blocks = N / BN;
  if (p.splits < 1) p.splits = 1;
  p.kb_per_split = (p.num_k_blocks + p.splits - 1) / p.splits;
  p.splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;   // no empty split
  p.m_blk0 = 0;
  p.plane_rows = p.M;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_bf16(&ta, A_or_X, arank, adims, abox))) return rc;
  if (p.conv == 2) {
    // B = X^T in three w-shifted planes [3*C, V, T, G*G] (channel-major); a 64-token k-block is 64 consecutive (h,w) positions of a frame
    p.cblocks_per_tap = p.wg_C / BN;
    uint32_t bb4[4] = {(uint32_t)BK, 1, 1, (uint32_t)(BN / ctas)};
    if ((rc = make_tmap_bf16(&tb, W, 4, bdims5, bb4))) return rc;
  } else {
    uint64_t db[2] = {(uint64_t)K_total, (uint64_t)N};
    uint32_t bb[2] = {BK, (uint32_t)(BN / ctas)};
    if ((rc = make_tmap_bf16(&tb, W, 2, db, bb))) return rc;
  }
  (void)conv;
  // Few tiles, long K (the text projection: 4 [DET] rows x 4096 x 4096 is 16 tiles, then ONE tile, each streaming its weights
  // through a single CTA at ~0.9 TB/s): spread K over the idle SMs as split-K partial planes and finish with the fix-up kernel.
  {
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    const bool simple = p.splits == 1 && p.conv == 0 && !p.resid && !p.resid16 && !p.gate_alpha && !p.dact_pre && !p.out2 && p.act != 1;
    if (simple && workspace && tiles * ctas * 2 <= num_sms() && p.num_k_blocks >= 32 && max_ctas == 0 && !tail_split_disabled()) {
      int s = num_sms() / (tiles * ctas);
      if (s > 16) s = 16;
      while (s > 1 && p.num_k_blocks / s < 4) --s;
      const int plane_rows = p.num_m_blocks * BM * ctas;
      if (s >= 2 && (size_t)s * plane_rows * p.N * sizeof(float) <= workspace_bytes) {
        GemmParams pb = p;
        pb.splits = s; pb.kb_per_split = (p.num_k_blocks + s - 1) / s;
        pb.splits = (p.num_k_blocks + pb.kb_per_split - 1) / pb.kb_per_split;
        pb.plane_rows = plane_rows;
        pb.out = workspace; pb.out_f32 = 1; pb.bias = nullptr; pb.act = 0;
        int rc2 = ctas == 2 ? launch_gemm<256, 2>(pb, ta, tb, max_ctas, st)
                            : (BN == 256 ? launch_gemm<256, 1>(pb, ta, tb, max_ctas, st) : launch_gemm<128, 1>(pb, ta, tb, max_ctas, st));
        if (rc2) return rc2;
        const long long n4 = (long long)p.M * (p.N / 4);
        const int grid = (int)((n4 + 255) / 256 < 4 * num_sms() ? (n4 + 255) / 256 : 4 * num_sms());
        gemm_tail_fixup_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(workspace), pb.splits, plane_rows, 0, p.M, p.N, p.bias, nullptr, nullptr,
                                                     p.act, p.out_f32 ? reinterpret_cast<float*>(p.out) : nullptr,
                                                     p.out_f32 ? nullptr : reinterpret_cast<__nv_bfloat16*>(p.out));
        grove_count_launch();
        GROVE_CHECK_LAUNCH();
        return GROVE_OK;
      }
    }
  }
  // straight-line epilogues for the two hot bf16-output forms (bias, optional GELU, nothing else)
  const bool plain_bf16 = !p.out_f32 && !p.resid && !p.resid16 && !p.gate_alpha && !p.out2 && !p.dact_pre && p.splits == 1 && p.conv != 2 && (p.act == 0 || p.act == 1) &&
                          !generic_epilogue_only();
  if (p.ln_stats && !(ctas == 2 && plain_bf16)) {
    grove_set_error("the LayerNorm-folded epilogue needs a plain bf16-output GEMM on CTA-pair tiles (N %% 256 == 0, M >= 256)");
    return GROVE_ERR_UNSUPPORTED;
  }
  if (ctas == 2 && plain_bf16 && p.ln_stats)
    return p.act == 1 ? launch_gemm<256, 2, 6>(p, ta, tb, max_ctas, st) : launch_gemm<256, 2, 5>(p, ta, tb, max_ctas, st);
  if (ctas == 2 && plain_bf16) return p.act == 1 ? launch_gemm<256, 2, 2>(p, ta, tb, max_ctas, st) : launch_gemm<256, 2, 1>(p, ta, tb, max_ctas, st);
  // the fc2 input gradient: bf16 output scaled by the GELU derivative of the saved pre-activation
  const bool dact_bf16 = !p.out_f32 && !p.resid && !p.resid16 && !p.gate_alpha && !p.out2 && p.dact_pre && p.dact == 1 && p.splits == 1 && p.conv != 2 && p.act == 0 &&
                         !generic_epilogue_only();
  if (ctas == 2 && dact_bf16) return launch_gemm<256, 2, 7>(p, ta, tb, max_ctas, st);
  const bool resid_f32 = p.out_f32 && p.resid && !p.resid16 && !p.dact_pre && p.splits == 1 && p.conv != 2 && p.act != 1 && !p.out2_pre &&
                         !generic_epilogue_only();
  const bool resid_b16 = !p.out_f32 && p.resid16 && !p.resid && !p.dact_pre && p.splits == 1 && p.conv != 2 && p.act != 1 && !p.out2 &&
                         !generic_epilogue_only();
  if (p.ln_stats_out && !(ctas == 2 && resid_b16)) {
    grove_set_error("row statistics are produced by the bf16 residual-stream epilogue on CTA-pair tiles only");
    return GROVE_ERR_UNSUPPORTED;
  }
  if (ctas == 2 && (resid_f32 || resid_b16)) {
    // Wave quantisation: 256x256 tiles on P = 74 CTA pairs.  With N = 768 and M = 32768 (proj, fc2, the Conv3d adapter) there are
    // 384 tiles = 5.19 waves, so the last wave runs 14 tiles on 74 pairs.  When the remainder is small the trailing m-blocks are
    // computed instead as split-K partial planes (t*nb tiles x s splits <= P units of K/s k-blocks each, one short wave) and a
    // fix-up kernel applies the epilogue to those rows: 5 + 1/s waves instead of 6.
    const int pairs = (max_ctas > 0 ? max_ctas : num_sms()) / 2;
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    const int full_waves = pairs > 0 ? tiles / pairs : 0, rem = pairs > 0 ? tiles % pairs : 0;
    // (measured: pays off for the K = 20736 adapter conv, 858 -> 778 us; LOSES on fc2 with K = 3072, 124 -> 134 us, where the two extra
    // launches and the partial-plane traffic outweigh 0.75 of a 20 us wave -- hence the k-block threshold)
    if (workspace && full_waves >= 1 && rem > 0 && 2 * rem <= pairs && p.M % (BM * 2) == 0 && p.num_k_blocks >= 96 && p.resid_mod == 0 && !p.ln_stats_out && !tail_split_disabled()) {
      const int t = (rem + p.num_n_blocks - 1) / p.num_n_blocks;            // trailing m-blocks taken out of the main launch
      const int tail_tiles = t * p.num_n_blocks;
      int s = tail_tiles > 0 ? pairs / tail_tiles : 0;
      if (s > 8) s = 8;
      while (s >= 2 && p.num_k_blocks / s < 8) --s;                        // keep every unit at least 8 k-blocks long
      const size_t need = (size_t)s * t * BM * 2 * p.N * sizeof(float);
      if (s >= 2 && t < p.num_m_blocks && need <= workspace_bytes) {
        GemmParams pa = p;
        pa.num_m_blocks = p.num_m_blocks - t;
        int rc = resid_b16 ? (epi4_staged() ? launch_gemm<256, 2, 4>(pa, ta, tb, max_ctas, st) : launch_gemm<256, 2, 8>(pa, ta, tb, max_ctas, st))
                           : launch_gemm<256, 2, 3>(pa, ta, tb, max_ctas, st);
        if (rc) return rc;
        GemmParams pb = p;
        pb.m_blk0 = p.num_m_blocks - t; pb.num_m_blocks = t;
        pb.splits = s; pb.kb_per_split = (p.num_k_blocks + s - 1) / s;
        pb.splits = (p.num_k_blocks + pb.kb_per_split - 1) / pb.kb_per_split;
        pb.plane_rows = t * BM * 2;
        pb.out = workspace; pb.out_f32 = 1;
        pb.bias = nullptr; pb.resid = nullptr; pb.resid16 = nullptr; pb.gate_alpha = nullptr; pb.act = 0; pb.out2 = nullptr;
        rc = launch_gemm<256, 2>(pb, ta, tb, max_ctas, st);
        if (rc) return rc;
        const int row0 = pb.m_blk0 * BM * 2;
        const long long n4 = (long long)(p.M - row0) * (p.N / 4);
        const int grid = (int)((n4 + 255) / 256 < 4 * num_sms() ? (n4 + 255) / 256 : 4 * num_sms());
        if (resid_b16)
          gemm_tail_fixup_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(workspace), pb.splits, pb.plane_rows, row0, p.M, p.N, p.bias, nullptr,
                                                       p.gate_alpha, p.act, nullptr, reinterpret_cast<__nv_bfloat16*>(p.out), p.resid16);
        else
          gemm_tail_fixup_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(workspace), pb.splits, pb.plane_rows, row0, p.M, p.N, p.bias, p.resid,
                                                       p.gate_alpha, p.act, reinterpret_cast<float*>(p.out), p.out2);
        grove_count_launch();
        GROVE_CHECK_LAUNCH();
        return GROVE_OK;
      }
    }
    if (resid_b16) return epi4_staged() ? launch_gemm<256, 2, 4>(p, ta, tb, max_ctas, st) : launch_gemm<256, 2, 8>(p, ta, tb, max_ctas, st);
    return launch_gemm<256, 2, 3>(p, ta, tb, max_ctas, st);
  }
  if (ctas == 2) return launch_gemm<256, 2>(p, ta, tb, max_ctas, st);
  return BN == 256 ? launch_gemm<256, 1>(p, ta, tb, max_ctas, st) : launch_gemm<128, 1>(p, ta, tb, max_ctas, st);
}

}  // namespace grove

using namespace grove;

static int fill_epilogue(GemmParams& p, const grove_gemm_epilogue* e, int M, int N) {
  p.bias = e ? e->bias : nullptr;
  p.resid = e ? e->resid : nullptr;
  p.resid16 = e ? reinterpret_cast<const __nv_bfloat16*>(e->resid_bf16) : nullptr;
  p.ln_stats = e ? e->ln_stats : nullptr;
  p.ln_csum = e ? e->ln_colsum : nullptr;
  p.ln_parts = e ? e->ln_parts : 0;
  p.ln_eps = e ? e->ln_eps : 0.f;
  p.ln_inv_k = 0.f;
  p.ln_stats_out = e ? e->ln_stats_out : nullptr;
  p.resid_mod = e ? e->resid_row_mod : 0;
  p.gate_alpha = e ? e->gate_alpha : nullptr;
  p.act = e ? e->act : 0;
  p.out_f32 = e ? e->out_f32 : 0;
  p.out2 = e ? reinterpret_cast<__nv_bfloat16*>(e->out2_bf16) : nullptr;
  p.out2_pre = e ? e->out2_pre_act : 0;
  p.dact_pre = e ? reinterpret_cast<const __nv_bfloat16*>(e->dact_pre) : nullptr;
  p.dact = e ? e->dact : 0;
  p.splits = (e && e->splits > 1) ? e->splits : 1;
  GROVE_CHECK_ARG(p.act >= 0 && p.act <= 2);
  GROVE_CHECK_ARG(!p.dact_pre || ((p.dact == 1 || p.dact == 2) && !p.out_f32 && ((uintptr_t)p.dact_pre & 15) == 0));
  GROVE_CHECK_ARG(p.out2_pre == 0 || (p.out2_pre == 1 && !p.out_f32) || (p.out2_pre == 2 && p.out_f32));
  // split-K writes raw fp32 partial planes [splits, M, N]; reduce them with grove_reduce_partials_f32
  GROVE_CHECK_ARG(p.splits == 1 || (p.out_f32 && !p.bias && !p.resid && !p.resid16 && !p.gate_alpha && !p.act && !p.out2));
  GROVE_CHECK_ARG(!p.resid16 || (!p.resid && !p.out_f32 && p.resid_mod == 0 && ((uintptr_t)p.resid16 & 15) == 0));
  GROVE_CHECK_ARG(!p.ln_stats || (p.ln_csum && p.ln_parts > 0 && p.ln_parts <= 16 && ((uintptr_t)p.ln_stats & 7) == 0 && ((uintptr_t)p.ln_csum & 15) == 0));
  GROVE_CHECK_ARG(!p.ln_stats_out || ((uintptr_t)p.ln_stats_out & 7) == 0);
  GROVE_CHECK_ARG(((uintptr_t)p.bias & 15) == 0 && ((uintptr_t)p.resid & 15) == 0 && ((uintptr_t)p.out2 & 15) == 0);
  (void)M; (void)N;
  return GROVE_OK;
}

extern "C" int grove_gemm_bf16(const void* A, const void* W, void* out, int M, int N, int K, const grove_gemm_epilogue* epi,
                               cudaStream_t stream) {
  GROVE_CHECK_ARG(A && W && out && M > 0 && N > 0 && K > 0);
  GROVE_CHECK_ARG(N % 128 == 0 && K % 8 == 0);
  GROVE_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)out & 15) == 0);
  GemmParams p{};
  p.M = M; p.N = N;
  p.num_k_blocks = (K + BK - 1) / BK;
  p.conv = 0;
  p.out = out;
  int rc = fill_epilogue(p, epi, M, N);
  if (rc) return rc;
  p.ln_inv_k = 1.0f / (float)K;
  uint64_t da[2] = {(uint64_t)K, (uint64_t)M};
  uint32_t ba[2] = {BK, BM};
  return dispatch_gemm(p, A, W, K, N, M, false, da, ba, 2, epi ? epi->max_ctas : 0, epi ? epi->force_ctas : 0, stream, nullptr,
                       epi ? epi->workspace : nullptr, epi && epi->workspace ? (size_t)epi->workspace_bytes : 0);
}

extern "C" int grove_conv_gemm_bf16(const void* X, const void* Wp, void* out, int V, int T, int G, int C, int N, int kt,
                                    const grove_gemm_epilogue* epi, cudaStream_t stream) {
  GROVE_CHECK_ARG(X && Wp && out && V > 0 && T > 0 && G > 0 && C > 0 && N > 0);
  GROVE_CHECK_ARG(kt == 1 || kt == 3);
  GROVE_CHECK_ARG(G <= 128 && 128 % G == 0 && (G * G) % 128 == 0);  // a 128-token M tile is whole grid rows of one frame
  GROVE_CHECK_ARG(C % 64 == 0 && N % 128 == 0);
  GROVE_CHECK_ARG(((uintptr_t)X & 15) == 0 && ((uintptr_t)Wp & 15) == 0 && ((uintptr_t)out & 15) == 0);
  const int ntaps = kt * 9;
  GemmParams p{};
  p.M = V * T * G * G; p.N = N;
  p.kc_blocks = C / BK;
  p.num_k_blocks = ntaps * p.kc_blocks;
  p.conv = 1; p.G = G; p.rows_per_tile = BM / G; p.tiles_per_frame = G * G / BM; p.T = T; p.kt = kt;
  p.out = out;
  int rc = fill_epilogue(p, epi, p.M, N);
  if (rc) return rc;
  uint64_t da[5] = {(uint64_t)C, (uint64_t)G, (uint64_t)G, (uint64_t)T, (uint64_t)V};
  uint32_t ba[5] = {BK, (uint32_t)G, (uint32_t)(BM / G), 1, 1};
  return dispatch_gemm(p, X, Wp, ntaps * C, N, p.M, true, da, ba, 5, epi ? epi->max_ctas : 0, epi ? epi->force_ctas : 0, stream, nullptr,
                       epi ? epi->workspace : nullptr, epi && epi->workspace ? (size_t)epi->workspace_bytes : 0);
}

/* Weight gradient of the implicit-GEMM convolutions: dWp[N, taps*C] (tap-major, fp32) = sum over tokens of
 * dY[token, n] * X[token + shift(tap), c].  dYt is dY transposed ([N, tokens] bf16); Xt3 holds X channel-major in three
 * w-shifted planes ([3, C, V,T,G,G] bf16, grove_transpose_shift3_to_bf16): K runs over tokens, the B operand is fetched with a
 * 5-D TMA box whose origin carries the tap's (h,t) shift (zero fill = 'same' padding) and whose plane carries the w shift. */
extern "C" int grove_conv_wgrad_bf16(const void* dYt, const void* Xt, float* dWp, int V, int T, int G, int C, int N, int kt, int splits,
                                     cudaStream_t stream) {
  GROVE_CHECK_ARG(dYt && Xt && dWp && V > 0 && T > 0 && G > 0 && C > 0 && N > 0);
  GROVE_CHECK_ARG(kt == 1 || kt == 3);
  GROVE_CHECK_ARG(64 % G == 0 && (G * G) % 64 == 0 && G >= 8);   // the h shift moves the box origin by G elements: 16-byte aligned
  GROVE_CHECK_ARG(C % 256 == 0);   // an n-block of 256 output columns is (tap, channel block)
  GROVE_CHECK_ARG(((uintptr_t)dYt & 15) == 0 && ((uintptr_t)Xt & 15) == 0 && ((uintptr_t)dWp & 15) == 0);
  const int ntaps = kt * 9;
  const long long tokens = (long long)V * T * G * G;
  GemmParams p{};
  p.M = N; p.N = ntaps * C;
  p.num_k_blocks = (int)(tokens / BK);
  p.conv = 2; p.G = G; p.T = T; p.kt = kt;
  p.kblocks_per_frame = G * G / BK; p.rows_per_kblock = BK / G;
  p.out = dWp; p.out_f32 = 1;
  p.splits = splits > 1 ? splits : 1;
  uint64_t da[2] = {(uint64_t)tokens, (uint64_t)N};
  uint32_t ba[2] = {BK, BM};
  p.wg_C = C;
  uint64_t db5[4] = {(uint64_t)G * G, (uint64_t)T, (uint64_t)V, (uint64_t)3 * C};
  return dispatch_gemm(p, dYt, Xt, (int)tokens, p.N, p.M, true, da, ba, 2, 0, 0, stream, db5);
}
