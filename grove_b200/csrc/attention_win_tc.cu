// Windowed attention (14x14 windows) with decomposed rel-pos bias on tcgen05 / TMEM / TMA
// (image_encoder.py:243-259 window path, 301-326, 329-384, 420-458).
//
// Persistent CTAs walk units (frame, window, head).  A unit's 196 window tokens arrive with ONE 4-D TMA box (64 ch x 14 x 14 x 1)
// per operand from the UNPARTITIONED token-major qkv tensor: rows land in window order, out-of-grid tokens are zero-filled by
// TMA and then overwritten with k = b_k, v = b_v by a dedicated "fixer" warp — exactly the reference's zero padding AFTER norm1
// (pad tokens take softmax mass, pad queries are never stored).  Per unit, two 128-row query tiles:
//   T = Q.[Rh;Rw]^T (UMMA 128x64xHD) -> rel_h / rel_w gathered per row into registers (x log2 e)
//   S = Q.K^T      (UMMA 128x208xHD) -> one-pass exact softmax in registers (two threads per row: keys [0,112) | [112,196))
//   P (bf16) is written back INTO the S columns of tensor memory and consumed from there (TS-mode UMMA): O = P.V (128xHDx208)
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-9 softmax/epilogue, 10 pad-token fixer.  TMEM: S0 | S1 | T/O = 208+208+80 columns.
#include <cuda.h>

#include "common.cuh"
#include "grove_b200.h"
#include "tmem_ldst.cuh"

namespace grove {

// In-kernel timeline probes (-DGROVE_WIN_PROBE; read with grove_win_probe_read, profiles/win_probe.py).  Compiled out by default.
#ifdef GROVE_WIN_PROBE
__device__ long long g_win_probe[4096];
#define WPROBE(slot) do { if (probe_on) g_win_probe[(slot)] = clock64(); } while (0)
#else
#define WPROBE(slot) do { } while (0)
#endif

constexpr int kWinThreads = 352;
constexpr int kWS = 14, kWQ = 196, kWK = 208;   // window side, tokens, keys padded to 13 UMMA k-steps

struct WinTmaps { CUtensorMap qkv, qkv_x, tab, tab_x; };

struct WinParams {
  const __nv_bfloat16* qkv_bias;   // [3*D] bf16
  __nv_bfloat16* out;              // [F,G,G,D]
  float* lse;                      // optional [F*G*G, heads]: log2-domain log-sum-exp of every real query row (training forward)
  int G, heads, nW, units;
};

template <int HD>
struct WinCfg {
  static constexpr bool kX = HD > 64;
  static constexpr int kQMain = 256 * 128, kQBytes = kQMain + (kX ? 256 * 32 : 0);
  static constexpr int kKMain = kWK * 128, kKBytes = ((kKMain + (kX ? kWK * 32 : 0)) + 1023) / 1024 * 1024;
  static constexpr int kTabMain = 64 * 128, kTabBytes = kTabMain + (kX ? 64 * 32 : 0);
  static constexpr int kStage = 128 * 64 * 4;
  static constexpr int kTxTile = kWQ * 128 + (kX ? kWQ * 32 : 0);          // bytes one window box delivers
  static constexpr int kSmem = kQBytes + 3 * kKBytes + kTabBytes + kStage + 4096 /*xch*/ + 1024 /*align*/ + 512 /*barriers*/;
};

// one query tile of the unit: bias + max + exp for this thread's keys, P -> tensor memory.  HS = 0: keys [0,112), 1: keys [112,196)
template <int HS>
__device__ __forceinline__ float win_softmax_tile(uint32_t tS, uint32_t tlane, const float (&relh)[8], const float (&relw)[14], float c_scale,
                                                  float* xch_max, int row, float& row_max) {
  constexpr int NK = HS == 0 ? 112 : 84;
  uint32_t s[112];
  if (HS == 0) {
    tmem_ld_x32(tS + 0 + tlane, reinterpret_cast<uint32_t(&)[32]>(s[0]));
    tmem_ld_x32(tS + 32 + tlane, reinterpret_cast<uint32_t(&)[32]>(s[32]));
    tmem_ld_x32(tS + 64 + tlane, reinterpret_cast<uint32_t(&)[32]>(s[64]));
    tmem_ld_x16(tS + 96 + tlane, reinterpret_cast<uint32_t(&)[16]>(s[96]));
  } else {
    tmem_ld_x16(tS + 112 + tlane, reinterpret_cast<uint32_t(&)[16]>(s[0]));
    tmem_ld_x32(tS + 128 + tlane, reinterpret_cast<uint32_t(&)[32]>(s[16]));
    tmem_ld_x32(tS + 160 + tlane, reinterpret_cast<uint32_t(&)[32]>(s[48]));
    tmem_ld_x4(tS + 192 + tlane, reinterpret_cast<uint32_t(&)[4]>(s[80]));
  }
  tmem_ld_wait();
  float mx = -INFINITY;
#ifndef GROVE_WIN_SCALAR
  // packed fp32 (FFMA2 / FADD2 / FMNMX3): the softmax warps are issue-bound here (computing exponentials on the FMA pipe instead of the
  // MUFU, i.e. MORE instructions, measured 123 -> 134-145 us per layer), so every per-element step works on a pair of keys.  A pair
  // (k, k+1) with k even never straddles a window row (14 is even): one rel_h value, one rel_w pair.
  const float2 c2 = make_float2(c_scale, c_scale);
  float mx1 = -INFINITY;
#pragma unroll
  for (int k = 0; k < NK; k += 2) {   // local key k -> window row k/14 (relh is already offset to this thread's first key row), column k%14
    float2 v = __ffma2_rn(make_float2(__uint_as_float(s[k]), __uint_as_float(s[k + 1])), c2, make_float2(relh[k / kWS], relh[k / kWS]));
    v = __fadd2_rn(v, make_float2(relw[k % kWS], relw[k % kWS + 1]));
    s[k] = __float_as_uint(v.x); s[k + 1] = __float_as_uint(v.y);
    if (k & 2) mx1 = fmaxf(mx1, fmaxf(v.x, v.y));
    else mx = fmaxf(mx, fmaxf(v.x, v.y));
  }
  mx = fmaxf(mx, mx1);
#else
#pragma unroll
  for (int k = 0; k < NK; ++k) {   // local key k -> window row k/14 (relh is already offset to this thread's first key row), column k%14
    const float v = fmaf(__uint_as_float(s[k]), c_scale, relh[k / kWS]) + relw[k % kWS];
    s[k] = __float_as_uint(v);
    mx = fmaxf(mx, v);
  }
#endif
  xch_max[HS * 128 + row] = mx;
  asm volatile("bar.sync 1, 256;" ::: "memory");   // also orders: both threads of the row have read S before P overwrites it
  mx = fmaxf(mx, xch_max[(HS ^ 1) * 128 + row]);
  row_max = mx;
  float lsum = 0.f;
#ifndef GROVE_WIN_SCALAR
  float2 lsum2 = make_float2(0.f, 0.f);
  const float2 nmx2 = make_float2(-mx, -mx);
#endif
  uint32_t pk[56];
#pragma unroll
#ifndef GROVE_WIN_POLY
#define GROVE_WIN_POLY 0x00                               /* bit i: pair i of every 8 computes 2^x on the FMA pipe (ex2_fma2) instead of the MUFU */
#endif
  for (int k = 0; k < NK; k += 2) {
    float p0, p1;
    if ((GROVE_WIN_POLY >> ((k >> 1) & 7)) & 1) {
      const float2 pp = ex2_fma2(make_float2(__uint_as_float(s[k]) - mx, __uint_as_float(s[k + 1]) - mx));
      p0 = pp.x; p1 = pp.y;
    } else {
#ifndef GROVE_WIN_SCALAR
      const float2 e = __fadd2_rn(make_float2(__uint_as_float(s[k]), __uint_as_float(s[k + 1])), nmx2);
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e.x));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e.y));
#else
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(__uint_as_float(s[k]) - mx));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(__uint_as_float(s[k + 1]) - mx));
#endif
    }
#ifndef GROVE_WIN_SCALAR
    lsum2 = __fadd2_rn(lsum2, make_float2(p0, p1));
#else
    lsum += p0 + p1;
#endif
    pk[k >> 1] = pack_bf16(p0, p1);
  }
#ifndef GROVE_WIN_SCALAR
  lsum = lsum2.x + lsum2.y;
#endif
  if (HS == 0) {     // keys 0..111 -> packed columns 0..55
    tmem_st_x32(tS + 0 + tlane, reinterpret_cast<const uint32_t(&)[32]>(pk[0]));
    tmem_st_x16(tS + 32 + tlane, reinterpret_cast<const uint32_t(&)[16]>(pk[32]));
    tmem_st_x8(tS + 48 + tlane, reinterpret_cast<const uint32_t(&)[8]>(pk[48]));
  } else {           // keys 112..195 -> columns 56..97, keys 196..207 (columns 98..103) are zero probability
#pragma unroll
    for (int i = 42; i < 48; ++i) pk[i] = 0u;
    tmem_st_x8(tS + 56 + tlane, reinterpret_cast<const uint32_t(&)[8]>(pk[0]));
    tmem_st_x32(tS + 64 + tlane, reinterpret_cast<const uint32_t(&)[32]>(pk[8]));
    tmem_st_x8(tS + 96 + tlane, reinterpret_cast<const uint32_t(&)[8]>(pk[40]));
  }
  tmem_st_wait();
  return lsum;
}

template <int HD>
__global__ void __launch_bounds__(kWinThreads, 1)
attn_window_tc_kernel(const __grid_constant__ WinTmaps tm, const WinParams p) {
  using Cfg = WinCfg<HD>;
  constexpr bool kX = Cfg::kX;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = s0, sK = sQ + Cfg::kQBytes, sV = sK + Cfg::kKBytes /* two buffers */, sTab = sV + 2 * Cfg::kKBytes;
  const uint32_t sStage = sTab + Cfg::kTabBytes, sXch = sStage + Cfg::kStage, bar0 = sXch + 4096;
  uint8_t* smem_al = smem_raw + (s0 - smem_u32(smem_raw));
  float* stage_f = reinterpret_cast<float*>(smem_al + (sStage - s0));
  float* xch_f = reinterpret_cast<float*>(smem_al + (sXch - s0));   // [0,256): max  [256,768): sums [tile][hs][row]
  enum { TAB_FULL = 0, QK_FULL, QK_EMPTY, K_FIX, V_FULL, V_EMPTY = V_FULL + 2, V_FIX = V_EMPTY + 2, T_FULL = V_FIX + 2, T_READ, S_FULL,
         P_FULL = S_FULL + 2, O_FULL = P_FULL + 2, O_READ, NUM_BARS };
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_slot = bar0 + 8u * NUM_BARS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.heads * HD;
  const int nWW = p.nW * p.nW;
#ifdef GROVE_WIN_PROBE
  const bool probe_on = blockIdx.x == 5 && lane == 0;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.qkv);
    mbar_init(bar(TAB_FULL), 1); mbar_init(bar(QK_FULL), 1); mbar_init(bar(QK_EMPTY), 1); mbar_init(bar(K_FIX), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(V_FULL + i), 1); mbar_init(bar(V_EMPTY + i), 1); mbar_init(bar(V_FIX + i), 1);
      mbar_init(bar(S_FULL + i), 1); mbar_init(bar(P_FULL + i), 8);
    }
    mbar_init(bar(T_FULL), 1); mbar_init(bar(T_READ), 8); mbar_init(bar(O_FULL), 1); mbar_init(bar(O_READ), 8);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // rows the TMA never writes must be finite: zero the whole Q / K / V area once
  for (uint32_t a = sQ + threadIdx.x * 16; a < sTab; a += kWinThreads * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                          // barrier init / TMEM allocation / the zero fill above overlapped the predecessor's last wave
  pdl_launch_dependents();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS0 = tmem_base, tTO = tmem_base + 416;

  auto decode = [&](int u, int& h, int& wy, int& wx, int& f) {
    h = u % p.heads;
    const int w = (u / p.heads) % nWW;
    f = u / (p.heads * nWW);
    wy = w / p.nW; wx = w % p.nW;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    {
      auto load_box = [&](uint32_t dst_main, uint32_t dst_tail, uint32_t b, int col, int wx, int wy, int f) {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst_main),
                     "l"(reinterpret_cast<uint64_t>(&tm.qkv)), "r"(b), "r"(col), "r"(wx * kWS), "r"(wy * kWS), "r"(f) : "memory");
        if (kX)
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst_tail),
                       "l"(reinterpret_cast<uint64_t>(&tm.qkv_x)), "r"(b), "r"(col + 64), "r"(wx * kWS), "r"(wy * kWS), "r"(f) : "memory");
      };
      if (elect_one()) {
        mbar_expect_tx(bar(TAB_FULL), Cfg::kTabBytes);
        tma_load_2d(sTab, &tm.tab, bar(TAB_FULL), 0, 0);
        if (kX) tma_load_2d(sTab + Cfg::kTabMain, &tm.tab_x, bar(TAB_FULL), 64, 0);
      }
      __syncwarp();
      uint32_t cnt = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++cnt) {
        int h, wy, wx, f;
        decode(u, h, wy, wx, f);
        const uint32_t vb = cnt & 1u;
        mbar_wait2(bar(QK_EMPTY), (cnt & 1u) ^ 1u, bar(V_EMPTY + vb), ((cnt >> 1) & 1u) ^ 1u);
        if (elect_one()) {     // uniform-datapath instructions: see elect_one() in common.cuh
          mbar_expect_tx(bar(QK_FULL), 2 * Cfg::kTxTile);
          load_box(sQ, sQ + Cfg::kQMain, bar(QK_FULL), h * HD, wx, wy, f);
          load_box(sK, sK + Cfg::kKMain, bar(QK_FULL), D + h * HD, wx, wy, f);
          mbar_expect_tx(bar(V_FULL + vb), Cfg::kTxTile);
          load_box(sV + vb * Cfg::kKBytes, sV + vb * Cfg::kKBytes + Cfg::kKMain, bar(V_FULL + vb), 2 * D + h * HD, wx, wy, f);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64), idesc_s = umma_idesc_bf16(128, kWK);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16), idesc_ox = umma_idesc_bf16(128, 16) | (1u << 16);
    auto qk_mma = [&](uint32_t d, uint32_t a_main, uint32_t a_tail, uint32_t b_main, uint32_t b_tail, uint32_t idesc) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16(d, umma_desc_sw128(a_main + k * 32), umma_desc_sw128(b_main + k * 32), idesc, k != 0);
      if (kX) tc_mma_f16(d, umma_desc_sw32(a_tail), umma_desc_sw32(b_tail), idesc, 1);
    };
    mbar_wait(bar(TAB_FULL), 0);
    uint32_t cnt = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++cnt) {
      const uint32_t vb = cnt & 1u;
      const uint32_t vbase = sV + vb * Cfg::kKBytes;
      WPROBE(100 + 16 * (cnt & 7) + 0);
      mbar_wait2(bar(QK_FULL), cnt & 1u, bar(K_FIX), cnt & 1u);
      WPROBE(100 + 16 * (cnt & 7) + 1);
      if (cnt > 0) mbar_wait(bar(O_READ), 1u);            // previous unit's second epilogue has drained the T/O columns
      WPROBE(100 + 16 * (cnt & 7) + 2);
      tc_fence_after();
      if (elect_one()) {
        qk_mma(tTO, sQ, sQ + Cfg::kQMain, sTab, sTab + Cfg::kTabMain, idesc_t);
        tc_commit(bar(T_FULL));
        qk_mma(tS0, sQ, sQ + Cfg::kQMain, sK, sK + Cfg::kKMain, idesc_s);
        tc_commit(bar(S_FULL + 0));
      }
      __syncwarp();
      WPROBE(100 + 16 * (cnt & 7) + 3);
      mbar_wait(bar(T_READ), 0u);                           // tile 0's T is in registers
      WPROBE(100 + 16 * (cnt & 7) + 4);
      tc_fence_after();
      if (elect_one()) {
        qk_mma(tTO, sQ + 16384, sQ + Cfg::kQMain + 4096, sTab, sTab + Cfg::kTabMain, idesc_t);
        tc_commit(bar(T_FULL));
        qk_mma(tS0 + kWK, sQ + 16384, sQ + Cfg::kQMain + 4096, sK, sK + Cfg::kKMain, idesc_s);
        tc_commit(bar(S_FULL + 1));
        tc_commit(bar(QK_EMPTY));                            // Q and K may be overwritten by the next unit's loads
      }
      __syncwarp();
      WPROBE(100 + 16 * (cnt & 7) + 5);
      mbar_wait(bar(T_READ), 1u);
      WPROBE(100 + 16 * (cnt & 7) + 6);
      mbar_wait2(bar(V_FULL + vb), (cnt >> 1) & 1u, bar(V_FIX + vb), (cnt >> 1) & 1u);
      WPROBE(100 + 16 * (cnt & 7) + 7);
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        mbar_wait(bar(P_FULL + t), cnt & 1u);
        WPROBE(100 + 16 * (cnt & 7) + 8 + 3 * t);
        if (t == 1) mbar_wait(bar(O_READ), 0u);             // tile 0's O has been read
        WPROBE(100 + 16 * (cnt & 7) + 9 + 3 * t);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kWK / 16; ++kk) {
            const uint32_t ta = tS0 + t * kWK + kk * 8;
            tc_mma_f16_ts(tTO, ta, umma_desc_sw128(vbase + kk * 2048), idesc_o, kk != 0);
            if (kX) tc_mma_f16_ts(tTO + 64, ta, umma_desc_sw32(vbase + Cfg::kKMain + kk * 512), idesc_ox, kk != 0);
          }
          tc_commit(bar(O_FULL));
          if (t == 1) tc_commit(bar(V_EMPTY + vb));
        }
        __syncwarp();
        WPROBE(100 + 16 * (cnt & 7) + 10 + 3 * t);
      }
    }
  } else if (warp == 10) {
    // ===================== pad-token fixer: out-of-grid window tokens get k = b_k, v = b_v =====================
    uint32_t cnt = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++cnt) {
      int h, wy, wx, f;
      decode(u, h, wy, wx, f);
      const bool edge = (wy + 1) * kWS > p.G || (wx + 1) * kWS > p.G;
      const uint32_t vb = cnt & 1u;
#pragma unroll 1
      for (int which = 1; which <= 2; ++which) {             // 1: K, 2: V
        const uint32_t base = which == 1 ? sK : sV + vb * Cfg::kKBytes;
        if (which == 1) mbar_wait(bar(QK_FULL), cnt & 1u);
        else mbar_wait(bar(V_FULL + vb), (cnt >> 1) & 1u);
        if (edge) {
          const __nv_bfloat16* bsrc = p.qkv_bias + which * D + h * HD;
          for (int r = lane; r < kWQ; r += 32) {
            if (wy * kWS + r / kWS >= p.G || wx * kWS + r % kWS >= p.G) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(bsrc + c * 8));
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + r * 128 + ((c ^ (r & 7)) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
              }
              if (kX) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const uint4 v = __ldg(reinterpret_cast<const uint4*>(bsrc + 64 + c * 8));
                  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + Cfg::kKMain + r * 32 + ((c ^ ((r >> 2) & 1)) << 4)), "r"(v.x), "r"(v.y),
                               "r"(v.z), "r"(v.w) : "memory");
                }
              }
            }
          }
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(which == 1 ? K_FIX : V_FIX + vb));
      }
    }
  } else {
    // ===================== softmax / epilogue warps: two threads per query row =====================
    const int quad = warp & 3;
    const int hs = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    constexpr float kL2e = 1.4426950408889634f;
    const float c_scale = (HD == 64 ? 0.125f : 0.11180339887498949f) * kL2e;
    auto softmax_sync = []() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    auto stage_at = [&](int e) { return stage_f[row * 64 + ((((e >> 2) ^ (row & 7)) << 2) | (e & 3))]; };
    uint32_t cnt = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++cnt) {
      int h, wy, wx, f;
      decode(u, h, wy, wx, f);
      float lsum[2], rmax[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int q = min(t * 128 + row, kWQ - 1);          // rows >= 196 compute on a clamped position and are never stored
        const int qh = q / kWS, qw = q % kWS;
        // ---- rel-pos products of this tile: T_h = columns [0,27), T_w = columns [32,59)
        WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 8 * t + 0);
        mbar_wait(bar(T_FULL), (uint32_t)t);
        WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 8 * t + 1);
        tc_fence_after();
        {
          uint32_t r[32];
          tmem_ld_x32(tTO + hs * 32 + tlane, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stage_f + row * 64 + (((hs * 8 + (j >> 2)) ^ (row & 7)) << 2)) =
                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(T_READ));
        softmax_sync();
        float relh[8], relw[14];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kh = hs * 8 + i;                        // thread 0 of the row owns key rows 0..7, thread 1 rows 8..13
          relh[i] = kh < kWS ? stage_at(qh + (kWS - 1) - kh) * kL2e : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) relw[i] = stage_at(32 + qw + (kWS - 1) - i) * kL2e;
        softmax_sync();
        // ---- softmax of this tile
        WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 8 * t + 2);
        mbar_wait(bar(S_FULL + t), cnt & 1u);
        WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 8 * t + 3);
        tc_fence_after();
        const uint32_t tS = tS0 + t * kWK;
        lsum[t] = hs == 0 ? win_softmax_tile<0>(tS, tlane, relh, relw, c_scale, xch_f, row, rmax[t])
                          : win_softmax_tile<1>(tS, tlane, relh, relw, c_scale, xch_f, row, rmax[t]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(P_FULL + t));
        WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 8 * t + 4);
        xch_f[256 + (t * 2 + hs) * 128 + row] = lsum[t];
      }
      softmax_sync();
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 8 * t + 5);
        mbar_wait(bar(O_FULL), (uint32_t)t);
        WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 8 * t + 6);
        tc_fence_after();
        uint32_t r[32], rx[8];
        tmem_ld_x32(tTO + hs * 32 + tlane, r);
        if (kX) tmem_ld_x8(tTO + 64 + hs * 8 + tlane, rx);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(O_READ));
        const float ltot = lsum[t] + xch_f[256 + (t * 2 + (hs ^ 1)) * 128 + row];
        const float inv = 1.f / ltot;
        const int q = t * 128 + row;
        const int gy = wy * kWS + q / kWS, gx = wx * kWS + q % kWS;
        if (p.lse != nullptr && hs == 0 && q < kWQ && gy < p.G && gx < p.G)
          p.lse[((size_t)(f * p.G + gy) * p.G + gx) * p.heads + h] = rmax[t] + log2f(ltot);
        if (q < kWQ && gy < p.G && gx < p.G) {
          __nv_bfloat16* orow = p.out + ((size_t)(f * p.G + gy) * p.G + gx) * D + h * HD;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(orow + hs * 32 + j) =
                make_uint4(pack_bf16(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv), pack_bf16(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv),
                           pack_bf16(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv), pack_bf16(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv));
          if (kX)
            *reinterpret_cast<uint4*>(orow + 64 + hs * 8) =
                make_uint4(pack_bf16(__uint_as_float(rx[0]) * inv, __uint_as_float(rx[1]) * inv), pack_bf16(__uint_as_float(rx[2]) * inv, __uint_as_float(rx[3]) * inv),
                           pack_bf16(__uint_as_float(rx[4]) * inv, __uint_as_float(rx[5]) * inv), pack_bf16(__uint_as_float(rx[6]) * inv, __uint_as_float(rx[7]) * inv));
        }
      }
      WPROBE(1000 + (warp - 2) * 256 + 32 * (cnt & 7) + 16 + 7);
      softmax_sync();   // the sum exchange slots are rewritten by the next unit
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_tmap_bf16_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows);
int make_tmap_bf16_nd(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box);

}  // namespace grove
using namespace grove;

#ifdef GROVE_WIN_PROBE
extern "C" int grove_win_probe_read(long long* host, int n) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host, g_win_probe, sizeof(long long) * n) == cudaSuccess ? 0 : 1;
}
#endif

template <int HD>
static int launch_window_tc(const void* qkv, const void* qkv_bias, const void* tab, void* out, float* lse, int F, int G, int heads, cudaStream_t stream) {
  using Cfg = WinCfg<HD>;
  const int D = heads * HD;
  WinTmaps tm;
  int rc;
  uint64_t dims[4] = {(uint64_t)3 * D, (uint64_t)G, (uint64_t)G, (uint64_t)F};
  uint32_t box[4] = {64, kWS, kWS, 1};
  if ((rc = make_tmap_bf16_nd(&tm.qkv, qkv, 4, dims, box))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.tab, tab, HD, 64, 64, 64))) return rc;
  if (HD > 64) {
    uint32_t boxx[4] = {16, kWS, kWS, 1};
    if ((rc = make_tmap_bf16_nd(&tm.qkv_x, qkv, 4, dims, boxx))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.tab_x, tab, HD, 64, 16, 64))) return rc;
  } else {
    tm.qkv_x = tm.qkv; tm.tab_x = tm.tab;
  }
  WinParams p;
  p.qkv_bias = reinterpret_cast<const __nv_bfloat16*>(qkv_bias);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.G = G; p.heads = heads; p.nW = (G + kWS - 1) / kWS;
  p.units = F * p.nW * p.nW * heads;
  cudaError_t e = cudaFuncSetAttribute(attn_window_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
  if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute(%d): %s", Cfg::kSmem, cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.units < sms ? p.units : sms;
  grove_launch_pdl(attn_window_tc_kernel<HD>, dim3(grid), dim3(kWinThreads), Cfg::kSmem, stream, tm, p);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_attn_window_relpos_tc_fwd_lse(const void* qkv, const void* qkv_bias_bf16, const void* rel_table, void* out, float* lse, int F, int G,
                                                   int heads, int hd, int ws, cudaStream_t stream);

extern "C" int grove_attn_window_relpos_tc_fwd(const void* qkv, const void* qkv_bias_bf16, const void* rel_table, void* out, int F, int G, int heads,
                                               int hd, int ws, cudaStream_t stream) {
  return grove_attn_window_relpos_tc_fwd_lse(qkv, qkv_bias_bf16, rel_table, out, nullptr, F, G, heads, hd, ws, stream);
}

extern "C" int grove_attn_window_relpos_tc_fwd_lse(const void* qkv, const void* qkv_bias_bf16, const void* rel_table, void* out, float* lse, int F, int G,
                                                   int heads, int hd, int ws, cudaStream_t stream) {
  GROVE_CHECK_ARG(qkv && qkv_bias_bf16 && rel_table && out && F > 0 && G > 0 && heads > 0);
  if ((hd != 64 && hd != 80) || ws != kWS) {
    grove_set_error("grove_attn_window_relpos_tc_fwd: head dim 64 / 80 and window 14 are built (got hd=%d ws=%d)", hd, ws);
    return GROVE_ERR_UNSUPPORTED;
  }
  GROVE_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)rel_table & 15) == 0 && ((uintptr_t)qkv_bias_bf16 & 15) == 0 && ((uintptr_t)out & 15) == 0);
  return hd == 64 ? launch_window_tc<64>(qkv, qkv_bias_bf16, rel_table, out, lse, F, G, heads, stream)
                  : launch_window_tc<80>(qkv, qkv_bias_bf16, rel_table, out, lse, F, G, heads, stream);
}
