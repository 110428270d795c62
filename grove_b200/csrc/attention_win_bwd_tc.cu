// Backward of the WINDOWED attention (14x14 windows, decomposed rel-pos bias) on tcgen05 / TMEM / TMA: image_encoder.py:243-259 window
// path, 301-326, 329-384, 420-458 under autograd; the blocks are frozen (train.py:254-255), so only d(qkv) is produced.  Replaces the
// four warp-level launches of attention_bwd.cu (relpos_kernel<0>, attn_bwd_q_kernel, attn_bwd_kv_kernel, relpos_kernel<1>) for head dims 64 and 80
// (80 = 64 + a 16-wide tail: SWIZZLE_32B operand tiles, one extra k-step / one extra N = 16 product per UMMA group).
//
//   S[q,k] = scale q.k + q.Rh[qh-kh+13] + q.Rw[qw-kw+13],  P = exp(S - lse_q),  dP = dO V^T,  dS = P (dP - D_q)
//   dQ = scale dS K + sum_c A[q,c] R[idx(q,c)]  (A_h[q,kh] = sum_kw dS, A_w[q,kw] = sum_kh dS),  dK = scale dS^T Q,  dV = P^T dO
//
// A window is self-contained (196 queries x 196 keys), so ONE persistent CTA does a whole unit (frame, window, head): its Q, K, V, dO
// tiles arrive with one 4-D TMA box each from the UNPARTITIONED token-major tensors (rows in window order, out-of-grid tokens zero-filled;
// a fixer warp rewrites out-of-grid K / V rows to b_k / b_v — the zero padding AFTER norm1 of the reference: pad keys take softmax mass,
// pad queries carry no cotangent).  Four passes per unit, every product a single-shot UMMA:
//   query passes (2 x 128 queries, TMEM lane = query):
//     T  = Q_t [Rh;Rw]^T (128x64x64)   -> rel_h / rel_w of the row gathered into registers; also kept, with lse and D, in a shared table
//     S  = Q_t K^T, dP = dO_t V^T (128x208x64) -> dS in registers: A_h / A_w row sums, scale*dS (bf16) back into tensor memory
//     dQ = (scale dS) K  (TS-mode, 128x64x208)  +  dT Tab (SS-mode, 128x64x64; dT = A scattered to the table rows, bf16 tile in smem)
//   key passes (2 x 128 keys, TMEM lane = key; the TRANSPOSED products, so P^T / dS^T are directly the A operands):
//     S^T = K_t Q^T, dP^T = V_t dO^T (128x208x64) -> P^T, dS^T (bf16) back into tensor memory (bias / lse / D from the shared table)
//     dV = P^T dO, dK = dS^T Q  (TS-mode, 128x64x208)
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-9 elementwise (two threads per row: columns [0,112) | [112,208)), 10 pad-token fixer.
// TMEM: S | dP | T/dQ/dV = 208 + 208 + HD columns; packed bf16 results overwrite the fp32 columns their own thread has already read
// (thread 0 ascending into [0,56), thread 1 descending into [160,208)), which leaves [64, 64 + HD) of the S region free for dK.
#include <cuda.h>

#include "common.cuh"
#include "grove_b200.h"
#include "tmem_ldst.cuh"

namespace grove {

constexpr int kWbThreads = 352;
constexpr int kBS = 14, kBQ = 196, kBK = 208;   // window side, tokens, tokens padded to 13 UMMA k-steps
// shared bias table of a unit, TRANSPOSED: row c (0..13: rel_h of key row c, x log2 e, - lse; 14..27: rel_w of key column c - 14, x log2 e;
// 28: -D), one float per query, so a key-pass thread (fixed kh, kw) reads four consecutive queries with one 16-byte load per row
constexpr int kRelRows = 29, kRelStride = 212;

struct WinBwdTmaps { CUtensorMap qkv, dO, rh, rw, qkv_x, dO_x, rh_x, rw_x; };   // _x: the 16-wide tail of an 80-wide head (SWIZZLE_32B)

struct WinBwdParams {
  const __nv_bfloat16* qkv_bias;   // [3*D] bf16
  const float* lse;                // [M, heads] log2-domain log-sum-exp of the forward kernel
  const float* dsum;               // [M, heads] D = rowsum(dO * O)
  __nv_bfloat16* dqkv;             // [M, 3*D]
  int G, heads, nW, units;
};

// One operand = 208 rows (196 real, 12 zero) x 64 bf16 with SWIZZLE_128B [+ 208 x 16 bf16 with SWIZZLE_32B for an 80-wide head].  As the A
// operand of the second 128-row tile it is read 48 rows past its end: those rows only produce TMEM lanes >= 208, which nothing ever reads back.
template <int HD>
struct WinBwdCfg {
  static constexpr bool kX = HD > 64;
  static constexpr int kMain = kBK * 128, kOp = kMain + (kX ? 7 * 1024 : 0);
  static constexpr int kTabMain = 64 * 128, kTab = kTabMain + (kX ? 64 * 32 : 0);
  static constexpr int kDT = 128 * 128, kStage = 128 * 64 * 4;
  static constexpr int kRel = (kRelRows * kRelStride * 4 + 1023) / 1024 * 1024;
  static constexpr int kXch = 2 * 7 * 128 * 4;
  static constexpr int kTx = 4 * (kBQ * 128 + (kX ? kBQ * 32 : 0));     // bytes the window boxes of Q, K, V, dO deliver
  static constexpr int kSmem = 4 * kOp + kTab + kDT + kStage + kRel + kXch + 1024 /*align*/ + 512 /*barriers*/;
};

template <int W>
__device__ __forceinline__ void wb_ld(uint32_t taddr, uint32_t (&r)[W]) {
  if constexpr (W == 32) tmem_ld_x32(taddr, r);
  else tmem_ld_x16(taddr, r);
}
template <int W>
__device__ __forceinline__ void wb_st(uint32_t taddr, const uint32_t (&r)[W]) {
  if constexpr (W == 16) tmem_st_x16(taddr, r);
  else tmem_st_x8(taddr, r);
}
__device__ __forceinline__ float wb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// query pass, one chunk of W key columns starting at key BASE; relh[i] belongs to key row KH0 + i.  scale*dS -> columns PCOL.. of the dP region.
// Packed fp32 throughout (the elementwise warps are issue-bound): a pair of keys (k, k+1), k even, never straddles a window row, so it takes
// one rel_h value and one rel_w pair; the A_h / A_w sums are float2 accumulators (A_h folded at the end).
template <int BASE, int W, int PCOL, int KH0>
__device__ __forceinline__ void wb_q_chunk(uint32_t tS, uint32_t tDP, uint32_t tlane, const float (&relh)[8], const float (&relw)[14], float c_l2, float scale,
                                           float my_d, float2 (&ah2)[8], float2 (&aw2)[7]) {
  uint32_t rs[W], rd[W], pd[W / 2];
  wb_ld<W>(tS + BASE + tlane, rs);
  wb_ld<W>(tDP + BASE + tlane, rd);
  tmem_ld_wait();
  const float2 c2 = make_float2(c_l2, c_l2), nd2 = make_float2(-my_d, -my_d), sc2 = make_float2(scale, scale);
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    const int k = BASE + j;
    if (k < kBQ) {
      const int kh = k / kBS - KH0, kw = k % kBS;
      float2 v = __ffma2_rn(make_float2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), c2, make_float2(relh[kh], relh[kh]));
      v = __fadd2_rn(v, make_float2(relw[kw], relw[kw + 1]));
      const float2 p = make_float2(wb_ex2(v.x), wb_ex2(v.y));
      const float2 ds = __fmul2_rn(p, __fadd2_rn(make_float2(__uint_as_float(rd[j]), __uint_as_float(rd[j + 1])), nd2));
      ah2[kh] = __fadd2_rn(ah2[kh], ds);
      aw2[kw >> 1] = __fadd2_rn(aw2[kw >> 1], ds);
      const float2 o = __fmul2_rn(ds, sc2);
      pd[j >> 1] = pack_bf16(o.x, o.y);
    } else {
      pd[j >> 1] = 0u;                                   // keys 196..207 do not exist
    }
  }
  wb_st<W / 2>(tDP + PCOL + tlane, pd);
}

// key pass, one chunk of W query columns starting at query BASE.  P^T -> the S region, dS^T -> the dP region, columns PCOL..
// ph / pw / pd: this thread's rel_h row, rel_w row and the -D row of the transposed bias table (four queries per 16-byte load)
template <int BASE, int W, int PCOL>
__device__ __forceinline__ void wb_k_chunk(uint32_t tS, uint32_t tDP, uint32_t tlane, const float* ph, const float* pw, const float* pdn, float c_l2) {
  uint32_t rs[W], rd[W], pp[W / 2], pd[W / 2];
  wb_ld<W>(tS + BASE + tlane, rs);
  wb_ld<W>(tDP + BASE + tlane, rd);
  tmem_ld_wait();
  const float2 c2 = make_float2(c_l2, c_l2);
#pragma unroll
  for (int j = 0; j < W; j += 4) {
    const int q = BASE + j;
    if (q < kBQ) {                                       // 196 is a multiple of four: a group never straddles the end
      const float4 h4 = *reinterpret_cast<const float4*>(ph + q), w4 = *reinterpret_cast<const float4*>(pw + q);
      const float4 d4 = *reinterpret_cast<const float4*>(pdn + q);
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const float2 b = __fadd2_rn(e == 0 ? make_float2(h4.x, h4.y) : make_float2(h4.z, h4.w), e == 0 ? make_float2(w4.x, w4.y) : make_float2(w4.z, w4.w));
        const float2 v = __ffma2_rn(make_float2(__uint_as_float(rs[j + e]), __uint_as_float(rs[j + e + 1])), c2, b);
        const float2 p = make_float2(wb_ex2(v.x), wb_ex2(v.y));
        const float2 ds = __fmul2_rn(p, __fadd2_rn(make_float2(__uint_as_float(rd[j + e]), __uint_as_float(rd[j + e + 1])),
                                                   e == 0 ? make_float2(d4.x, d4.y) : make_float2(d4.z, d4.w)));
        pp[(j + e) >> 1] = pack_bf16(p.x, p.y);
        pd[(j + e) >> 1] = pack_bf16(ds.x, ds.y);
      }
    } else {
      pp[j >> 1] = pp[(j >> 1) + 1] = 0u;
      pd[j >> 1] = pd[(j >> 1) + 1] = 0u;
    }
  }
  wb_st<W / 2>(tS + PCOL + tlane, pp);
  wb_st<W / 2>(tDP + PCOL + tlane, pd);
}

__device__ __forceinline__ void wb_store32(__nv_bfloat16* o, const uint32_t (&r)[32], float sc) {
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    *reinterpret_cast<uint4*>(o + j) =
        make_uint4(pack_bf16(__uint_as_float(r[j]) * sc, __uint_as_float(r[j + 1]) * sc), pack_bf16(__uint_as_float(r[j + 2]) * sc, __uint_as_float(r[j + 3]) * sc),
                   pack_bf16(__uint_as_float(r[j + 4]) * sc, __uint_as_float(r[j + 5]) * sc), pack_bf16(__uint_as_float(r[j + 6]) * sc, __uint_as_float(r[j + 7]) * sc));
}

__device__ __forceinline__ void wb_store8(__nv_bfloat16* o, const uint32_t (&r)[8], float sc) {
  *reinterpret_cast<uint4*>(o) =
      make_uint4(pack_bf16(__uint_as_float(r[0]) * sc, __uint_as_float(r[1]) * sc), pack_bf16(__uint_as_float(r[2]) * sc, __uint_as_float(r[3]) * sc),
                 pack_bf16(__uint_as_float(r[4]) * sc, __uint_as_float(r[5]) * sc), pack_bf16(__uint_as_float(r[6]) * sc, __uint_as_float(r[7]) * sc));
}

template <int HD>
__global__ void __launch_bounds__(kWbThreads, 1)
attn_window_bwd_tc_kernel(const __grid_constant__ WinBwdTmaps tm, const WinBwdParams p) {
  using C = WinBwdCfg<HD>;
  constexpr bool kX = C::kX;
  constexpr float kScale = HD == 64 ? 0.125f : 0.11180339887498949f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = s0, sK = sQ + C::kOp, sV = sK + C::kOp, sDO = sV + C::kOp, sTab = sDO + C::kOp, sDT = sTab + C::kTab;
  const uint32_t sStage = sDT + C::kDT, sRel = sStage + C::kStage, sXch = sRel + C::kRel, bar0 = sXch + C::kXch;
  uint8_t* smem_al = smem_raw + (s0 - smem_u32(smem_raw));
  float* stage_f = reinterpret_cast<float*>(smem_al + (sStage - s0));
  float* rel_s = reinterpret_cast<float*>(smem_al + (sRel - s0));
  float* xch_f = reinterpret_cast<float*>(smem_al + (sXch - s0));     // [hs][kw][row]
  uint8_t* dt_b = smem_al + (sDT - s0);
  enum { TAB_FULL = 0, LOAD_FULL, OPS_EMPTY, FIX_DONE, T_FULL, S_FULL, P_FULL, O_FULL, ACC_READ, NUM_BARS };
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_slot = bar0 + 8u * NUM_BARS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.heads * HD;
  const int nWW = p.nW * p.nW;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.qkv);
    tma_prefetch_desc(&tm.dO);
    mbar_init(bar(TAB_FULL), 1); mbar_init(bar(LOAD_FULL), 1); mbar_init(bar(OPS_EMPTY), 1); mbar_init(bar(FIX_DONE), 1);
    mbar_init(bar(T_FULL), 1); mbar_init(bar(S_FULL), 1); mbar_init(bar(P_FULL), 8); mbar_init(bar(O_FULL), 1); mbar_init(bar(ACC_READ), 8);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // rows the TMA never writes (196..255 of every operand, the table's unused rows) must be finite: zero everything once
  for (uint32_t a = sQ + threadIdx.x * 16; a < sStage; a += kWbThreads * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < kRelRows * kRelStride; i += kWbThreads) rel_s[i] = 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tDP = tmem_base + kBK, tTO = tmem_base + 2 * kBK, tDK = tmem_base + 64;

  auto decode = [&](int u, int& h, int& wy, int& wx, int& f) {
    h = u % p.heads;
    const int w = (u / p.heads) % nWW;
    f = u / (p.heads * nWW);
    wy = w / p.nW; wx = w % p.nW;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(bar(TAB_FULL), 2 * 32 * 128 + (kX ? 2 * 32 * 32 : 0));
      tma_load_2d(sTab, &tm.rh, bar(TAB_FULL), 0, 0);              // rows 0..26 = Rh, 27..31 zero-filled (out of bounds)
      tma_load_2d(sTab + 32 * 128, &tm.rw, bar(TAB_FULL), 0, 0);   // rows 32..58 = Rw
      if (kX) {
        tma_load_2d(sTab + C::kTabMain, &tm.rh_x, bar(TAB_FULL), 64, 0);
        tma_load_2d(sTab + C::kTabMain + 32 * 32, &tm.rw_x, bar(TAB_FULL), 64, 0);
      }
    }
    __syncwarp();
    uint32_t cnt = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++cnt) {
      int h, wy, wx, f;
      decode(u, h, wy, wx, f);
      mbar_wait(bar(OPS_EMPTY), (cnt & 1u) ^ 1u);                  // every product of the previous unit has retired
      if (elect_one()) {
        mbar_expect_tx(bar(LOAD_FULL), C::kTx);
        tma_load_4d_cta(sQ, &tm.qkv, bar(LOAD_FULL), h * HD, wx * kBS, wy * kBS, f);
        tma_load_4d_cta(sDO, &tm.dO, bar(LOAD_FULL), h * HD, wx * kBS, wy * kBS, f);
        tma_load_4d_cta(sK, &tm.qkv, bar(LOAD_FULL), D + h * HD, wx * kBS, wy * kBS, f);
        tma_load_4d_cta(sV, &tm.qkv, bar(LOAD_FULL), 2 * D + h * HD, wx * kBS, wy * kBS, f);
        if (kX) {
          tma_load_4d_cta(sQ + C::kMain, &tm.qkv_x, bar(LOAD_FULL), h * HD + 64, wx * kBS, wy * kBS, f);
          tma_load_4d_cta(sDO + C::kMain, &tm.dO_x, bar(LOAD_FULL), h * HD + 64, wx * kBS, wy * kBS, f);
          tma_load_4d_cta(sK + C::kMain, &tm.qkv_x, bar(LOAD_FULL), D + h * HD + 64, wx * kBS, wy * kBS, f);
          tma_load_4d_cta(sV + C::kMain, &tm.qkv_x, bar(LOAD_FULL), 2 * D + h * HD + 64, wx * kBS, wy * kBS, f);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64), idesc_s = umma_idesc_bf16(128, kBK);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);        // B read MN-major ([k][d] tile, d contiguous)
    constexpr uint32_t idesc_ox = umma_idesc_bf16(128, 16) | (1u << 16);
    // D = A_tile B^T over the head dim, both K-major; a / b are operand bases, `t` picks the 128-row tile of A
    auto ss = [&](uint32_t d, uint32_t a, int t, uint32_t b, uint32_t b_tail, uint32_t idesc) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16(d, umma_desc_sw128(a + t * 16384 + k * 32), umma_desc_sw128(b + k * 32), idesc, k != 0);
      if (kX) tc_mma_f16(d, umma_desc_sw32(a + C::kMain + t * 4096), umma_desc_sw32(b_tail), idesc, 1);
    };
    auto ts = [&](uint32_t d, uint32_t a_region, uint32_t b) {                 // D = A[tmem, 128 x 208 packed] B[208 x HD]
#pragma unroll
      for (int kk = 0; kk < kBK / 16; ++kk) {
        const uint32_t ta = a_region + (kk < 7 ? kk * 8 : 160 + (kk - 7) * 8);
        tc_mma_f16_ts(d, ta, umma_desc_sw128(b + kk * 2048), idesc_o, kk != 0);
        if (kX) tc_mma_f16_ts(d + 64, ta, umma_desc_sw32(b + C::kMain + kk * 512), idesc_ox, kk != 0);
      }
    };
    mbar_wait(bar(TAB_FULL), 0);
    uint32_t cnt = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++cnt) {
      mbar_wait2(bar(LOAD_FULL), cnt & 1u, bar(FIX_DONE), cnt & 1u);
#pragma unroll 1
      for (int ps = 0; ps < 4; ++ps) {
        const int t = ps & 1;
        if (cnt > 0 || ps > 0) mbar_wait(bar(ACC_READ), (uint32_t)(ps + 3) & 1u);     // the previous pass' accumulators have been read
        tc_fence_after();
        if (elect_one()) {
          if (ps < 2) {
            ss(tTO, sQ, t, sTab, sTab + C::kTabMain, idesc_t);
            tc_commit(bar(T_FULL));
            ss(tS, sQ, t, sK, sK + C::kMain, idesc_s);
            ss(tDP, sDO, t, sV, sV + C::kMain, idesc_s);
          } else {
            ss(tS, sK, t, sQ, sQ + C::kMain, idesc_s);
            ss(tDP, sV, t, sDO, sDO + C::kMain, idesc_s);
          }
          tc_commit(bar(S_FULL));
        }
        __syncwarp();
        mbar_wait(bar(P_FULL), (uint32_t)ps & 1u);
        tc_fence_after();
        if (elect_one()) {
          if (ps < 2) {
            ts(tTO, tDP, sK);                                                  // dQ = (scale dS) K
#pragma unroll
            for (int k = 0; k < 4; ++k) {                                       // + dT Tab
              tc_mma_f16(tTO, umma_desc_sw128(sDT + k * 32), umma_desc_sw128(sTab + k * 2048), idesc_o, 1);
              if (kX) tc_mma_f16(tTO + 64, umma_desc_sw128(sDT + k * 32), umma_desc_sw32(sTab + C::kTabMain + k * 512), idesc_ox, 1);
            }
          } else {
            ts(tTO, tS, sDO);                                                  // dV = P^T dO
            ts(tDK, tDP, sQ);                                                  // dK = dS^T Q
          }
          tc_commit(bar(O_FULL));
          if (ps == 3) tc_commit(bar(OPS_EMPTY));
        }
        __syncwarp();
      }
    }
  } else if (warp == 10) {
    // ===================== pad-token fixer: out-of-grid window tokens get k = b_k, v = b_v =====================
    uint32_t cnt = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++cnt) {
      int h, wy, wx, f;
      decode(u, h, wy, wx, f);
      const bool edge = (wy + 1) * kBS > p.G || (wx + 1) * kBS > p.G;
      mbar_wait(bar(LOAD_FULL), cnt & 1u);
      if (edge) {
#pragma unroll 1
        for (int which = 1; which <= 2; ++which) {             // 1: K, 2: V
          const uint32_t base = which == 1 ? sK : sV;
          const __nv_bfloat16* bsrc = p.qkv_bias + which * D + h * HD;
          for (int r = lane; r < kBQ; r += 32) {
            if (wy * kBS + r / kBS >= p.G || wx * kBS + r % kBS >= p.G) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(bsrc + c * 8));
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + r * 128 + ((c ^ (r & 7)) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
              }
              if (kX) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const uint4 v = __ldg(reinterpret_cast<const uint4*>(bsrc + 64 + c * 8));
                  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + C::kMain + r * 32 + ((c ^ ((r >> 2) & 1)) << 4)), "r"(v.x), "r"(v.y),
                               "r"(v.z), "r"(v.w) : "memory");
                }
              }
            }
          }
        }
        fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(FIX_DONE));
    }
  } else {
    // ===================== elementwise warps: two threads per row =====================
    const int quad = warp & 3;
    const int hs = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int et = (warp - 2) * 32 + lane;                    // 0..255
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;
    constexpr float kL2e = 1.4426950408889634f;
    const float c_l2 = kScale * kL2e;
    auto sync256 = []() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    auto stage_at = [&](int e) { return stage_f[row * 64 + ((((e >> 2) ^ (row & 7)) << 2) | (e & 3))]; };
    auto dt_put = [&](int col, float v) {                     // dT[row, col] (bf16, the SWIZZLE_128B K-major tile the UMMA reads)
      *reinterpret_cast<__nv_bfloat16*>(dt_b + row * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2) = __float2bfloat16(v);
    };
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      int h, wy, wx, f;
      decode(u, h, wy, wx, f);
      // ---------------- query passes ----------------
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int q = t * 128 + row, qc = min(q, kBQ - 1);
        const int qh = qc / kBS, qw = qc % kBS;
        const int gy = wy * kBS + qh, gx = wx * kBS + qw;
        const bool valid = q < kBQ && gy < p.G && gx < p.G;    // rows past the window and out-of-grid queries: lse = D = 0, dO = 0 -> dS = 0
        const size_t tok = (size_t)(f * p.G + gy) * p.G + gx;
        const float my_lse = valid ? __ldg(p.lse + tok * p.heads + h) : 0.f;
        const float my_d = valid ? __ldg(p.dsum + tok * p.heads + h) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(dt_b + et * 64 + i * 16) = make_uint4(0, 0, 0, 0);
        mbar_wait(bar(T_FULL), (uint32_t)t);
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
        sync256();
        float relh[8], relw[14];
        float2 ah2[8], aw2[7];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kh = hs * 8 + i;                          // thread 0 of the row owns key rows 0..7, thread 1 rows 8..13
          relh[i] = kh < kBS ? fmaf(stage_at(qh + (kBS - 1) - kh), kL2e, -my_lse) : 0.f;
          ah2[i] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 14; ++i) relw[i] = stage_at(32 + qw + (kBS - 1) - i) * kL2e;
#pragma unroll
        for (int i = 0; i < 7; ++i) aw2[i] = make_float2(0.f, 0.f);
        if (q < kBK) {                                        // the key passes read bias, lse and D of every query from this (transposed) table
          float* rr = rel_s + q;
          if (hs == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i * kRelStride] = relh[i];
#pragma unroll
            for (int i = 0; i < 14; ++i) rr[(14 + i) * kRelStride] = relw[i];
            rr[28 * kRelStride] = -my_d;
          } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) rr[(8 + i) * kRelStride] = relh[i];
          }
        }
        mbar_wait(bar(S_FULL), (uint32_t)t);
        tc_fence_after();
        if (hs == 0) {
          wb_q_chunk<0, 32, 0, 0>(tS, tDP, tlane, relh, relw, c_l2, kScale, my_d, ah2, aw2);
          wb_q_chunk<32, 32, 16, 0>(tS, tDP, tlane, relh, relw, c_l2, kScale, my_d, ah2, aw2);
          wb_q_chunk<64, 32, 32, 0>(tS, tDP, tlane, relh, relw, c_l2, kScale, my_d, ah2, aw2);
          wb_q_chunk<96, 16, 48, 0>(tS, tDP, tlane, relh, relw, c_l2, kScale, my_d, ah2, aw2);
        } else {
          wb_q_chunk<176, 32, 192, 8>(tS, tDP, tlane, relh, relw, c_l2, kScale, my_d, ah2, aw2);
          wb_q_chunk<144, 32, 176, 8>(tS, tDP, tlane, relh, relw, c_l2, kScale, my_d, ah2, aw2);
          wb_q_chunk<112, 32, 160, 8>(tS, tDP, tlane, relh, relw, c_l2, kScale, my_d, ah2, aw2);
        }
        float ah[8], aw[14];
#pragma unroll
        for (int i = 0; i < 8; ++i) ah[i] = ah2[i].x + ah2[i].y;
#pragma unroll
        for (int i = 0; i < 7; ++i) { aw[2 * i] = aw2[i].x; aw[2 * i + 1] = aw2[i].y; }
#pragma unroll
        for (int i = 0; i < 7; ++i) xch_f[(hs * 7 + i) * 128 + row] = aw[hs == 0 ? 7 + i : i];     // the partial sums the partner thread finishes
        sync256();
        // dT[q, qh + 13 - kh] = A_h[q, kh],  dT[q, 32 + qw + 13 - kw] = A_w[q, kw]: the cotangent of T = Q Tab^T
        if (hs == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dt_put(qh + (kBS - 1) - i, ah[i]);
#pragma unroll
          for (int i = 0; i < 7; ++i) dt_put(32 + qw + (kBS - 1) - i, aw[i] + xch_f[(7 + i) * 128 + row]);
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i) dt_put(qh + (kBS - 1) - (8 + i), ah[i]);
#pragma unroll
          for (int i = 7; i < 14; ++i) dt_put(32 + qw + (kBS - 1) - i, aw[i] + xch_f[(i - 7) * 128 + row]);
        }
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(P_FULL));
        // ---- dQ
        mbar_wait(bar(O_FULL), (uint32_t)t);
        tc_fence_after();
        uint32_t r[32], rx[8];
        tmem_ld_x32(tTO + hs * 32 + tlane, r);
        if (kX) tmem_ld_x8(tTO + 64 + hs * 8 + tlane, rx);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(ACC_READ));
        if (valid) {
          __nv_bfloat16* o = p.dqkv + tok * (3 * D) + h * HD;
          wb_store32(o + hs * 32, r, 1.0f);
          if (kX) wb_store8(o + 64 + hs * 8, rx, 1.0f);
        }
      }
      sync256();                                              // the bias table of this unit is complete
      // ---------------- key passes ----------------
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int k = t * 128 + row, kc = min(k, kBQ - 1);
        const int kh = kc / kBS, kw = kc % kBS;
        const int gy = wy * kBS + kh, gx = wx * kBS + kw;
        const bool valid = k < kBQ && gy < p.G && gx < p.G;    // pad keys take part in the softmax but their cotangents reach only the frozen biases
        const size_t tok = (size_t)(f * p.G + gy) * p.G + gx;
        const float *rel_kh = rel_s + kh * kRelStride, *rel_kw = rel_s + (14 + kw) * kRelStride, *rel_d = rel_s + 28 * kRelStride;
        mbar_wait(bar(S_FULL), (uint32_t)t);
        tc_fence_after();
        if (hs == 0) {
          wb_k_chunk<0, 32, 0>(tS, tDP, tlane, rel_kh, rel_kw, rel_d, c_l2);
          wb_k_chunk<32, 32, 16>(tS, tDP, tlane, rel_kh, rel_kw, rel_d, c_l2);
          wb_k_chunk<64, 32, 32>(tS, tDP, tlane, rel_kh, rel_kw, rel_d, c_l2);
          wb_k_chunk<96, 16, 48>(tS, tDP, tlane, rel_kh, rel_kw, rel_d, c_l2);
        } else {
          wb_k_chunk<176, 32, 192>(tS, tDP, tlane, rel_kh, rel_kw, rel_d, c_l2);
          wb_k_chunk<144, 32, 176>(tS, tDP, tlane, rel_kh, rel_kw, rel_d, c_l2);
          wb_k_chunk<112, 32, 160>(tS, tDP, tlane, rel_kh, rel_kw, rel_d, c_l2);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(P_FULL));
        // ---- dV, dK
        mbar_wait(bar(O_FULL), (uint32_t)t);
        tc_fence_after();
        uint32_t rv[32], rk[32], rvx[8], rkx[8];
        tmem_ld_x32(tTO + hs * 32 + tlane, rv);
        tmem_ld_x32(tDK + hs * 32 + tlane, rk);
        if (kX) { tmem_ld_x8(tTO + 64 + hs * 8 + tlane, rvx); tmem_ld_x8(tDK + 64 + hs * 8 + tlane, rkx); }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(ACC_READ));
        if (valid) {
          __nv_bfloat16* o = p.dqkv + tok * (3 * D) + h * HD;
          wb_store32(o + D + hs * 32, rk, kScale);
          wb_store32(o + 2 * D + hs * 32, rv, 1.0f);
          if (kX) { wb_store8(o + D + 64 + hs * 8, rkx, kScale); wb_store8(o + 2 * D + 64 + hs * 8, rvx, 1.0f); }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_tmap_bf16_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows);
int make_tmap_bf16_nd(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box);

// called from run_attn_bwd (attention_bwd.cu) for windowed layers with head dim 64 / 80 when the forward saved its log-sum-exp.
// Rh, Rw: bf16 [27, HD]; lse, dsum: fp32 [M, heads]; writes all three slots of dqkv [M, 3*D] for every token.
template <int HD>
static int launch_window_bwd(const void* qkv, const void* qkv_bias, const void* Rh, const void* Rw, const void* dO, const float* lse, const float* dsum,
                             void* dqkv, int F, int G, int heads, cudaStream_t st) {
  using C = WinBwdCfg<HD>;
  const int D = heads * HD;
  WinBwdTmaps tm;
  int rc;
  uint64_t dims[4] = {(uint64_t)3 * D, (uint64_t)G, (uint64_t)G, (uint64_t)F};
  uint64_t dims_o[4] = {(uint64_t)D, (uint64_t)G, (uint64_t)G, (uint64_t)F};
  uint32_t box[4] = {64, kBS, kBS, 1};
  if ((rc = make_tmap_bf16_nd(&tm.qkv, qkv, 4, dims, box))) return rc;
  if ((rc = make_tmap_bf16_nd(&tm.dO, dO, 4, dims_o, box))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.rh, Rh, HD, 2 * kBS - 1, 64, 32))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm.rw, Rw, HD, 2 * kBS - 1, 64, 32))) return rc;
  if (HD > 64) {
    uint32_t boxx[4] = {16, kBS, kBS, 1};
    if ((rc = make_tmap_bf16_nd(&tm.qkv_x, qkv, 4, dims, boxx))) return rc;
    if ((rc = make_tmap_bf16_nd(&tm.dO_x, dO, 4, dims_o, boxx))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.rh_x, Rh, HD, 2 * kBS - 1, 16, 32))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.rw_x, Rw, HD, 2 * kBS - 1, 16, 32))) return rc;
  } else {
    tm.qkv_x = tm.qkv; tm.dO_x = tm.dO; tm.rh_x = tm.rh; tm.rw_x = tm.rw;
  }
  WinBwdParams p;
  p.qkv_bias = reinterpret_cast<const __nv_bfloat16*>(qkv_bias);
  p.lse = lse; p.dsum = dsum;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.G = G; p.heads = heads; p.nW = (G + kBS - 1) / kBS;
  p.units = F * p.nW * p.nW * heads;
  static_assert(C::kSmem <= 232448, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(attn_window_bwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
  if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute(%d): %s", C::kSmem, cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.units < sms ? p.units : sms;
  attn_window_bwd_tc_kernel<HD><<<grid, kWbThreads, C::kSmem, st>>>(tm, p);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

int launch_attn_window_bwd_tc(const void* qkv, const void* qkv_bias, const void* Rh, const void* Rw, const void* dO, const float* lse, const float* dsum,
                              void* dqkv, int F, int G, int heads, int hd, cudaStream_t st) {
  return hd == 64 ? launch_window_bwd<64>(qkv, qkv_bias, Rh, Rw, dO, lse, dsum, dqkv, F, G, heads, st)
                  : launch_window_bwd<80>(qkv, qkv_bias, Rh, Rw, dO, lse, dsum, dqkv, F, G, heads, st);
}

}  // namespace grove
