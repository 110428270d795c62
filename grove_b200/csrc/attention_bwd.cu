// Backward of the SAM ViT attention with decomposed relative-position bias (image_encoder.py:301-326, 420-458) for the
// training step of the grounding branch (BASELINE config 4).  The encoder blocks are frozen (train.py:254-255), so only
// d(qkv) is produced — no weight / table gradients.
//
//   S[i,j] = scale q_i.k_j + q_i.Rh[iy-jy+S-1] + q_i.Rw[ix-jx+S-1],  P = softmax_j S,  O = P V
//
// Pipeline (one call = one attention layer, window or global):
//   1. relpos_bias_kernel   rel[i, 0:S] = q_i.Rh[iy-jy+S-1], rel[i, S:2S] = q_i.Rw[ix-jx+S-1]        (fp32, like the reference's einsums)
//      rowdot_kernel        Dsum[i] = dO_i . O_i
//   2. attn_bwd_q_kernel    per 64-query tile: sweep keys for the log-sum-exp, sweep again for dS = P (dP - Dsum);
//                           dq_core = scale dS K ;  A[i, jy] = sum_jx dS, A[i, S+jx] = sum_jy dS   (the rel-pos bias' cotangents)
//   3. attn_bwd_kv_kernel   per 64-key tile (transposed orientation): dV = P^T dO, dK = scale dS^T Q
//   4. relpos_bias_bwd      dq = dq_core + sum_c A[i,c] . R[idx(i,c)]  -> bf16 into the q slot of dqkv
// Steps 2-3 run on warp-level mma.sync (m16n8k16 bf16, fp32 accumulate); scores never reach HBM.  Windowed blocks address the
// UNPARTITIONED token-major tensors; window positions outside the image are the zero-padded tokens of window_partition
// (image_encoder.py:344-348): as keys they carry k = b_k, v = b_v and receive softmax mass, as queries they are skipped.
#include <cstdlib>

#include "grove_b200.h"
#include "mma_sync.cuh"

namespace grove {
using namespace ws;

constexpr float kL2e = 1.4426950408889634f;
typedef __nv_bfloat16 bf16_t;

// region-local index t (row-major in an S x S region at (y0,x0)) -> token id (>= 0), -1 = padded window position, -2 = past the region
template <int S, bool WIN>
struct RegionMap {
  int y0, x0, G;
  __device__ __forceinline__ int tok(int t) const {
    if (t >= S * S) return -2;
    const int gy = y0 + t / S, gx = x0 + t % S;
    if (WIN && (gy >= G || gx >= G)) return -1;
    return gy * G + gx;
  }
};

template <int S>
struct RelStride { static constexpr int v = (S == 14) ? 36 : 2 * S + 4; };   // 16-byte aligned rows, 2*v mod 32 == 8 (bank spread)

template <int HD, class F>
__device__ __forceinline__ void load_tile64(uint32_t sbase, F rowptr, int tid) {
  constexpr int CH = HD / 8;
  for (int i = tid; i < 64 * CH; i += 128) {
    const int r = i / CH, c = i % CH;
    const bf16_t* p = rowptr(r);
    if (p) cp_async16(tile_addr<HD>(sbase, r, c), p + c * 8);
    else st_smem16(tile_addr<HD>(sbase, r, c), make_uint4(0, 0, 0, 0));
  }
}

// acc[j] (16 rows x 64 cols, 8 n-tiles) = A(16 x HD, register fragments) . Y^T  with Y a [64][HD] tile in shared memory
template <int HD>
__device__ __forceinline__ void mma_a_yt(float (&acc)[8][4], const uint32_t (&af)[HD / 16][4], uint32_t sY, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      uint32_t bb[4];
      ldsm_x4(tile_addr<HD>(sY, jp * 16 + (lane & 7) + (lane >> 4) * 8, 2 * ks + ((lane >> 3) & 1)), bb);
      mma16816(acc[2 * jp], af[ks], bb[0], bb[1]);
      mma16816(acc[2 * jp + 1], af[ks], bb[2], bb[3]);
    }
}

// out (16 rows x HD) += P(16 x 64, fp32 accumulator fragments, rounded to bf16) . Y  with Y a [64][HD] tile in shared memory
template <int HD>
__device__ __forceinline__ void mma_p_y(float (&out)[HD / 8][4], const float (&p)[8][4], uint32_t sY, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    pa[1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    pa[2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    pa[3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < HD / 16; ++dp) {
      uint32_t bb[4];
      ldsm_x4_t(tile_addr<HD>(sY, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dp + (lane >> 4)), bb);
      mma16816(out[2 * dp], pa, bb[0], bb[1]);
      mma16816(out[2 * dp + 1], pa, bb[2], bb[3]);
    }
  }
}

// =====================================================================================================================
// Query side: LSE, dq_core, A
// =====================================================================================================================
template <int S, int HD, bool WIN>
__global__ void __launch_bounds__(128, 1)
attn_bwd_q_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ qkv_bias, const bf16_t* __restrict__ dO, const float* __restrict__ rel,
                  const float* __restrict__ Dsum, const float* __restrict__ lse_in, float* __restrict__ lse_out, float* __restrict__ dq_out,
                  float* __restrict__ A_out, int G, int heads) {
  constexpr int NT = (S * S + 63) / 64, KS = HD / 16, NTD = HD / 8, TILEB = 64 * HD * 2, RS = RelStride<S>::v;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = ((smem_u32(smem_raw) + 127u) & ~127u) - smem_u32(smem_raw);
  const uint32_t s0 = smem_u32(smem_raw) + pad;
  const uint32_t sQ = s0, sDO = s0 + TILEB, sK = s0 + 2 * TILEB, sV = s0 + 4 * TILEB;
  float* sRel = reinterpret_cast<float*>(smem_raw + pad + 6 * TILEB);   // [64][RS], pre-multiplied by log2(e)
  float* sA = sRel + 64 * RS;                                           // [64][RS]
  float* sDS = sA + 64 * RS;                                            // [64][65]
  int* sTok = reinterpret_cast<int*>(sDS + 64 * 65);                    // [64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int qt = blockIdx.x, region = blockIdx.y / heads, h = blockIdx.y % heads, f = blockIdx.z;
  const int N = G * G, Dm = heads * HD;
  const size_t tstride = (size_t)3 * Dm;
  RegionMap<S, WIN> map;
  map.G = G;
  if (WIN) { const int nwx = (G + S - 1) / S; map.y0 = (region / nwx) * S; map.x0 = (region % nwx) * S; }
  else { map.y0 = 0; map.x0 = 0; }

  if (tid < 64) sTok[tid] = map.tok(qt * 64 + tid);
  if (!__syncthreads_or(tid < 64 && sTok[tid] >= 0)) return;   // no real query in this tile (also publishes sTok)

  const bf16_t* qkv_f = qkv + (size_t)f * N * tstride + h * HD;
  load_tile64<HD>(sQ, [&](int r) { const int t = sTok[r]; return t >= 0 ? qkv_f + (size_t)t * tstride : nullptr; }, tid);
  load_tile64<HD>(sDO, [&](int r) { const int t = sTok[r]; return t >= 0 ? dO + ((size_t)f * N + t) * Dm + h * HD : nullptr; }, tid);
  cp_async_commit();
  for (int i = tid; i < 64 * 2 * S; i += 128) {
    const int r = i / (2 * S), c = i % (2 * S), t = sTok[r];
    sRel[r * RS + c] = t >= 0 ? rel[(((size_t)f * N + t) * heads + h) * (2 * S) + c] * kL2e : 0.f;
    sA[r * RS + c] = 0.f;
  }
  // S <= 16 (the 14x14 windows): the rel-pos bias cotangents A_h[i, jy] = sum_jx dS, A_w[i, jx] = sum_jy dS are one more tensor-core
  // product per key tile, dS (16 x 64, already in bf16 A fragments for dq) times a ONE-HOT matrix E_kt [64 key columns][32]: column jy of
  // the first 16 and column 16 + jx of the second 16 are 1 for key t = kt*64 + c (exact in bf16; the sums accumulate in fp32 fragments).
  // It replaces a serial two-lanes-per-row loop over the staged dS tile that cost as much as the three real products together.
  constexpr bool kHot = S <= 16;
  const uint32_t sE = smem_u32(sDS);                     // kHot: the dS staging area holds the NT one-hot tiles (4 KB each) instead
  if (kHot) {
    __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(sDS);
    for (int i = tid; i < NT * 64 * 32; i += 128) {
      const int t = (i >> 11) * 64 + ((i >> 5) & 63), col = i & 31;
      const bool one = t < S * S && (col < 16 ? t / S == col : t % S == col - 16);
      e[i] = __float2bfloat16(one ? 1.f : 0.f);
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  const int r0 = warp * 16;
  uint32_t qf[KS][4], dof[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    ldsm_x4(tile_addr<HD>(sQ, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), qf[ks]);
    ldsm_x4(tile_addr<HD>(sDO, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), dof[ks]);
  }
  const int row_l[2] = {r0 + g, r0 + g + 8};
  const int tokr[2] = {sTok[row_l[0]], sTok[row_l[1]]};
  float drow[2];
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) drow[rs] = tokr[rs] >= 0 ? Dsum[((size_t)f * N + tokr[rs]) * heads + h] : 0.f;
  const float scale = rsqrtf((float)HD), scale_l2 = scale * kL2e;

  auto kptr = [&](int kt, int which) {
    return [=](int r) -> const bf16_t* {
      const int t = map.tok(kt * 64 + r);
      if (t >= 0) return qkv_f + (size_t)t * tstride + which * Dm;
      if (t == -1) return qkv_bias + which * Dm + h * HD;
      return nullptr;
    };
  };
  auto score = [&](float (&s)[8][4], int kt) {   // s <- log2-domain logits, -inf past the region
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rs = e >> 1, t = kt * 64 + 8 * j + 2 * t4 + (e & 1);
        float v = -INFINITY;
        if (t < S * S) {
          const int jy = t / S, jx = t - jy * S;
          v = s[j][e] * scale_l2 + sRel[row_l[rs] * RS + jy] + sRel[row_l[rs] * RS + S + jx];
        }
        s[j][e] = v;
      }
  };

  // ---- sweep 1: log-sum-exp per query row (skipped when the forward kernel saved it: lse_in)
  float lse2[2];
  if (lse_in != nullptr) {
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) lse2[rs] = tokr[rs] >= 0 ? lse_in[((size_t)f * N + tokr[rs]) * heads + h] : INFINITY;
  } else {
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  load_tile64<HD>(sK, kptr(0, 1), tid);
  cp_async_commit();
  for (int kt = 0; kt < NT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < NT) { load_tile64<HD>(sK + (buf ^ 1) * TILEB, kptr(kt + 1, 1), tid); cp_async_commit(); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    float s[8][4];
    mma_a_yt<HD>(s, qf, sK + buf * TILEB, lane);
    score(s, kt);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      mx[rs] = fmaxf(mx[rs], __shfl_xor_sync(0xffffffffu, mx[rs], 1));
      mx[rs] = fmaxf(mx[rs], __shfl_xor_sync(0xffffffffu, mx[rs], 2));
      const float mn = fmaxf(m_run[rs], mx[rs]);
      l_run[rs] *= exp2f(m_run[rs] - mn);
      m_run[rs] = mn;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) l_run[e >> 1] += exp2f(s[j][e] - m_run[e >> 1]);
    __syncthreads();
  }
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    l_run[rs] += __shfl_xor_sync(0xffffffffu, l_run[rs], 1);
    l_run[rs] += __shfl_xor_sync(0xffffffffu, l_run[rs], 2);
    lse2[rs] = tokr[rs] >= 0 ? m_run[rs] + log2f(l_run[rs]) : INFINITY;
    if (t4 == 0 && tokr[rs] >= 0) lse_out[((size_t)f * N + tokr[rs]) * heads + h] = lse2[rs];
  }
  }

  // ---- sweep 2: dS, dq_core, A
  float dq[NTD][4], acc_a[4][4];
#pragma unroll
  for (int j = 0; j < NTD; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) acc_a[j][0] = acc_a[j][1] = acc_a[j][2] = acc_a[j][3] = 0.f;
  load_tile64<HD>(sK, kptr(0, 1), tid);
  load_tile64<HD>(sV, kptr(0, 2), tid);
  cp_async_commit();
  for (int kt = 0; kt < NT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < NT) {
      load_tile64<HD>(sK + (buf ^ 1) * TILEB, kptr(kt + 1, 1), tid);
      load_tile64<HD>(sV + (buf ^ 1) * TILEB, kptr(kt + 1, 2), tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else cp_async_wait<0>();
    __syncthreads();
    float s[8][4], dp[8][4];
    mma_a_yt<HD>(s, qf, sK + buf * TILEB, lane);
    score(s, kt);
    mma_a_yt<HD>(dp, dof, sV + buf * TILEB, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rs = e >> 1;
        const float p = exp2f(s[j][e] - lse2[rs]);   // 0 past the region (s = -inf) and for non-real rows (lse = +inf)
        const float ds = p * (dp[j][e] - drow[rs]);
        s[j][e] = ds;
        if (!kHot) sDS[row_l[rs] * 65 + 8 * j + 2 * t4 + (e & 1)] = ds;
      }
    mma_p_y<HD>(dq, s, sK + buf * TILEB, lane);
    if (kHot) mma_p_y<32>(acc_a, s, sE + kt * 4096, lane);
    __syncwarp();
    if (!kHot) {  // rel-pos cotangents: two lanes per row — lane&1 == 0 sums runs of equal jy, lane&1 == 1 scatters by jx
      const int r = r0 + (lane >> 1);
      const float* dsr = sDS + r * 65;
      float* ar = sA + r * RS;
      const int tbase = kt * 64;
      const int nval = min(64, S * S - tbase);
      if ((lane & 1) == 0) {
        int cur = tbase / S;
        float sum = 0.f;
        for (int c = 0; c < nval; ++c) {
          const int jy = (tbase + c) / S;
          if (jy != cur) { ar[cur] += sum; sum = 0.f; cur = jy; }
          sum += dsr[c];
        }
        ar[cur] += sum;
      } else {
        for (int c = 0; c < nval; ++c) ar[S + (tbase + c) % S] += dsr[c];
      }
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int rs = 0; rs < 2; ++rs)
    if (tokr[rs] >= 0) {
      float* o = dq_out + ((size_t)f * N + tokr[rs]) * Dm + h * HD + 2 * t4;
#pragma unroll
      for (int j = 0; j < NTD; ++j) *reinterpret_cast<float2*>(o + 8 * j) = make_float2(dq[j][2 * rs] * scale, dq[j][2 * rs + 1] * scale);
    }
  if (kHot) {
#pragma unroll
    for (int rs = 0; rs < 2; ++rs)
      if (tokr[rs] >= 0) {
        float* ab = A_out + (((size_t)f * N + tokr[rs]) * heads + h) * (2 * S);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int col = 8 * (nt & 1) + 2 * t4 + e2;       // jy (nt < 2) or jx (nt >= 2)
            if (col < S) ab[(nt < 2 ? 0 : S) + col] = acc_a[nt][2 * rs + e2];
          }
      }
  } else {
    for (int i = tid; i < 64 * 2 * S; i += 128) {
      const int r = i / (2 * S), c = i % (2 * S), t = sTok[r];
      if (t >= 0) A_out[(((size_t)f * N + t) * heads + h) * (2 * S) + c] = sA[r * RS + c];
    }
  }
}


// =====================================================================================================================
// Query side, global attention fast path (S = G, 64 % G == 0: a 64-key tile is 64/G whole grid rows).  The log-sum-exp comes
// from the forward kernel (grove_attn_global_relpos_fwd_lse), so there is a single key sweep; rel_w lives in registers for the
// whole kernel (each thread always sees the same key columns), rel_h is one value per (row, tile row); the bias cotangents
// need no scratch: A_w accumulates in registers at fixed fragment positions, A_h is a quad-reduced row sum per tile written
// straight to global.  65-80 KB of shared memory and <= 255 registers: two CTAs per SM.
// =====================================================================================================================
template <int G, int HD>
__global__ void __launch_bounds__(128, 2)
attn_bwd_q_global_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ dO, const float* __restrict__ rel, const float* __restrict__ Dsum,
                         const float* __restrict__ lse, float* __restrict__ dq_out, float* __restrict__ A_out, int heads) {
  constexpr int S = G, N = G * G, NT = N / 64, KS = HD / 16, NTD = HD / 8, TILEB = 64 * HD * 2, RPT = 64 / G, RH = G + 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = ((smem_u32(smem_raw) + 127u) & ~127u) - smem_u32(smem_raw);
  const uint32_t s0 = smem_u32(smem_raw) + pad;
  const uint32_t sQ = s0, sDO = s0 + TILEB, sK = s0 + 2 * TILEB, sV = s0 + 4 * TILEB;
  float* sRelH = reinterpret_cast<float*>(smem_raw + pad + 6 * TILEB);   // [64][RH], pre-multiplied by log2(e)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int qt = blockIdx.x, h = blockIdx.y, f = blockIdx.z;
  const int Dm = heads * HD;
  const size_t tstride = (size_t)3 * Dm;
  const bf16_t* qkv_f = qkv + (size_t)f * N * tstride + h * HD;
  const size_t row0 = (size_t)f * N + qt * 64;                 // first query token (global row index) of this tile

  load_tile64<HD>(sQ, [&](int r) { return qkv_f + (size_t)(qt * 64 + r) * tstride; }, tid);
  load_tile64<HD>(sDO, [&](int r) { return dO + (row0 + r) * Dm + h * HD; }, tid);
  auto load_kv = [&](int kt, int buf) {
    load_tile64<HD>(sK + buf * TILEB, [&](int r) { return qkv_f + (size_t)(kt * 64 + r) * tstride + Dm; }, tid);
    load_tile64<HD>(sV + buf * TILEB, [&](int r) { return qkv_f + (size_t)(kt * 64 + r) * tstride + 2 * Dm; }, tid);
    cp_async_commit();
  };
  load_kv(0, 0);      // same commit group as Q / dO
  for (int i = tid; i < 64 * G; i += 128) {
    const int r = i / G, c = i % G;
    sRelH[r * RH + c] = rel[((row0 + r) * heads + h) * (2 * S) + c] * kL2e;
  }
  cp_async_wait<0>();
  __syncthreads();

  const int r0 = warp * 16;
  uint32_t qf[KS][4], dof[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    ldsm_x4(tile_addr<HD>(sQ, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), qf[ks]);
    ldsm_x4(tile_addr<HD>(sDO, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), dof[ks]);
  }
  const int row_l[2] = {r0 + g, r0 + g + 8};
  float drow[2], lse2[2], relw[2][16], aw[2][16];
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    const size_t rh = (row0 + row_l[rs]) * heads + h;
    drow[rs] = Dsum[rh];
    lse2[rs] = lse[rh];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        relw[rs][2 * j + e] = rel[rh * (2 * S) + S + (8 * j + 2 * t4 + e) % G] * kL2e;
        aw[rs][2 * j + e] = 0.f;
      }
  }
  const float scale = rsqrtf((float)HD), scale_l2 = scale * kL2e;
  float dq[NTD][4];
#pragma unroll
  for (int j = 0; j < NTD; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;

  for (int kt = 0; kt < NT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < NT) { load_kv(kt + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    float s[8][4], dp[8][4];
    mma_a_yt<HD>(s, qf, sK + buf * TILEB, lane);
    mma_a_yt<HD>(dp, dof, sV + buf * TILEB, lane);
    float ah[2][RPT];
#pragma unroll
    for (int rs = 0; rs < 2; ++rs)
#pragma unroll
      for (int q = 0; q < RPT; ++q) ah[rs][q] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rs = e >> 1, sub = (RPT == 1) ? 0 : (j * 8) / G;          // which grid row of the tile this column belongs to
        const float v = s[j][e] * scale_l2 + sRelH[row_l[rs] * RH + kt * RPT + sub] + relw[rs][2 * j + (e & 1)];
        const float ds = exp2f(v - lse2[rs]) * (dp[j][e] - drow[rs]);
        s[j][e] = ds;
        aw[rs][2 * j + (e & 1)] += ds;
        ah[rs][sub] += ds;
      }
    mma_p_y<HD>(dq, s, sK + buf * TILEB, lane);
#pragma unroll
    for (int rs = 0; rs < 2; ++rs)
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        float a = ah[rs][q];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (t4 == 0) A_out[((row0 + row_l[rs]) * heads + h) * (2 * S) + kt * RPT + q] = a;
      }
    __syncthreads();
  }
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    float* o = dq_out + (row0 + row_l[rs]) * Dm + h * HD + 2 * t4;
#pragma unroll
    for (int j = 0; j < NTD; ++j) *reinterpret_cast<float2*>(o + 8 * j) = make_float2(dq[j][2 * rs] * scale, dq[j][2 * rs + 1] * scale);
    float* ao = A_out + ((row0 + row_l[rs]) * heads + h) * (2 * S) + S;
    if (G == 64) {
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(ao + 8 * j + 2 * t4) = make_float2(aw[rs][2 * j], aw[rs][2 * j + 1]);
    } else {   // G == 32: columns c and c + 32 of a tile are the same jx
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float2*>(ao + 8 * j + 2 * t4) = make_float2(aw[rs][2 * j] + aw[rs][2 * (j + 4)], aw[rs][2 * j + 1] + aw[rs][2 * (j + 4) + 1]);
    }
  }
}

// =====================================================================================================================
// Key/value side (transposed orientation: rows = keys, columns = queries)
// =====================================================================================================================
// GFAST (global attention, 64 % S == 0): a 64-key tile is 64/S whole grid rows, so of each streamed query's 2S bias values only the
// S rel_w entries and the tile's 64/S rel_h entries are fetched (half the L2 traffic of the generic layout), two CTAs per SM.
template <int S, int HD, bool WIN, bool GFAST>
__global__ void __launch_bounds__(128, GFAST ? 2 : 1)
attn_bwd_kv_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ qkv_bias, const bf16_t* __restrict__ dO, const float* __restrict__ rel,
                   const float* __restrict__ Dsum, const float* __restrict__ lse, bf16_t* __restrict__ dqkv, int G, int heads) {
  constexpr int NT = (S * S + 63) / 64, KS = HD / 16, NTD = HD / 8, TILEB = 64 * HD * 2, RS = GFAST ? S + 4 : RelStride<S>::v;
  constexpr int RPT = GFAST ? 64 / S : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = ((smem_u32(smem_raw) + 127u) & ~127u) - smem_u32(smem_raw);
  const uint32_t s0 = smem_u32(smem_raw) + pad;
  const uint32_t sK = s0, sV = s0 + TILEB, sQ = s0 + 2 * TILEB, sDO = s0 + 4 * TILEB;
  float* sRel = reinterpret_cast<float*>(smem_raw + pad + 6 * TILEB);   // [2][64][RS] raw bias rows of the streamed queries
  float* sLse = sRel + 2 * 64 * RS;                                     // [2][64]
  float* sD = sLse + 128;                                               // [2][64]
  int* sTok = reinterpret_cast<int*>(sD + 128);                         // [64] key tokens
  const uint32_t sRel_u = s0 + 6 * TILEB;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int kt = blockIdx.x, region = blockIdx.y / heads, h = blockIdx.y % heads, f = blockIdx.z;
  const int N = G * G, Dm = heads * HD;
  const size_t tstride = (size_t)3 * Dm;
  RegionMap<S, WIN> map;
  map.G = G;
  if (WIN) { const int nwx = (G + S - 1) / S; map.y0 = (region / nwx) * S; map.x0 = (region % nwx) * S; }
  else { map.y0 = 0; map.x0 = 0; }

  if (tid < 64) sTok[tid] = map.tok(kt * 64 + tid);
  if (!__syncthreads_or(tid < 64 && sTok[tid] >= 0)) return;   // only padded / nonexistent keys: nothing to write

  const bf16_t* qkv_f = qkv + (size_t)f * N * tstride + h * HD;
  // pad keys would only produce gradients of the frozen qkv bias: load them as zeros, never stored
  load_tile64<HD>(sK, [&](int r) { const int t = sTok[r]; return t >= 0 ? qkv_f + (size_t)t * tstride + Dm : nullptr; }, tid);
  load_tile64<HD>(sV, [&](int r) { const int t = sTok[r]; return t >= 0 ? qkv_f + (size_t)t * tstride + 2 * Dm : nullptr; }, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int r0 = warp * 16;
  uint32_t kf[KS][4], vf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    ldsm_x4(tile_addr<HD>(sK, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), kf[ks]);
    ldsm_x4(tile_addr<HD>(sV, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), vf[ks]);
  }
  int jy[2], jx[2];
  bool kvalid[2];
  int tokk[2];
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    const int t = kt * 64 + r0 + g + 8 * rs;
    tokk[rs] = sTok[r0 + g + 8 * rs];
    kvalid[rs] = tokk[rs] >= 0;
    jy[rs] = min(t / S, S - 1);
    jx[rs] = t % S;
  }
  const float scale = rsqrtf((float)HD), scale_l2 = scale * kL2e;
  (void)qkv_bias;

  auto load_q = [&](int qt, int buf) {
    load_tile64<HD>(sQ + buf * TILEB, [&](int r) { const int t = map.tok(qt * 64 + r); return t >= 0 ? qkv_f + (size_t)t * tstride : nullptr; }, tid);
    load_tile64<HD>(sDO + buf * TILEB, [&](int r) { const int t = map.tok(qt * 64 + r); return t >= 0 ? dO + ((size_t)f * N + t) * Dm + h * HD : nullptr; },
                    tid);
    constexpr int C4 = (GFAST ? S : 2 * S) / 4;   // 16-byte chunks per bias row (2S floats, or only the S rel_w entries)
    for (int i = tid; i < 64 * C4; i += 128) {
      const int r = i / C4, c = i % C4, t = map.tok(qt * 64 + r);
      const uint32_t dst = sRel_u + ((buf * 64 + r) * RS + c * 4) * 4;
      if (t >= 0) cp_async16(dst, rel + (((size_t)f * N + t) * heads + h) * (2 * S) + (GFAST ? S : 0) + c * 4);
      else st_smem16(dst, make_uint4(0, 0, 0, 0));
    }
    if (GFAST)   // the rel_h entries of this key tile's grid rows
      for (int i = tid; i < 64 * RPT; i += 128) {
        const int r = i / RPT, q = i % RPT;
        sRel[(buf * 64 + r) * RS + S + q] = rel[(((size_t)f * N + qt * 64 + r) * heads + h) * (2 * S) + kt * RPT + q];
      }
    if (tid < 64) {
      const int t = map.tok(qt * 64 + tid);
      sLse[buf * 64 + tid] = t >= 0 ? lse[((size_t)f * N + t) * heads + h] : INFINITY;
      sD[buf * 64 + tid] = t >= 0 ? Dsum[((size_t)f * N + t) * heads + h] : 0.f;
    }
    cp_async_commit();
  };

  float dk[NTD][4], dv[NTD][4];
#pragma unroll
  for (int j = 0; j < NTD; ++j) dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  __syncthreads();   // K/V staging fully read before anything else is overwritten (buffers are distinct, kept for clarity)
  load_q(0, 0);
  for (int qt = 0; qt < NT; ++qt) {
    const int buf = qt & 1;
    if (qt + 1 < NT) { load_q(qt + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    float s[8][4], dp[8][4];
    mma_a_yt<HD>(s, kf, sQ + buf * TILEB, lane);     // S^T = K Q^T
    mma_a_yt<HD>(dp, vf, sDO + buf * TILEB, lane);   // dP^T = V dO^T
    const float* relb = sRel + buf * 64 * RS;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rs = e >> 1, c = 8 * j + 2 * t4 + (e & 1);
        const float v = s[j][e] * scale_l2 + (GFAST ? relb[c * RS + S + (jy[rs] - kt * RPT)] + relb[c * RS + jx[rs]]
                                                    : relb[c * RS + jy[rs]] + relb[c * RS + S + jx[rs]]) * kL2e;
        const float p = kvalid[rs] ? exp2f(v - sLse[buf * 64 + c]) : 0.f;
        s[j][e] = p;
        dp[j][e] = p * (dp[j][e] - sD[buf * 64 + c]);
      }
    mma_p_y<HD>(dv, s, sDO + buf * TILEB, lane);   // dV += P^T dO
    mma_p_y<HD>(dk, dp, sQ + buf * TILEB, lane);   // dK += dS^T Q
    __syncthreads();
  }
#pragma unroll
  for (int rs = 0; rs < 2; ++rs)
    if (kvalid[rs]) {
      bf16_t* o = dqkv + ((size_t)f * N + tokk[rs]) * tstride + h * HD + 2 * t4;
#pragma unroll
      for (int j = 0; j < NTD; ++j) {
        *reinterpret_cast<uint32_t*>(o + Dm + 8 * j) = pack_bf16(dk[j][2 * rs] * scale, dk[j][2 * rs + 1] * scale);
        *reinterpret_cast<uint32_t*>(o + 2 * Dm + 8 * j) = pack_bf16(dv[j][2 * rs], dv[j][2 * rs + 1]);
      }
    }
}

// =====================================================================================================================
// rel-pos bias terms and their backward; Dsum
// =====================================================================================================================
// CTA = 64 consecutive tokens x one head.  MODE 0: rel[i, c] = q_i . R[idx(i,c)].   MODE 1: dq_i = dq_core_i + sum_c A[i,c] R[idx(i,c)].
template <int S, int HD, bool WIN, int MODE>
__global__ void __launch_bounds__(256) relpos_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ Rh, const bf16_t* __restrict__ Rw,
                                                     float* __restrict__ rel, const float* __restrict__ A, const float* __restrict__ dq_core,
                                                     bf16_t* __restrict__ dqkv, int G, int heads) {
  constexpr int TR = 2 * S - 1, HP = HD + 1;
  extern __shared__ float smf[];
  float* sT = smf;                  // [2*TR][HP]  rows 0..TR-1 = Rh, TR.. = Rw
  float* sX = sT + 2 * TR * HP;     // MODE 0: q rows [64][HP];  MODE 1: A rows [64][2S+1]
  const int tid = threadIdx.x, h = blockIdx.y;
  const long long row0 = (long long)blockIdx.x * 64;
  const int N = G * G, Dm = heads * HD;
  for (int i = tid; i < 2 * TR * HD; i += 256) {
    const int r = i / HD, d = i % HD;
    sT[r * HP + d] = __bfloat162float(r < TR ? Rh[r * HD + d] : Rw[(r - TR) * HD + d]);
  }
  if (MODE == 0) {
    for (int i = tid; i < 64 * HD; i += 256) {
      const int r = i / HD, d = i % HD;
      sX[r * HP + d] = __bfloat162float(qkv[(size_t)(row0 + r) * 3 * Dm + h * HD + d]);
    }
  } else {
    for (int i = tid; i < 64 * 2 * S; i += 256) {
      const int r = i / (2 * S), c = i % (2 * S);
      sX[r * (2 * S + 1) + c] = A[((size_t)(row0 + r) * heads + h) * (2 * S) + c];
    }
  }
  __syncthreads();
  if (MODE == 0) {
    // one thread = four bias columns (c, c + S/2, c + S, c + 3S/2) of one query: q[d] is read once for four dot products (the kernel is
    // bound by shared-memory loads: 5 instead of 8 per four multiply-adds); neighbouring lanes keep neighbouring table rows (no conflicts)
    static_assert((2 * S) % 4 == 0, "bias columns come in groups of four");
    constexpr int CQ = 2 * S / 4;
    for (int i = tid; i < 64 * CQ; i += 256) {
      const int r = i / CQ, c0 = i % CQ;
      const int tok = (int)((row0 + r) % N), gy = tok / G, gx = tok % G;
      const int ly = WIN ? gy % S : gy, lx = WIN ? gx % S : gx;
      const float* q = sX + r * HP;
      const float* t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j * CQ;
        t[j] = sT + (c < S ? (ly - c + S - 1) : TR + (lx - (c - S) + S - 1)) * HP;
      }
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int d = 0; d < HD; ++d) {
        const float qd = q[d];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] += qd * t[j][d];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) rel[((size_t)(row0 + r) * heads + h) * (2 * S) + c0 + j * CQ] = a[j];
    }
  } else {
    for (int i = tid; i < 64 * HD; i += 256) {
      const int r = i / HD, d = i % HD;
      const int tok = (int)((row0 + r) % N), gy = tok / G, gx = tok % G;
      const int ly = WIN ? gy % S : gy, lx = WIN ? gx % S : gx;
      const float* a = sX + r * (2 * S + 1);
      float acc = dq_core[(size_t)(row0 + r) * Dm + h * HD + d];
#pragma unroll 4
      for (int c = 0; c < S; ++c) acc += a[c] * sT[(ly - c + S - 1) * HP + d] + a[S + c] * sT[(TR + lx - c + S - 1) * HP + d];
      dqkv[(size_t)(row0 + r) * 3 * Dm + h * HD + d] = __float2bfloat16(acc);
    }
  }
}

// Dsum[row, h] = sum_d dO[row, h, d] * O[row, h, d]
__global__ void rowdot_kernel(const bf16_t* __restrict__ a, const bf16_t* __restrict__ b, float* __restrict__ out, long long rows_heads, int HD) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows_heads) return;
  const uint4* pa = reinterpret_cast<const uint4*>(a + i * HD);
  const uint4* pb = reinterpret_cast<const uint4*>(b + i * HD);
  float s = 0.f;
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 x = pa[c], y = pb[c];
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 u = unpack_bf16(xs[j]), v = unpack_bf16(ys[j]);
      s += u.x * v.x + u.y * v.y;
    }
  }
  out[i] = s;
}

// attention_bwd_tc.cu
int launch_attn_bwd_kv_tc(const void* qkv, const void* dO, const float* rel, const float* lse, const float* dsum, float* aux, void* dqkv, int F,
                          int G, int heads, int hd, cudaStream_t st);
int launch_attn_bwd_q_tc(const void* qkv, const void* dO, const void* Rh, const void* Rw, float* rel, const float* lse, const float* dsum, void* dqkv,
                         int F, int G, int heads, int hd, cudaStream_t st);

// attention_win_bwd_tc.cu
int launch_attn_window_bwd_tc(const void* qkv, const void* qkv_bias, const void* Rh, const void* Rw, const void* dO, const float* lse, const float* dsum,
                              void* dqkv, int F, int G, int heads, int hd, cudaStream_t st);

template <int S, int HD, bool WIN>
static int run_attn_bwd(const bf16_t* qkv, const bf16_t* qkv_bias, const bf16_t* Rh, const bf16_t* Rw, const bf16_t* O, const bf16_t* dO, bf16_t* dqkv,
                        float* ws, const float* lse_fwd, int F, int G, int heads, cudaStream_t st) {
  constexpr int NT = (S * S + 63) / 64, TILEB = 64 * HD * 2, RS = RelStride<S>::v;
  constexpr bool GFAST = !WIN && (64 % S == 0) && S >= 32;
  const long long M = (long long)F * G * G;
  float* rel = ws;
  float* A = rel + M * heads * 2 * S;
  float* Dsum = A + M * heads * 2 * S;
  float* lse = Dsum + M * heads;
  float* dqc = lse + M * heads;
  const int regions = WIN ? ((G + S - 1) / S) * ((G + S - 1) / S) : 1;
  const int smem_rel = (2 * (2 * S - 1) * (HD + 1) + 64 * (HD + 1 > 2 * S + 1 ? HD + 1 : 2 * S + 1)) * (int)sizeof(float);
  const int smem_q = 128 + 6 * TILEB + (2 * 64 * RS + 64 * 65) * 4 + 64 * 4;
  const int smem_qf = 128 + 6 * TILEB + 64 * (S + 1) * 4;
  const int smem_kv = 128 + 6 * TILEB + (2 * 64 * (GFAST ? S + 4 : RS) + 256) * 4 + 64 * 4;
  static GrovePerDeviceOnce attr;
  if (attr.first_time()) {
    cudaFuncSetAttribute(relpos_kernel<S, HD, WIN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_rel);
    cudaFuncSetAttribute(relpos_kernel<S, HD, WIN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_rel);
    cudaFuncSetAttribute(attn_bwd_q_kernel<S, HD, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q);
    cudaFuncSetAttribute(attn_bwd_kv_kernel<S, HD, WIN, GFAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kv);
    if constexpr (GFAST) cudaFuncSetAttribute(attn_bwd_q_global_kernel<S, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_qf);
  }
  if constexpr (WIN && (HD == 64 || HD == 80) && S == 14) {
    // windows with the forward's log-sum-exp at hand: the whole layer in one persistent tcgen05 kernel (attention_win_bwd_tc.cu)
    static const bool use_mma_sync = []() { const char* e = getenv("GROVE_BWD_MMA_SYNC"); return e && e[0] == '1'; }();
    if (lse_fwd != nullptr && !use_mma_sync) {
      rowdot_kernel<<<(unsigned)((M * heads + 255) / 256), 256, 0, st>>>(dO, O, Dsum, M * heads, HD);
      grove_count_launch();
      return launch_attn_window_bwd_tc(qkv, qkv_bias, Rh, Rw, dO, lse_fwd, Dsum, dqkv, F, G, heads, HD, st);
    }
  }
  if constexpr (GFAST && (HD == 64 || HD == 80) && (S == 64 || S == 32)) {
    // global layers with the forward's log-sum-exp at hand: query side then key side on tcgen05 / TMEM (attention_bwd_tc.cu).  The query
    // side also produces the bias rows `rel` the key side reads and back-projects the bias cotangents through the tables itself, so the
    // two relpos_kernel launches of the warp-level path are gone.  GROVE_BWD_MMA_SYNC=1 keeps the warp-level kernels (A/B, cross-check).
    static const bool use_mma_sync = []() { const char* e = getenv("GROVE_BWD_MMA_SYNC"); return e && e[0] == '1'; }();
    if (lse_fwd != nullptr && !use_mma_sync) {
      rowdot_kernel<<<(unsigned)((M * heads + 255) / 256), 256, 0, st>>>(dO, O, Dsum, M * heads, HD);
      grove_count_launch();
      int rc = launch_attn_bwd_q_tc(qkv, dO, Rh, Rw, rel, lse_fwd, Dsum, dqkv, F, G, heads, HD, st);
      if (rc) return rc;
      float* aux = dqc + M * heads * HD;
      return launch_attn_bwd_kv_tc(qkv, dO, rel, lse_fwd, Dsum, aux, dqkv, F, G, heads, HD, st);
    }
  }
  const dim3 grid_rel((unsigned)(M / 64), heads);
  relpos_kernel<S, HD, WIN, 0><<<grid_rel, 256, smem_rel, st>>>(qkv, Rh, Rw, rel, nullptr, nullptr, nullptr, G, heads);
  rowdot_kernel<<<(unsigned)((M * heads + 255) / 256), 256, 0, st>>>(dO, O, Dsum, M * heads, HD);
  const dim3 grid(NT, regions * heads, F);
  bool fast_q = false;
  if constexpr (GFAST) {
    if (lse_fwd != nullptr) {      // the forward kernel's log-sum-exp: one key sweep instead of two
      attn_bwd_q_global_kernel<S, HD><<<dim3(NT, heads, F), 128, smem_qf, st>>>(qkv, dO, rel, Dsum, lse_fwd, dqc, A, heads);
      lse = const_cast<float*>(lse_fwd);
      fast_q = true;
    }
  }
  if (!fast_q) {
    // windows (and the small global grids): two key sweeps, or one when the forward kernel saved the row log-sum-exp
    attn_bwd_q_kernel<S, HD, WIN><<<grid, 128, smem_q, st>>>(qkv, qkv_bias, dO, rel, Dsum, lse_fwd, lse, dqc, A, G, heads);
    if (lse_fwd != nullptr) lse = const_cast<float*>(lse_fwd);
  }
  attn_bwd_kv_kernel<S, HD, WIN, GFAST><<<grid, 128, smem_kv, st>>>(qkv, qkv_bias, dO, rel, Dsum, lse, dqkv, G, heads);
  relpos_kernel<S, HD, WIN, 1><<<grid_rel, 256, smem_rel, st>>>(qkv, Rh, Rw, nullptr, A, dqc, dqkv, G, heads);
  grove_count_launch(5);
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

}  // namespace grove
using namespace grove;

extern "C" long long grove_attn_relpos_bwd_workspace_bytes(int F, int G, int heads, int hd, int ws) {
  const long long M = (long long)F * G * G, S = ws > 0 ? ws : G;
  return (M * heads * 2 * S * 2 + M * heads * 2 + M * heads * hd + M * heads * 4 /* aux of the tcgen05 key-side kernel */) * (long long)sizeof(float);
}

extern "C" int grove_attn_relpos_bwd_lse(const void* qkv, const void* qkv_bias_bf16, const void* rel_pos_h, const void* rel_pos_w, const void* att,
                                         const void* datt, void* dqkv, void* workspace, const float* lse_fwd, int F, int G, int heads, int hd, int ws,
                                         cudaStream_t stream) {
  GROVE_CHECK_ARG(qkv && rel_pos_h && rel_pos_w && att && datt && dqkv && workspace && F > 0 && G > 0 && heads > 0);
  GROVE_CHECK_ARG((G * G) % 64 == 0 && (ws == 0 || qkv_bias_bf16));
  GROVE_CHECK_ARG(((uintptr_t)workspace & 15) == 0);
  auto q = (const bf16_t*)qkv; auto qb = (const bf16_t*)qkv_bias_bf16; auto rh = (const bf16_t*)rel_pos_h; auto rw = (const bf16_t*)rel_pos_w;
  auto o = (const bf16_t*)att; auto d = (const bf16_t*)datt; auto out = (bf16_t*)dqkv; auto w = (float*)workspace;
#define RUN(S_, HD_, WIN_) return run_attn_bwd<S_, HD_, WIN_>(q, qb, rh, rw, o, d, out, w, lse_fwd, F, G, heads, stream)
  if (ws == 14) {
    if (hd == 64) RUN(14, 64, true);
    if (hd == 80) RUN(14, 80, true);
  } else if (ws == 0) {
    if (G == 64 && hd == 64) RUN(64, 64, false);
    if (G == 64 && hd == 80) RUN(64, 80, false);
    if (G == 32 && hd == 64) RUN(32, 64, false);
    if (G == 32 && hd == 80) RUN(32, 80, false);
    if (G == 16 && hd == 64) RUN(16, 64, false);
    if (G == 16 && hd == 80) RUN(16, 80, false);
  }
#undef RUN
  grove_set_error("attention backward is built for window 14 or global grids 16/32/64 with head dim 64/80 (got ws=%d G=%d hd=%d)", ws, G, hd);
  return GROVE_ERR_UNSUPPORTED;
}

extern "C" int grove_attn_relpos_bwd(const void* qkv, const void* qkv_bias_bf16, const void* rel_pos_h, const void* rel_pos_w, const void* att,
                                     const void* datt, void* dqkv, void* workspace, int F, int G, int heads, int hd, int ws, cudaStream_t stream) {
  return grove_attn_relpos_bwd_lse(qkv, qkv_bias_bf16, rel_pos_h, rel_pos_w, att, datt, dqkv, workspace, nullptr, F, G, heads, hd, ws, stream);
}
