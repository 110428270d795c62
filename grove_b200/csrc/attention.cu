// Fused attention with SAM's decomposed relative-position bias (image_encoder.py:301-326, 420-458).
//
// Round-1 implementation: flash-style online softmax on warp-level mma.sync (m16n8k16 bf16, fp32 accumulate),
// scores never touch HBM.  The rel-pos bias q.Rh[qh-kh+S-1] + q.Rw[qw-kw+S-1] (UNSCALED q, :313-315) is
// produced in-kernel by small extra MMAs against the rel-pos tables staged in shared memory.
//   * global blocks: per CTA 128 queries of one (frame, head); K/V streamed in 64-key blocks (cp.async double
//     buffer); rel_w lives in registers for the whole key loop, rel_h is recomputed every 8 key rows.
//   * window blocks: per CTA one (frame, window, head); operates on the UNPARTITIONED token-major qkv — pad
//     tokens (grid 64 -> 70) are synthesised in shared memory as k = b_k, v = b_v, pad queries are skipped.
#include "common.cuh"
#include "grove_b200_legacy.h"

namespace grove {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_smem16(uint32_t dst, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 64-element (128 B) bf16 rows, 16-byte chunks XOR-swizzled by the row: conflict-free ldmatrix
__device__ __forceinline__ uint32_t sw(uint32_t base, int row, int chunk) { return base + row * 128 + ((chunk ^ (row & 7)) << 4); }

constexpr float kLog2e = 1.4426950408889634f;

// =====================================================================================================
// Global attention.  qkv [F, N, 3, heads, 64] bf16; Rh/Rw [2G-1, 64] bf16; out [F, N, heads*64] bf16.
// =====================================================================================================
template <int G>
__global__ void __launch_bounds__(256, 1)
attn_global_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ Rh, const __nv_bfloat16* __restrict__ Rw,
                   __nv_bfloat16* __restrict__ out, int heads) {
  constexpr int N = G * G;
  constexpr int NB = N / 64;          // key blocks
  constexpr int NKW = G / 4;          // rel_w values a thread needs per row (its key columns mod G)
  constexpr int RPB = 64 / G;         // key grid-rows per 64-key block (1 for G=64, 2 for G=32)
  constexpr int TROWS = 2 * G - 1;    // rel-pos table rows
  constexpr int TPASS = (TROWS + 63) / 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t sQ = s0, sK = s0 + 16384, sV = sK + 16384, sRh = sV + 16384, sRw = sRh + 16384;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, f = blockIdx.z;
  const int D = heads * 64;
  const size_t tok_stride = (size_t)3 * D;
  const __nv_bfloat16* qbase = qkv + ((size_t)f * N) * tok_stride + h * 64;
  const __nv_bfloat16* kbase = qbase + D;
  const __nv_bfloat16* vbase = qbase + 2 * D;

  // ---- prologue: Q tile + rel-pos tables
  for (int i = tid; i < 128 * 8; i += 256) {
    const int r = i >> 3, c = i & 7;
    cp_async16(sw(sQ, r, c), qbase + (size_t)(q0 + r) * tok_stride + c * 8);
  }
  for (int i = tid; i < 128 * 8; i += 256) {
    const int r = i >> 3, c = i & 7;
    if (r < TROWS) {
      cp_async16(sw(sRh, r, c), Rh + r * 64 + c * 8);
      cp_async16(sw(sRw, r, c), Rw + r * 64 + c * 8);
    } else {
      st_smem16(sw(sRh, r, c), make_uint4(0, 0, 0, 0));
      st_smem16(sw(sRw, r, c), make_uint4(0, 0, 0, 0));
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // Q fragments stay in registers for the whole kernel
  uint32_t qf[4][4];
  const int r0 = warp * 16;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(sw(sQ, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), qf[ks]);

  const int qh = (q0 + r0) / G;            // all 16 rows of a warp share the grid row (16 | G)
  const int qw_lo = (q0 + r0) % G + g;     // row g ; row g+8 has qw_lo + 8

  // ---- rel_w[row][kw] for this thread's key columns, via scratch in the (still unused) K/V region
  float relw[2][NKW];
  {
    const uint32_t scr = sK + warp * 4096;  // [16][64] fp32, column XOR-swizzled by row
#pragma unroll
    for (int pass = 0; pass < TPASS; ++pass) {
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          uint32_t b[4];
          ldsm_x4(sw(sRw, pass * 64 + jp * 16 + (lane & 7) + (lane >> 4) * 8, 2 * ks + ((lane >> 3) & 1)), b);
          mma16816(acc[2 * jp], qf[ks], b[0], b[1]);
          mma16816(acc[2 * jp + 1], qf[ks], b[2], b[3]);
        }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = g + (e >> 1) * 8, col = 8 * j + 2 * t + (e & 1);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(scr + (row * 64 + (col ^ ((row & 7) << 3))) * 4), "f"(acc[j][e]) : "memory");
        }
      __syncwarp();
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        const int row = g + rs * 8, qw = qw_lo + rs * 8;
#pragma unroll
        for (int i = 0; i < NKW; ++i) {
          const int kw = 8 * (i >> 1) + 2 * t + (i & 1);
          const int idx = qw - kw + (G - 1) - pass * 64;
          if (idx >= 0 && idx < 64) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(scr + (row * 64 + (idx ^ ((row & 7) << 3))) * 4));
            relw[rs][i] = v * kLog2e;
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();  // scratch region becomes the K/V ring

  auto load_kv = [&](int b, int buf) {
    for (int i = tid; i < 64 * 8; i += 256) {
      const int r = i >> 3, c = i & 7;
      const size_t off = (size_t)(b * 64 + r) * tok_stride + c * 8;
      cp_async16(sw(sK + buf * 8192, r, c), kbase + off);
      cp_async16(sw(sV + buf * 8192, r, c), vbase + off);
    }
    cp_async_commit();
  };
  load_kv(0, 0);

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float hacc[4] = {0.f, 0.f, 0.f, 0.f};
  const float scale_log2 = 0.125f * kLog2e;  // hd^-0.5 with hd = 64

  for (int b = 0; b < NB; ++b) {
    const int buf = b & 1;
    if (b + 1 < NB) { load_kv(b + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();

    // rel_h for the key grid-rows of this block: recomputed for 8 consecutive kh at a time
    const int kh_first = b * RPB;
    if ((kh_first & 7) == 0) {
      hacc[0] = hacc[1] = hacc[2] = hacc[3] = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t bb[4];
        const int trow = qh + (G - 1) - kh_first - (lane & 7);
        ldsm_x4(sw(sRh, trow, (lane >> 3) + 4 * half), bb);
        mma16816(hacc, qf[2 * half], bb[0], bb[1]);
        mma16816(hacc, qf[2 * half + 1], bb[2], bb[3]);
      }
    }
    float relh[2][RPB];
#pragma unroll
    for (int rr = 0; rr < RPB; ++rr) {
      const int j = (kh_first + rr) & 7;
      const int src = (lane & ~3) | (j >> 1);
      const float lo = __shfl_sync(0xffffffffu, (j & 1) ? hacc[1] : hacc[0], src);
      const float hi = __shfl_sync(0xffffffffu, (j & 1) ? hacc[3] : hacc[2], src);
      relh[0][rr] = lo * kLog2e;
      relh[1][rr] = hi * kLog2e;
    }

    // S = Q K^T
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    const uint32_t kb = sK + buf * 8192, vb = sV + buf * 8192;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t bb[4];
        ldsm_x4(sw(kb, jp * 16 + (lane & 7) + (lane >> 4) * 8, 2 * ks + ((lane >> 3) & 1)), bb);
        mma16816(s[2 * jp], qf[ks], bb[0], bb[1]);
        mma16816(s[2 * jp + 1], qf[ks], bb[2], bb[3]);
      }
    // scale + bias, online softmax (log2 domain)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rs = e >> 1;
        const int col = 8 * j + 2 * t + (e & 1);       // key index within the block
        const int rr = (RPB == 1) ? 0 : (col / G);
        const int wi = ((j % (G / 8)) << 1) | (e & 1);  // index into relw (kw = col % G)
        const float v = s[j][e] * scale_log2 + relh[rs][rr] + relw[rs][wi];
        s[j][e] = v;
        mx[rs] = fmaxf(mx[rs], v);
      }
    float alpha[2];
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      mx[rs] = fmaxf(mx[rs], __shfl_xor_sync(0xffffffffu, mx[rs], 1));
      mx[rs] = fmaxf(mx[rs], __shfl_xor_sync(0xffffffffu, mx[rs], 2));
      const float mn = fmaxf(m_run[rs], mx[rs]);
      alpha[rs] = exp2f(m_run[rs] - mn);
      m_run[rs] = mn;
    }
    float rsum[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[j][e] - m_run[e >> 1]);
        s[j][e] = p;
        rsum[e >> 1] += p;
      }
    l_run[0] = l_run[0] * alpha[0] + rsum[0];
    l_run[1] = l_run[1] * alpha[1] + rsum[1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
      o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bb[4];
        ldsm_x4_t(sw(vb, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dp + (lane >> 4)), bb);
        mma16816(o[2 * dp], pa, bb[0], bb[1]);
        mma16816(o[2 * dp + 1], pa, bb[2], bb[3]);
      }
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    l_run[rs] += __shfl_xor_sync(0xffffffffu, l_run[rs], 1);
    l_run[rs] += __shfl_xor_sync(0xffffffffu, l_run[rs], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  __nv_bfloat16* orow0 = out + ((size_t)f * N + q0 + r0 + g) * D + h * 64 + 2 * t;
  __nv_bfloat16* orow1 = orow0 + (size_t)8 * D;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(orow0 + 8 * j) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
    *reinterpret_cast<uint32_t*>(orow1 + 8 * j) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
  }
}

// =====================================================================================================
// Windowed attention (ws = 14).  qkv [F, G, G, 3, heads, 64] bf16 (unpartitioned); out [F, G, G, heads*64].
// =====================================================================================================
constexpr int WS = 14, WQ = 196, WQP = 208;  // window side, tokens, tokens padded to 13 m-tiles / 26 n-tiles

// row addressing of a [rows][HD] bf16 tile: HD = 64 -> 128-byte rows with the XOR swizzle; HD = 80 (ViT-H) -> plain 160-byte
// rows (8 consecutive rows land on banks 0,8,16,24,0,.. : a 2-way ldmatrix conflict, accepted)
template <int HD>
__device__ __forceinline__ uint32_t swh(uint32_t base, int row, int chunk) {
  if (HD == 64) return base + row * 128 + ((chunk ^ (row & 7)) << 4);
  return base + row * (HD * 2) + (chunk << 4);
}

template <int NT, int HD>  // n-tiles (8 keys each) in this key block; head dim
__device__ __forceinline__ void window_block(const uint32_t (&qf)[HD / 16][4], uint32_t sK, uint32_t sV, int key0, const float* relh_s,
                                             const float* relw_s, int qrow_lo, float (&o)[HD / 8][4], float (&m_run)[2], float (&l_run)[2],
                                             int lane) {
  const int g = lane >> 2, t = lane & 3;
  float s[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
      uint32_t bb[4];
      ldsm_x4(swh<HD>(sK, key0 + jp * 16 + (lane & 7) + (lane >> 4) * 8, 2 * ks + ((lane >> 3) & 1)), bb);
      mma16816(s[2 * jp], qf[ks], bb[0], bb[1]);
      mma16816(s[2 * jp + 1], qf[ks], bb[2], bb[3]);
    }
  const float scale_log2 = (HD == 64 ? 0.125f : 0.11180339887498949f) * kLog2e;   // hd^-0.5
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int rs = e >> 1;
      const int key = key0 + 8 * j + 2 * t + (e & 1);
      const int q = min(qrow_lo + rs * 8, WQ - 1);  // rows >= 196 are never stored; clamp keeps the reads in range
      float v = -INFINITY;
      if (key < WQ) v = s[j][e] * scale_log2 + relh_s[q * WS + key / WS] + relw_s[q * WS + key % WS];
      s[j][e] = v;
      mx[rs] = fmaxf(mx[rs], v);
    }
  float alpha[2];
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    mx[rs] = fmaxf(mx[rs], __shfl_xor_sync(0xffffffffu, mx[rs], 1));
    mx[rs] = fmaxf(mx[rs], __shfl_xor_sync(0xffffffffu, mx[rs], 2));
    const float mn = fmaxf(m_run[rs], mx[rs]);
    alpha[rs] = exp2f(m_run[rs] - mn);
    m_run[rs] = mn;
  }
  float rsum[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float p = exp2f(s[j][e] - m_run[e >> 1]);
      s[j][e] = p;
      rsum[e >> 1] += p;
    }
  l_run[0] = l_run[0] * alpha[0] + rsum[0];
  l_run[1] = l_run[1] * alpha[1] + rsum[1];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
    o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
  }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < HD / 16; ++dp) {
      uint32_t bb[4];
      ldsm_x4_t(swh<HD>(sV, key0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dp + (lane >> 4)), bb);
      mma16816(o[2 * dp], pa, bb[0], bb[1]);
      mma16816(o[2 * dp + 1], pa, bb[2], bb[3]);
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(128, HD == 64 ? 2 : 1)
attn_window_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ qkv_bias, const __nv_bfloat16* __restrict__ Rh,
                   const __nv_bfloat16* __restrict__ Rw, __nv_bfloat16* __restrict__ out, int G, int heads) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 127u) & ~127u;
  constexpr int RB = HD * 2, CH = HD / 8;   // bytes / 16-byte chunks per row
  const uint32_t sQ = s0, sK = sQ + WQP * RB, sV = sK + WQP * RB, sRh = sV + WQP * RB, sRw = sRh + 32 * RB;
  const uint32_t sBias = sRw + 32 * RB;  // relh [196][14] fp32 then relw [196][14] fp32
  float* relh_s = reinterpret_cast<float*>(smem_raw + (sBias - smem_u32(smem_raw)));
  float* relw_s = relh_s + WQ * WS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int nW = (G + WS - 1) / WS;
  const int wy = blockIdx.x / nW, wx = blockIdx.x % nW, h = blockIdx.y, f = blockIdx.z;
  const int D = heads * HD;

  // ---- stage Q, K, V of the window (pad tokens: k = b_k, v = b_v; rows >= 196: zeros) and the rel-pos tables
  for (int i = tid; i < 3 * WQP * CH; i += 128) {
    const int which = i / (WQP * CH), rem = i % (WQP * CH), row = rem / CH, c = rem % CH;
    const uint32_t dst = swh<HD>(sQ + which * (WQP * RB), row, c);
    const int gy = wy * WS + row / WS, gx = wx * WS + row % WS;
    if (row < WQ && gy < G && gx < G) {
      cp_async16(dst, qkv + (((size_t)(f * G + gy) * G + gx) * 3 + which) * D + h * HD + c * 8);
    } else if (row < WQ && which > 0) {
      st_smem16(dst, __ldg(reinterpret_cast<const uint4*>(qkv_bias + which * D + h * HD + c * 8)));
    } else {
      st_smem16(dst, make_uint4(0, 0, 0, 0));
    }
  }
  for (int i = tid; i < 2 * 32 * CH; i += 128) {
    const int which = i / (32 * CH), rem = i % (32 * CH), row = rem / CH, c = rem % CH;
    const uint32_t dst = swh<HD>(which ? sRw : sRh, row, c);
    if (row < 2 * WS - 1) cp_async16(dst, (which ? Rw : Rh) + row * HD + c * 8);
    else st_smem16(dst, make_uint4(0, 0, 0, 0));
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  for (int mt = warp; mt < WQP / 16; mt += 4) {
    const int r0 = mt * 16;
    uint32_t qf[HD / 16][4];
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) ldsm_x4(swh<HD>(sQ, r0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), qf[ks]);

    // rel-pos products against the whole (27-row) tables, scattered to relh_s[q][kh], relw_s[q][kw] (pre-scaled by log2 e)
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
          uint32_t bb[4];
          ldsm_x4(swh<HD>(which ? sRw : sRh, jp * 16 + (lane & 7) + (lane >> 4) * 8, 2 * ks + ((lane >> 3) & 1)), bb);
          mma16816(acc[2 * jp], qf[ks], bb[0], bb[1]);
          mma16816(acc[2 * jp + 1], qf[ks], bb[2], bb[3]);
        }
      float* dst = which ? relw_s : relh_s;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = r0 + g + (e >> 1) * 8, r = 8 * j + 2 * t + (e & 1);
          const int qc = which ? (q % WS) : (q / WS);
          const int kc = qc + (WS - 1) - r;
          if (q < WQ && kc >= 0 && kc < WS) dst[q * WS + kc] = acc[j][e] * kLog2e;
        }
    }
    __syncwarp();

    float o[HD / 8][4];
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    window_block<14, HD>(qf, sK, sV, 0, relh_s, relw_s, r0 + g, o, m_run, l_run, lane);
    window_block<12, HD>(qf, sK, sV, 112, relh_s, relw_s, r0 + g, o, m_run, l_run, lane);

#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      l_run[rs] += __shfl_xor_sync(0xffffffffu, l_run[rs], 1);
      l_run[rs] += __shfl_xor_sync(0xffffffffu, l_run[rs], 2);
    }
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      const int q = r0 + g + rs * 8;
      const int gy = wy * WS + q / WS, gx = wx * WS + q % WS;
      if (q < WQ && gy < G && gx < G) {
        const float inv = 1.f / l_run[rs];
        __nv_bfloat16* orow = out + ((size_t)(f * G + gy) * G + gx) * D + h * HD + 2 * t;
#pragma unroll
        for (int j = 0; j < HD / 8; ++j) *reinterpret_cast<uint32_t*>(orow + 8 * j) = pack_bf16(o[j][2 * rs] * inv, o[j][2 * rs + 1] * inv);
      }
    }
  }
}

constexpr int kGlobalSmem = 5 * 16384 + 128;
template <int HD>
constexpr int window_smem() { return 3 * WQP * HD * 2 + 2 * 32 * HD * 2 + 2 * WQ * WS * 4 + 128; }

}  // namespace grove
using namespace grove;

extern "C" int grove_attn_global_relpos_fwd_mma(const void* qkv, const void* rel_pos_h, const void* rel_pos_w, void* out, int F, int G,
                                            int heads, int hd, cudaStream_t stream) {
  GROVE_CHECK_ARG(qkv && rel_pos_h && rel_pos_w && out && F > 0 && heads > 0);
  if (hd != 64 || (G != 64 && G != 32)) {
    grove_set_error("grove_attn_global_relpos_fwd_mma: only hd=64 and G in {32,64} are built (got hd=%d G=%d)", hd, G);
    return GROVE_ERR_UNSUPPORTED;
  }
  GROVE_CHECK_ARG(F <= 65535 && heads <= 65535);
  dim3 grid(G * G / 128, heads, F);
  cudaError_t e;
  if (G == 64) {
    e = cudaFuncSetAttribute(attn_global_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGlobalSmem);
    if (e == cudaSuccess)
      attn_global_kernel<64><<<grid, 256, kGlobalSmem, stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)rel_pos_h,
                                                                 (const __nv_bfloat16*)rel_pos_w, (__nv_bfloat16*)out, heads);
  } else {
    e = cudaFuncSetAttribute(attn_global_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGlobalSmem);
    if (e == cudaSuccess)
      attn_global_kernel<32><<<grid, 256, kGlobalSmem, stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)rel_pos_h,
                                                                 (const __nv_bfloat16*)rel_pos_w, (__nv_bfloat16*)out, heads);
  }
  if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

template <int HD>
static int launch_window(const void* qkv, const void* qkv_bias_bf16, const void* rel_pos_h, const void* rel_pos_w, void* out, int F, int G,
                         int heads, cudaStream_t stream) {
  const int nW = (G + WS - 1) / WS;
  constexpr int smem = window_smem<HD>();
  cudaError_t e = cudaFuncSetAttribute(attn_window_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
  attn_window_kernel<HD><<<dim3(nW * nW, heads, F), 128, smem, stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)qkv_bias_bf16,
                                                                         (const __nv_bfloat16*)rel_pos_h, (const __nv_bfloat16*)rel_pos_w,
                                                                         (__nv_bfloat16*)out, G, heads);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_attn_window_relpos_fwd(const void* qkv, const void* qkv_bias_bf16, const void* rel_pos_h, const void* rel_pos_w,
                                            void* out, int F, int G, int heads, int hd, int ws, cudaStream_t stream) {
  GROVE_CHECK_ARG(qkv && qkv_bias_bf16 && rel_pos_h && rel_pos_w && out && F > 0 && G > 0 && heads > 0);
  if ((hd != 64 && hd != 80) || ws != 14) {
    grove_set_error("grove_attn_window_relpos_fwd: head dim 64 / 80 and window 14 are built (got hd=%d ws=%d)", hd, ws);
    return GROVE_ERR_UNSUPPORTED;
  }
  GROVE_CHECK_ARG(F <= 65535 && heads <= 65535);
  return hd == 64 ? launch_window<64>(qkv, qkv_bias_bf16, rel_pos_h, rel_pos_w, out, F, G, heads, stream)
                  : launch_window<80>(qkv, qkv_bias_bf16, rel_pos_h, rel_pos_w, out, F, G, heads, stream);
}
