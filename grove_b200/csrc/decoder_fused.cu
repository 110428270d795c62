// Token side of a TwoWayAttentionBlock (transformer.py:151-182) in two fused kernels around the token->image attention.
//
// The 6 tokens x 256 channels of an instance are tiny; what the unfused path (decoder_ops.cu: ~25 launches per layer, each 5-50 us of
// pure latency) paid for was launch / dependency latency and, for the MLP, re-reading 4 MB of fp32 weights once per 8-row group.
//   A  twoway_tokens_a_kernel   one CTA per instance: q/k/v projections -> 8-head self-attention over the 6 tokens -> out_proj (+ residual
//                               unless skip_first_layer_pe) -> norm1 -> q_proj of the token->image attention.
//   B  twoway_tokens_b_kernel   one CLUSTER of 8 CTAs per instance: token->image out_proj + residual -> norm2 -> MLP with the 2048 hidden
//                               units split over the cluster (each CTA streams 1/8 of lin1 / lin2, partial sums reduced through
//                               distributed shared memory) -> norm3 (row statistics exchanged through DSMEM) -> k/v projections of the
//                               image->token attention (and, after the last layer, the q projection of the final attention), 1/8 each.
// Weights are fp32 and TRANSPOSED ([K][N]) so that a warp reads 128 contiguous bytes per k; every thread owns one output column and keeps
// six row accumulators, the activations are broadcast from shared memory.  All arithmetic fp32, two-pass LayerNorm like nn.LayerNorm.
#include <cooperative_groups.h>

#include "common.cuh"
#include "grove_b200.h"

namespace cg = cooperative_groups;

namespace grove {

constexpr int kT = 6;        // tokens per instance: iou (1) + mask (4) + prompt (1), mask_decoder.py:165-172
constexpr int kC = 256;      // transformer_dim
constexpr int kCI = 128;     // cross-attention internal width (attention_downsample_rate 2)
constexpr int kHeads = 8;
constexpr int kClu = 8;      // CTAs per instance in kernel B
constexpr int kRingA = 3;    // cp.async ring depth of part A (one CTA per SM: 96 KB)
constexpr int kRingB = 2;    // ... of part B (two CTAs of a cluster may share an SM: 64 KB each)
constexpr int kRingTile = 32 * 256 * 4;

// acc[r] += sum_k act[r][k] * wT[k*ldw + col0 + tid] for r < 6, k < K, for the NC columns [col0, col0 + NC) (thread tid < NC owns one).
// The weights stream through shared memory in [32 k][NC] tiles with cp.async (NS-deep ring, whole CTA cooperates): with one CTA per SM
// the stream is pure latency, and register-staged loads (which ptxas interleaves with the FMAs whatever the source order) kept ~8 loads
// per thread in flight -- a tenth of the L2 bandwidth an SM can pull (119 us for part A, measured).  `act` lives in shared memory
// (row stride `lda` floats, broadcast float4 reads).  Ends with a __syncthreads(): the ring may be reused immediately.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int K, int NC, int NS>
__device__ __forceinline__ void cta_cols_dot(const float* __restrict__ wT, int ldw, int col0, const float* act, int lda, float* ring,
                                             float (&acc)[kT], int tid) {
  constexpr int KT = 32, C16 = NC / 4, CH = KT * C16, NTILES = K / KT;
  static_assert(K % KT == 0 && NC % 4 == 0 && NC <= 256, "");
  auto issue = [&](int kt) {
    if (kt < NTILES) {
      float* dst = ring + (kt % NS) * (KT * NC);
      for (int idx = tid; idx < CH; idx += 256) {
        const int row = idx / C16, c = idx % C16;
        cp_async16(dst + row * NC + c * 4, wT + (size_t)(kt * KT + row) * ldw + col0 + c * 4);
      }
    }
    cp_async_commit();               // always a group, so that the wait depth below is uniform
  };
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) issue(s);
#pragma unroll 1
  for (int kt = 0; kt < NTILES; ++kt) {
    issue(kt + NS - 1);
    cp_async_wait<NS - 1>();         // tile kt has landed (for this thread's copies; the barrier makes it true for everyone's)
    __syncthreads();
    if (tid < NC) {
      const float* wb = ring + (kt % NS) * (KT * NC) + tid;
#pragma unroll
      for (int i = 0; i < KT; i += 4) {
        const float w0 = wb[(i + 0) * NC], w1 = wb[(i + 1) * NC], w2 = wb[(i + 2) * NC], w3 = wb[(i + 3) * NC];
#pragma unroll
        for (int r = 0; r < kT; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(act + r * lda + kt * KT + i);
          acc[r] = fmaf(a.x, w0, fmaf(a.y, w1, fmaf(a.z, w2, fmaf(a.w, w3, acc[r]))));
        }
      }
    }
    __syncthreads();                 // the slot is refilled NS-1 iterations later
  }
}

// LayerNorm of row `r` (256 values in shared memory) by one warp: two-pass mean / centred variance; returns this lane's 8 outputs
__device__ __forceinline__ void ln_row_256(const float* row, const float* __restrict__ g, const float* __restrict__ be, float eps, int lane, float (&y)[8]) {
  float v[8], s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = row[i * 32 + lane]; s += v[i]; }
  const float mean = warp_sum(s) * (1.f / kC);
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] -= mean; qv += v[i] * v[i]; }
  const float rstd = rsqrtf(warp_sum(qv) * (1.f / kC) + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = v[i] * rstd * g[i * 32 + lane] + be[i * 32 + lane];
}

__global__ void __launch_bounds__(256) twoway_tokens_a_kernel(const float* __restrict__ queries, const float* __restrict__ tokens,
                                                              const grove_twoway_a_params p, float* __restrict__ q_out, float* __restrict__ qt_out) {
  __shared__ __align__(16) float xs[kT][kC];     // queries
  __shared__ __align__(16) float qin[kT][kC];    // queries + query_pe (= the original tokens); later norm1 output + query_pe
  __shared__ __align__(16) float Qs[kT][kC];     // later: pre-LayerNorm rows
  __shared__ __align__(16) float Ks[kT][kC];
  __shared__ __align__(16) float Vs[kT][kC];
  __shared__ __align__(16) float att[kT][kC];
  extern __shared__ __align__(16) float ring[];   // cp.async weight ring: kRingA stages of [32][256] fp32
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xq = queries + (size_t)b * kT * kC;
  const float* xt = tokens + (size_t)b * kT * kC;
  for (int i = tid; i < kT * kC; i += 256) {
    const float x = xq[i];
    (&xs[0][0])[i] = x;
    (&qin[0][0])[i] = p.skip_pe ? x : x + xt[i];
  }
  __syncthreads();
  {  // q, k from (queries + pe), v from queries (transformer.py:155-160)
    const int n = tid;
    float aq[kT], ak[kT], av[kT];
    const float bq = p.bq[n], bk = p.bk[n], bv = p.bv[n];
#pragma unroll
    for (int r = 0; r < kT; ++r) { aq[r] = bq; ak[r] = bk; av[r] = bv; }
    cta_cols_dot<kC, kC, kRingA>(p.wq_t, kC, 0, &qin[0][0], kC, ring, aq, tid);
    cta_cols_dot<kC, kC, kRingA>(p.wk_t, kC, 0, &qin[0][0], kC, ring, ak, tid);
    cta_cols_dot<kC, kC, kRingA>(p.wv_t, kC, 0, &xs[0][0], kC, ring, av, tid);
#pragma unroll
    for (int r = 0; r < kT; ++r) { Qs[r][n] = aq[r]; Ks[r][n] = ak[r]; Vs[r][n] = av[r]; }
  }
  __syncthreads();
  if (tid < kHeads * kT) {   // self-attention among the 6 tokens: thread = (head, query token), 32-dim heads
    const int h = tid / kT, tq = tid % kT;
    constexpr int dh = kC / kHeads;
    float s[kT], mx = -INFINITY;
#pragma unroll
    for (int tk = 0; tk < kT; ++tk) {
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < dh; ++d) a += Qs[tq][h * dh + d] * Ks[tk][h * dh + d];
      s[tk] = a * rsqrtf((float)dh);
      mx = fmaxf(mx, s[tk]);
    }
    float l = 0.f;
#pragma unroll
    for (int tk = 0; tk < kT; ++tk) { s[tk] = expf(s[tk] - mx); l += s[tk]; }
    const float inv = 1.f / l;
    for (int d = 0; d < dh; ++d) {
      float a = 0.f;
#pragma unroll
      for (int tk = 0; tk < kT; ++tk) a += s[tk] * Vs[tk][h * dh + d];
      att[tq][h * dh + d] = a * inv;
    }
  }
  __syncthreads();
  {  // out_proj; layer 0 (skip_first_layer_pe) REPLACES the queries, the others add the residual (transformer.py:155-161)
    const int n = tid;
    float ao[kT];
    const float bo = p.bo[n];
#pragma unroll
    for (int r = 0; r < kT; ++r) ao[r] = bo;
    cta_cols_dot<kC, kC, kRingA>(p.wo_t, kC, 0, &att[0][0], kC, ring, ao, tid);
#pragma unroll
    for (int r = 0; r < kT; ++r) Qs[r][n] = ao[r] + (p.skip_pe ? 0.f : xs[r][n]);
  }
  __syncthreads();
  if (warp < kT) {           // norm1; the token->image query input is norm1(...) + query_pe (always, also on layer 0: transformer.py:164)
    float y[8];
    ln_row_256(&Qs[warp][0], p.ln_g, p.ln_b, p.ln_eps, lane, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = i * 32 + lane;
      q_out[((size_t)b * kT + warp) * kC + c] = y[i];
      qin[warp][c] = y[i] + xt[warp * kC + c];
    }
  }
  __syncthreads();
  {
    const int n = tid;
    float a[kT];
    const float bq2 = n < kCI ? p.bq2[n] : 0.f;
#pragma unroll
    for (int r = 0; r < kT; ++r) a[r] = bq2;
    cta_cols_dot<kC, kCI, kRingA>(p.wq2_t, kCI, 0, &qin[0][0], kC, ring, a, tid);
    if (n < kCI) {
#pragma unroll
      for (int r = 0; r < kT; ++r) qt_out[((size_t)b * kT + r) * kCI + n] = a[r];
    }
  }
}

__global__ void __cluster_dims__(kClu, 1, 1) __launch_bounds__(256)
twoway_tokens_b_kernel(const float* __restrict__ queries, const float* __restrict__ att_in, const float* __restrict__ tokens,
                       const grove_twoway_b_params p, float* __restrict__ q_out, float* __restrict__ kt_out, float* __restrict__ vt_out,
                       float* __restrict__ qf_out) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ __align__(16) float a2[kT][kCI];       // token->image attention output (before out_proj)
  __shared__ __align__(16) float z[kT][kC];         // norm2 output
  __shared__ __align__(16) float hs[kT][256];       // this CTA's 256 hidden units
  __shared__ __align__(16) float mp[kT][kC];        // pre-norm2 rows, then this CTA's partial lin2 output (read by the whole cluster)
  __shared__ __align__(16) float wf[kT][kC];        // norm3 output (gathered from the cluster)
  __shared__ __align__(16) float wi[kT][kC];        // norm3 output + query_pe
  __shared__ float wsl[kT][32];                     // this CTA's 32-column slice of the norm3 output (read by the whole cluster)
  __shared__ float ssum[kT], ssq[kT];               // this CTA's slice statistics (read by the whole cluster)
  __shared__ float part[16][16][kT];
  extern __shared__ __align__(16) float ring[];     // cp.async weight ring: kRingB stages of [32][256] fp32
  const int rk = (int)cluster.block_rank();
  const int b = blockIdx.x / kClu, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xq = queries + (size_t)b * kT * kC;
  const float* xt = tokens + (size_t)b * kT * kC;
  for (int i = tid; i < kT * kCI; i += 256) (&a2[0][0])[i] = att_in[(size_t)b * kT * kCI + i];
  __syncthreads();
  {  // token->image out_proj + residual (transformer.py:164-168); computed redundantly by the 8 CTAs (128 KB of weights each)
    const int n = tid;
    float ao[kT];
    const float bo = p.bo[n];
#pragma unroll
    for (int r = 0; r < kT; ++r) ao[r] = bo;
    cta_cols_dot<kCI, kC, kRingB>(p.wo_t, kC, 0, &a2[0][0], kCI, ring, ao, tid);
#pragma unroll
    for (int r = 0; r < kT; ++r) mp[r][n] = ao[r] + xq[r * kC + n];
  }
  __syncthreads();
  if (warp < kT) {           // norm2
    float y[8];
    ln_row_256(&mp[warp][0], p.ln2_g, p.ln2_b, p.ln2_eps, lane, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[warp][i * 32 + lane] = y[i];
  }
  __syncthreads();
  {  // MLP lin1 + ReLU: hidden units [rk*256, rk*256+256)
    const int j = rk * 256 + tid;
    float ah[kT];
    const float b1 = p.b1[j];
#pragma unroll
    for (int r = 0; r < kT; ++r) ah[r] = b1;
    cta_cols_dot<kC, 256, kRingB>(p.w1_t, p.mlp_dim, rk * 256, &z[0][0], kC, ring, ah, tid);
#pragma unroll
    for (int r = 0; r < kT; ++r) hs[r][tid] = fmaxf(ah[r], 0.f);
  }
  __syncthreads();
  {  // MLP lin2, partial over this CTA's hidden units
    const int n = tid;
    float am[kT];
#pragma unroll
    for (int r = 0; r < kT; ++r) am[r] = 0.f;
    cta_cols_dot<256, kC, kRingB>(p.w2_t + (size_t)rk * 256 * kC, kC, 0, &hs[0][0], 256, ring, am, tid);
#pragma unroll
    for (int r = 0; r < kT; ++r) mp[r][n] = am[r];
  }
  cluster.sync();
  // reduce: this CTA owns output columns [rk*32, rk*32+32) of all 6 rows; warp r handles row r
  float v = 0.f, d = 0.f;
  const int col = rk * 32 + lane;
  if (warp < kT) {
#pragma unroll
    for (int c = 0; c < kClu; ++c) v += cluster.map_shared_rank(&mp[0][0], c)[warp * kC + col];
    v += p.b2[col] + z[warp][col];                    // + bias + residual (transformer.py:171-173)
    const float s = warp_sum(v);
    if (lane == 0) ssum[warp] = s;
  }
  cluster.sync();
  if (warp < kT) {           // norm3, two-pass across the cluster: mean first, then the centred sum of squares
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kClu; ++c) s += cluster.map_shared_rank(&ssum[0], c)[warp];
    d = v - s * (1.f / kC);
    const float q2 = warp_sum(d * d);
    if (lane == 0) ssq[warp] = q2;
  }
  cluster.sync();
  if (warp < kT) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kClu; ++c) s += cluster.map_shared_rank(&ssq[0], c)[warp];
    const float w = d * rsqrtf(s * (1.f / kC) + p.ln3_eps) * p.ln3_g[col] + p.ln3_b[col];
    wsl[warp][lane] = w;
    if (q_out) q_out[((size_t)b * kT + warp) * kC + col] = w;
  }
  cluster.sync();
  for (int i = tid; i < kT * kC; i += 256) {         // gather the norm3 rows from the 8 slices
    const int r = i / kC, c = i % kC;
    const float w = cluster.map_shared_rank(&wsl[0][0], c >> 5)[r * 32 + (c & 31)];
    wf[r][c] = w;
    wi[r][c] = w + xt[i];
  }
  __syncthreads();
  // k / v projections of the image->token attention (keys from queries + pe, values from queries; transformer.py:175-177) and, when
  // asked, the q projection of the NEXT token->image attention: this CTA produces outputs [rk*16, rk*16+16) of each, thread = (k-group, n)
  {
    const int g = tid >> 4, nl = tid & 15, n = rk * 16 + nl;
    const int nmat = qf_out ? 3 : 2;
    for (int mat = 0; mat < nmat; ++mat) {
      const float* wT = mat == 0 ? p.wk_t : (mat == 1 ? p.wv_t : p.wqf_t);
      const float* src = mat == 1 ? &wf[0][0] : &wi[0][0];
      float acc[kT];
#pragma unroll
      for (int r = 0; r < kT; ++r) acc[r] = 0.f;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int k = g + 16 * i;
        const float w = __ldg(wT + (size_t)k * kCI + n);
#pragma unroll
        for (int r = 0; r < kT; ++r) acc[r] = fmaf(src[r * kC + k], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kT; ++r) part[g][nl][r] = acc[r];
      __syncthreads();
      if (tid < 16 * kT) {
        const int nl2 = tid / kT, r = tid % kT, n2 = rk * 16 + nl2;
        float s = (mat == 0 ? p.bk : (mat == 1 ? p.bv : p.bqf))[n2];
#pragma unroll
        for (int gg = 0; gg < 16; ++gg) s += part[gg][nl2][r];
        float* o = mat == 0 ? kt_out : (mat == 1 ? vt_out : qf_out);
        o[((size_t)b * kT + r) * kCI + n2] = s;
      }
      __syncthreads();
    }
  }
  cluster.sync();            // no CTA may exit while a peer can still read its shared memory
}

// ---------------------------------------------------------------- token -> image attention, 256 threads per (instance, head)
// q fp32 [B,T,H*DH] (projected); k,v bf16 [*,N,H*DH] with row block src_of[b].  Each thread walks keys n = tid, tid+256, ... with an
// online-softmax state per token; states are merged warp-first (shuffles), then across the 8 warps through shared memory.
template <int T, int DH, int NT>
__global__ void __launch_bounds__(NT) t2i_attention_wide_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                                const __nv_bfloat16* __restrict__ v, const int* __restrict__ src_of,
                                                                float* __restrict__ out, float* __restrict__ lse_out, int N, int heads) {
  static_assert(DH == 16, "one key/value head row = two 16-byte loads");
  constexpr int NW = NT / 32;
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HD = heads * DH;
  const float scale_log2 = rsqrtf((float)DH) * 1.4426950408889634f;
  __shared__ float qs[T][DH];
  __shared__ float red_m[T][NW], red_l[T][NW];
  __shared__ float red_acc[T][DH][NW];
  if (tid < T * DH) qs[tid / DH][tid % DH] = q[((size_t)b * T + tid / DH) * HD + h * DH + tid % DH] * scale_log2;
  __syncthreads();
  const size_t base = (size_t)(src_of ? src_of[b] : b) * N * HD + h * DH;
  float m[T], l[T], acc[T][DH];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    m[t] = -INFINITY; l[t] = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) acc[t][d] = 0.f;
  }
  for (int n = tid; n < N; n += NT) {
    const uint4* kp = reinterpret_cast<const uint4*>(k + base + (size_t)n * HD);
    const uint4* vp = reinterpret_cast<const uint4*>(v + base + (size_t)n * HD);
    const uint4 k0 = __ldg(kp), k1 = __ldg(kp + 1), v0 = __ldg(vp), v1 = __ldg(vp + 1);
    float kf[DH], vf[DH];
    const uint32_t ku[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    const uint32_t vu[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = unpack_bf16(ku[i]), c = unpack_bf16(vu[i]);
      kf[2 * i] = a.x; kf[2 * i + 1] = a.y; vf[2 * i] = c.x; vf[2 * i + 1] = c.y;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s += qs[t][d] * kf[d];
      const float mn = fmaxf(m[t], s);
      const float a = exp2f(m[t] - mn), p = exp2f(s - mn);
      m[t] = mn;
      l[t] = l[t] * a + p;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc[t][d] = acc[t][d] * a + p * vf[d];
    }
  }
  // merge: warp maximum -> cross-warp maximum -> rescaled sums
  float gm[T], w[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float wm = warp_max(m[t]);
    if (lane == 0) red_m[t][warp] = wm;
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float x = -INFINITY;
#pragma unroll
    for (int i = 0; i < NW; ++i) x = fmaxf(x, red_m[t][i]);
    gm[t] = x;
    w[t] = (m[t] == -INFINITY) ? 0.f : exp2f(m[t] - x);
    const float ls = warp_sum(l[t] * w[t]);
    if (lane == 0) red_l[t][warp] = ls;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      const float s = warp_sum(acc[t][d] * w[t]);
      if (lane == 0) red_acc[t][d][warp] = s;
    }
  }
  __syncthreads();
  if (tid < T * DH) {
    const int t = tid / DH, d = tid % DH;
    float lsum = 0.f, a = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) { lsum += red_l[t][i]; a += red_acc[t][d][i]; }
    out[((size_t)b * T + t) * HD + h * DH + d] = a / lsum;
    if (lse_out && d == 0) {
      float x = -INFINITY;
#pragma unroll
      for (int i = 0; i < NW; ++i) x = fmaxf(x, red_m[t][i]);
      lse_out[((size_t)b * T + t) * heads + h] = x + log2f(lsum);   // log2 domain, scale folded in (backward pass)
    }
  }
}

}  // namespace grove
using namespace grove;

extern "C" int grove_twoway_block_tokens_a_fwd(const float* queries, const float* tokens, const grove_twoway_a_params* p, float* queries_out,
                                               float* qt_out, int B, int T, int C, cudaStream_t stream) {
  GROVE_CHECK_ARG(queries && tokens && p && queries_out && qt_out && B > 0);
  GROVE_CHECK_ARG(p->wq_t && p->bq && p->wk_t && p->bk && p->wv_t && p->bv && p->wo_t && p->bo && p->ln_g && p->ln_b && p->wq2_t && p->bq2);
  if (T != kT || C != kC) { grove_set_error("the fused two-way token kernels are built for 6 tokens x 256 channels (got %d x %d)", T, C); return GROVE_ERR_UNSUPPORTED; }
  static GrovePerDeviceOnce attr;
  if (attr.first_time()) cudaFuncSetAttribute(twoway_tokens_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingA * kRingTile);
  twoway_tokens_a_kernel<<<B, 256, kRingA * kRingTile, stream>>>(queries, tokens, *p, queries_out, qt_out);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_twoway_block_tokens_b_fwd(const float* queries, const float* att, const float* tokens, const grove_twoway_b_params* p,
                                               float* queries_out, float* kt_out, float* vt_out, float* qf_out, int B, int T, int C,
                                               cudaStream_t stream) {
  GROVE_CHECK_ARG(queries && att && tokens && p && queries_out && kt_out && vt_out && B > 0);
  GROVE_CHECK_ARG(p->wo_t && p->bo && p->ln2_g && p->ln2_b && p->w1_t && p->b1 && p->w2_t && p->b2 && p->ln3_g && p->ln3_b && p->wk_t && p->bk &&
                  p->wv_t && p->bv);
  GROVE_CHECK_ARG(qf_out == nullptr || (p->wqf_t && p->bqf));
  if (T != kT || C != kC || p->mlp_dim != kClu * 256) {
    grove_set_error("the fused two-way token kernels are built for 6 tokens x 256 channels, MLP width 2048 (got %d x %d, %d)", T, C, p->mlp_dim);
    return GROVE_ERR_UNSUPPORTED;
  }
  static GrovePerDeviceOnce attr;
  if (attr.first_time()) cudaFuncSetAttribute(twoway_tokens_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingB * kRingTile);
  twoway_tokens_b_kernel<<<B * kClu, 256, kRingB * kRingTile, stream>>>(queries, att, tokens, *p, queries_out, kt_out, vt_out, qf_out);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

// wide variant of grove_decoder_t2i_attention (same contract); decoder_ops.cu keeps the 128-thread kernel for the training path's tests
extern "C" int grove_decoder_t2i_attention_wide(const float* q, const void* k, const void* v, const int* src_of, float* out, float* lse_out, int B,
                                                int T, int N, int heads, int dh, cudaStream_t stream) {
  GROVE_CHECK_ARG(q && k && v && out && B > 0 && N > 0 && heads > 0);
  if (T != 6 || dh != 16) { grove_set_error("t2i attention is built for T=6 tokens, 16-dim heads (got T=%d dh=%d)", T, dh); return GROVE_ERR_UNSUPPORTED; }
  t2i_attention_wide_kernel<6, 16, 256><<<dim3(B, heads), 256, 0, stream>>>(q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, src_of, out, lse_out,
                                                                           N, heads);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}
