// Prompt-encoder / two-way box-decoder kernels (prompt_encoder.py:164-229, transformer.py:62-242,
// mask_decoder.py:155-205).  The four [N,256]<->128 image-side projections per layer run on the tcgen05 GEMM
// (gemm_tcgen05.cu); everything here is the HBM/latency-bound remainder:
//   * the 6-token side is kept in fp32 (tiny: 6 x 256 per instance),
//   * the image side ("keys", N x 256 per instance) is bf16 in HBM and is read/written once per stage,
//   * layer 0's image-side projections are shared by all phrases of a frame (keys only become
//     per-instance after the first image-to-token update), expressed through `src_of[b]` index arrays.
#include "common.cuh"
#include "grove_b200.h"

namespace grove {

// ---------------------------------------------------------------- gather rows (the [DET] hidden states)
template <bool SRC_F32>
__global__ void gather_rows_kernel(const void* __restrict__ src, const int* __restrict__ idx, __nv_bfloat16* __restrict__ dst, int D) {
  const int r = blockIdx.x;
  const size_t so = (size_t)idx[r] * D;
  for (int c = threadIdx.x * 2; c < D; c += blockDim.x * 2) {
    float a, b;
    if (SRC_F32) {
      const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(src) + so + c);
      a = v.x; b = v.y;
    } else {
      const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(src) + so + c));
      a = v.x; b = v.y;
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)r * D + c) = pack_bf16(a, b);
  }
}

// ---------------------------------------------------------------- dense positional encoding (random Fourier features)
__global__ void dense_pe_kernel(const float* __restrict__ gauss, float* __restrict__ pe, int G, int F2) {
  const int n = blockIdx.x;  // token = y*G + x
  const float cx = 2.f * (((float)(n % G) + 0.5f) / (float)G) - 1.f;
  const float cy = 2.f * (((float)(n / G) + 0.5f) / (float)G) - 1.f;
  for (int j = threadIdx.x; j < F2; j += blockDim.x) {
    const float a = 6.283185307179586f * (cx * gauss[j] + cy * gauss[F2 + j]);
    float s, c;
    sincosf(a, &s, &c);
    pe[(size_t)n * 2 * F2 + j] = s;
    pe[(size_t)n * 2 * F2 + F2 + j] = c;
  }
}

// keys[f,n,:] = bf16(emb[f,n,:] + no_mask[:])
__global__ void add_rowvec_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ vec, __nv_bfloat16* __restrict__ y,
                                       long long rows, int C) {
  const int c8 = C / 8;
  const long long total = rows * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const uint4 v = reinterpret_cast<const uint4*>(x)[i];
    const float4 a = *reinterpret_cast<const float4*>(vec + c), b = *reinterpret_cast<const float4*>(vec + c + 4);
    float2 p0 = unpack_bf16(v.x), p1 = unpack_bf16(v.y), p2 = unpack_bf16(v.z), p3 = unpack_bf16(v.w);
    reinterpret_cast<uint4*>(y)[i] = make_uint4(pack_bf16(p0.x + a.x, p0.y + a.y), pack_bf16(p1.x + a.z, p1.y + a.w),
                                                pack_bf16(p2.x + b.x, p2.y + b.y), pack_bf16(p3.x + b.z, p3.y + b.w));
  }
}

// ---------------------------------------------------------------- token -> image attention
// grid (B, heads); block 128.  q fp32 [B,T,H*DH]; k,v bf16 [*,N,H*DH] with row block src_of[b].
// Each thread walks keys n = tid, tid+128, ... keeping an online-softmax state per token; the 128 partial
// states are merged through shared memory.
template <int T, int DH>
__global__ void __launch_bounds__(128) t2i_attention_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                            const __nv_bfloat16* __restrict__ v, const int* __restrict__ src_of,
                                                            float* __restrict__ out, float* __restrict__ lse_out, int N, int heads) {
  static_assert(DH == 16, "one key/value head row = two 16-byte loads");
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int HD = heads * DH;
  const float scale_log2 = rsqrtf((float)DH) * 1.4426950408889634f;
  __shared__ float qs[T][DH];
  __shared__ float red_m[T][128], red_l[T][128];
  __shared__ float red_acc[T][DH][4];
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
  for (int n = tid; n < N; n += 128) {
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
  // merge the 128 per-thread states
#pragma unroll
  for (int t = 0; t < T; ++t) { red_m[t][tid] = m[t]; }
  __syncthreads();
  float gm[T], w[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float x = -INFINITY;
    for (int i = 0; i < 128; ++i) x = fmaxf(x, red_m[t][i]);  // smem broadcast reads
    gm[t] = x;
    w[t] = (m[t] == -INFINITY) ? 0.f : exp2f(m[t] - x);
    red_l[t][tid] = l[t] * w[t];
  }
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      const float s = warp_sum(acc[t][d] * w[t]);
      if (lane == 0) red_acc[t][d][warp] = s;
    }
  __syncthreads();
  if (tid < T * DH) {
    const int t = tid / DH, d = tid % DH;
    float lsum = 0.f;
    for (int i = 0; i < 128; ++i) lsum += red_l[t][i];
    const float a = red_acc[t][d][0] + red_acc[t][d][1] + red_acc[t][d][2] + red_acc[t][d][3];
    out[((size_t)b * T + t) * HD + h * DH + d] = a / lsum;
    if (lse_out && d == 0) lse_out[((size_t)b * T + t) * heads + h] = gm[t] + log2f(lsum);   // log2 domain, scale folded in (backward pass)
  }
}

constexpr int kI2tIter = 4;
// ---------------------------------------------------------------- image -> token attention (per key row, T tokens)
// thread = (row n, head h), head fastest so a warp reads 4 contiguous 256-byte rows.
// qi bf16 [*,N,H*DH] (row block src_of[b]); kt,vt fp32 [B,T,H*DH]; out bf16 [B,N,H*DH].
template <int T, int DH>
__global__ void __launch_bounds__(256) i2t_attention_kernel(const __nv_bfloat16* __restrict__ qi, const float* __restrict__ kt,
                                                            const float* __restrict__ vt, const int* __restrict__ src_of,
                                                            __nv_bfloat16* __restrict__ out, int N, int heads) {
  static_assert(DH == 16, "");
  extern __shared__ float sm[];
  const int b = blockIdx.y;
  const int HD = heads * DH;
  const int HS = DH + 1;  // padded head stride: conflict-free across heads
  float* ks = sm;                         // [T][heads][HS]
  float* vs = sm + T * heads * HS;
  for (int i = threadIdx.x; i < T * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD;
    ks[(t * heads + c / DH) * HS + c % DH] = kt[(size_t)b * T * HD + i] * (rsqrtf((float)DH) * 1.4426950408889634f);
    vs[(t * heads + c / DH) * HS + c % DH] = vt[(size_t)b * T * HD + i];
  }
  __syncthreads();
  // kI2tIter (row, head) pairs per thread: the 6 KB k/v preload above is amortised over 4x the rows (4096 CTAs of one pair each spent
  // most of their time on it)
#pragma unroll 1
  for (int it = 0; it < kI2tIter; ++it) {
    const int gid = (blockIdx.x * kI2tIter + it) * blockDim.x + threadIdx.x;
    const int n = gid / heads, h = gid % heads;
    if (n >= N) return;
    const size_t qoff = ((size_t)(src_of ? src_of[b] : b) * N + n) * HD + h * DH;
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(qi + qoff)), q1 = __ldg(reinterpret_cast<const uint4*>(qi + qoff) + 1);
    const uint32_t qu[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float qf[DH];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float2 a = unpack_bf16(qu[i]); qf[2 * i] = a.x; qf[2 * i + 1] = a.y; }
    float s[T], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) a += qf[d] * ks[(t * heads + h) * HS + d];
      s[t] = a;
      mx = fmaxf(mx, a);
    }
    float l = 0.f, o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float p = exp2f(s[t] - mx);
      l += p;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] += p * vs[(t * heads + h) * HS + d];
    }
    const float inv = 1.f / l;
    uint4 w0 = make_uint4(pack_bf16(o[0] * inv, o[1] * inv), pack_bf16(o[2] * inv, o[3] * inv), pack_bf16(o[4] * inv, o[5] * inv),
                          pack_bf16(o[6] * inv, o[7] * inv));
    uint4 w1 = make_uint4(pack_bf16(o[8] * inv, o[9] * inv), pack_bf16(o[10] * inv, o[11] * inv), pack_bf16(o[12] * inv, o[13] * inv),
                          pack_bf16(o[14] * inv, o[15] * inv));
    uint4* op = reinterpret_cast<uint4*>(out + ((size_t)b * N + n) * HD + h * DH);
    op[0] = w0;
    op[1] = w1;
  }
}

// keys_out[b,n,:] = bf16( LN( keys_in[src_of[b],n,:] + delta[b,n,:] ) ), C = 256, one warp per row (norm4, transformer.py:180)
__global__ void __launch_bounds__(256) keys_add_ln_kernel(const __nv_bfloat16* __restrict__ keys_in, const int* __restrict__ src_of,
                                                          const float* __restrict__ delta, const float* __restrict__ g,
                                                          const float* __restrict__ be, __nv_bfloat16* __restrict__ keys_out, int B, int N,
                                                          float eps) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), n = (int)(row % N);
  const size_t in_off = ((size_t)(src_of ? src_of[b] : b) * N + n) * 256 + lane * 8;
  const uint4 kv = *reinterpret_cast<const uint4*>(keys_in + in_off);
  const float4 d0 = *reinterpret_cast<const float4*>(delta + (size_t)row * 256 + lane * 8);
  const float4 d1 = *reinterpret_cast<const float4*>(delta + (size_t)row * 256 + lane * 8 + 4);
  float x[8];
  { float2 p = unpack_bf16(kv.x); x[0] = p.x + d0.x; x[1] = p.y + d0.y; }
  { float2 p = unpack_bf16(kv.y); x[2] = p.x + d0.z; x[3] = p.y + d0.w; }
  { float2 p = unpack_bf16(kv.z); x[4] = p.x + d1.x; x[5] = p.y + d1.y; }
  { float2 p = unpack_bf16(kv.w); x[6] = p.x + d1.z; x[7] = p.y + d1.w; }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] -= mean; qv += x[i] * x[i]; }
  const float rstd = rsqrtf(warp_sum(qv) * (1.f / 256.f) + eps);
  const float4 g0 = *reinterpret_cast<const float4*>(g + lane * 8), g1 = *reinterpret_cast<const float4*>(g + lane * 8 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(be + lane * 8), b1 = *reinterpret_cast<const float4*>(be + lane * 8 + 4);
  const float y0 = x[0] * rstd * g0.x + b0.x, y1 = x[1] * rstd * g0.y + b0.y, y2 = x[2] * rstd * g0.z + b0.z, y3 = x[3] * rstd * g0.w + b0.w;
  const float y4 = x[4] * rstd * g1.x + b1.x, y5 = x[5] * rstd * g1.y + b1.y, y6 = x[6] * rstd * g1.z + b1.z, y7 = x[7] * rstd * g1.w + b1.w;
  *reinterpret_cast<uint4*>(keys_out + (size_t)row * 256 + lane * 8) =
      make_uint4(pack_bf16(y0, y1), pack_bf16(y2, y3), pack_bf16(y4, y5), pack_bf16(y6, y7));
}

// ---------------------------------------------------------------- token-side fp32 helpers
// y[R,N] = act(x[R,K] . W[N,K]^T + b) (+ resid).  CTA = 8 rows x 16 outputs (two per warp); x rows in smem; lanes split K.
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                                           const float* __restrict__ resid, float* __restrict__ y, int R, int N, int K, int act) {
  extern __shared__ float xs[];  // [8][K]
  const int r0 = blockIdx.y * 8, n0 = blockIdx.x * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * K; i += 256) {
    const int r = r0 + i / K;
    xs[i] = r < R ? x[(size_t)r * K + i % K] : 0.f;
  }
  __syncthreads();
  for (int j = warp; j < 16; j += 8) {
    const int n = n0 + j;
    if (n >= N) break;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* wr = W + (size_t)n * K;
    for (int kk = lane * 4; kk < K; kk += 128) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wr + kk));
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(xs + r * K + kk);
        acc[r] += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = warp_sum(acc[r]);
    if (lane < 8 && r0 + lane < R) {
      float v = acc[0];
#pragma unroll
      for (int r = 1; r < 8; ++r) v = (lane == r) ? acc[r] : v;
      v += bias ? bias[n] : 0.f;
      if (act == 1) v = gelu_erf(v);
      else if (act == 2) v = fmaxf(v, 0.f);
      else if (act == 3) v = 1.f / (1.f + expf(-v));
      const size_t o = (size_t)(r0 + lane) * N + n;
      if (resid) v += resid[o];
      y[o] = v;
    }
  }
}

// self-attention among T (<= 8) tokens: one CTA per instance, thread = (head, query token)
__global__ void token_self_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                            float* __restrict__ out, int T, int heads, int dh) {
  const int b = blockIdx.x, HD = heads * dh;
  for (int i = threadIdx.x; i < heads * T; i += blockDim.x) {
    const int h = i / T, tq = i % T;
    const float* qp = q + ((size_t)b * T + tq) * HD + h * dh;
    float s[8], mx = -INFINITY;
    for (int tk = 0; tk < T; ++tk) {
      const float* kp = k + ((size_t)b * T + tk) * HD + h * dh;
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a += qp[d] * kp[d];
      s[tk] = a * rsqrtf((float)dh);
      mx = fmaxf(mx, s[tk]);
    }
    float l = 0.f;
    for (int tk = 0; tk < T; ++tk) { s[tk] = expf(s[tk] - mx); l += s[tk]; }
    for (int d = 0; d < dh; ++d) {
      float a = 0.f;
      for (int tk = 0; tk < T; ++tk) a += s[tk] * v[((size_t)b * T + tk) * HD + h * dh + d];
      out[((size_t)b * T + tq) * HD + h * dh + d] = a / l;
    }
  }
}

// y = LN(x (+ r)); y2 = y + add2 (optional).  One warp per row, C % 32 == 0, C <= 1024.
__global__ void __launch_bounds__(256) add_layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ g,
                                                                const float* __restrict__ be, float* __restrict__ y, const float* __restrict__ add2,
                                                                float* __restrict__ y2, int R, int C, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= R) return;
  float v[32];
  const int per = C / 32;
  float s = 0.f;
  for (int i = 0; i < per; ++i) {
    const size_t o = (size_t)row * C + i * 32 + lane;
    v[i] = x[o] + (r ? r[o] : 0.f);
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float qv = 0.f;
  for (int i = 0; i < per; ++i) { v[i] -= mean; qv += v[i] * v[i]; }
  const float rstd = rsqrtf(warp_sum(qv) / (float)C + eps);
  for (int i = 0; i < per; ++i) {
    const int c = i * 32 + lane;
    const size_t o = (size_t)row * C + c;
    const float out = v[i] * rstd * g[c] + be[c];
    y[o] = out;
    if (y2) y2[o] = out + add2[o];
  }
}


// ---------------------------------------------------------------- final out-proj + LayerNorm + heads for the prompt token -> packed record
// One CTA (256 threads) per instance; only token `tok` (index 1 + num_mask_tokens) feeds the heads (mask_decoder.py:192):
//   hs   = LN_final(queries[b,tok,:] + out_proj(att[b,tok,:]))        (transformer.py:99-104)
//   box  = sigmoid(W2 relu(W0 hs + b0) + b2),  logit = Wt hs + bt      (mask_decoder.py:80-85, 198-203)
//   rec[b, 0:4] = box (cx, cy, w, h),  rec[b, 4] = logit               (the record the config-5 all-gather ships)
// A warp owns 32 consecutive outputs of each matrix-vector product; its lanes split K with float4 loads (coalesced weight rows).
template <int K>
__device__ __forceinline__ float warp_dot(const float* __restrict__ w, const float* __restrict__ xs, int lane) {
  float a = 0.f;
#pragma unroll
  for (int kk = lane * 4; kk < K; kk += 128) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + kk));
    const float4 xv = *reinterpret_cast<const float4*>(xs + kk);
    a += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
  }
  return warp_sum(a);
}

__global__ void __launch_bounds__(256) decoder_heads_kernel(const float* __restrict__ queries, const float* __restrict__ att, const float* __restrict__ Wo,
                                                            const float* __restrict__ bo, const float* __restrict__ lg, const float* __restrict__ lb,
                                                            float eps, const float* __restrict__ W0, const float* __restrict__ b0,
                                                            const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ Wt,
                                                            const float* __restrict__ bt, float* __restrict__ rec, float* __restrict__ hs_out, int T,
                                                            int tok) {
  constexpr int C = 256, CI = 128;
  __shared__ __align__(16) float xs[C], hs[C], h1[C];
  __shared__ float red[2][8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t row = (size_t)b * T + tok;
  if (tid < CI) xs[tid] = att[row * CI + tid];
  __syncthreads();
  for (int j = warp * 32; j < warp * 32 + 32; ++j) {
    const float d = warp_dot<CI>(Wo + (size_t)j * CI, xs, lane);
    if (lane == 0) h1[j] = d + bo[j] + queries[row * C + j];
  }
  __syncthreads();
  // LayerNorm over the 256 channels (two-pass, like nn.LayerNorm)
  const float v = h1[tid];
  float s = warp_sum(v);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mean += red[0][i];
  mean *= (1.f / C);
  const float dv = v - mean;
  s = warp_sum(dv * dv);
  if (lane == 0) red[1][warp] = s;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) var += red[1][i];
  const float y = dv * rsqrtf(var * (1.f / C) + eps) * lg[tid] + lb[tid];
  hs[tid] = y;
  if (hs_out) hs_out[(size_t)b * C + tid] = y;
  __syncthreads();
  for (int j = warp * 32; j < warp * 32 + 32; ++j) {
    const float d = warp_dot<C>(W0 + (size_t)j * C, hs, lane);
    if (lane == 0) xs[j] = fmaxf(d + b0[j], 0.f);     // xs is free again: CI <= C
  }
  __syncthreads();
  if (warp < 4) {
    const float d = warp_dot<C>(W2 + (size_t)warp * C, xs, lane);
    if (lane == 0) rec[(size_t)b * 5 + warp] = 1.f / (1.f + expf(-(d + b2[warp])));
  } else if (warp == 4) {
    float d = 0.f;
    if (Wt) d = warp_dot<C>(Wt, hs, lane) + bt[0];
    if (lane == 0) rec[(size_t)b * 5 + 4] = d;
  }
}

}  // namespace grove
using namespace grove;

static inline int grid_cap(long long n, int block, int per_sm = 8) {
  long long gsz = (n + block - 1) / block, cap = (long long)kNumSMs * per_sm;
  return (int)(gsz < cap ? (gsz > 0 ? gsz : 1) : cap);
}

extern "C" int grove_gather_rows_bf16(const void* src, int src_is_f32, const int* row_idx, void* dst, int n_rows, int D, cudaStream_t stream) {
  GROVE_CHECK_ARG(src && row_idx && dst && n_rows > 0 && D % 2 == 0);
  if (src_is_f32) gather_rows_kernel<true><<<n_rows, 256, 0, stream>>>(src, row_idx, (__nv_bfloat16*)dst, D);
  else gather_rows_kernel<false><<<n_rows, 256, 0, stream>>>(src, row_idx, (__nv_bfloat16*)dst, D);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_dense_pe(const float* gauss, float* pe, int G, int F2, cudaStream_t stream) {
  GROVE_CHECK_ARG(gauss && pe && G > 0 && F2 > 0);
  dense_pe_kernel<<<G * G, 128, 0, stream>>>(gauss, pe, G, F2);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_add_rowvec_bf16(const void* x, const float* vec, void* y, long long rows, int C, cudaStream_t stream) {
  GROVE_CHECK_ARG(x && vec && y && rows > 0 && C % 8 == 0);
  add_rowvec_bf16_kernel<<<grid_cap(rows * (C / 8), 256), 256, 0, stream>>>((const __nv_bfloat16*)x, vec, (__nv_bfloat16*)y, rows, C);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_decoder_t2i_attention(const float* q, const void* k, const void* v, const int* src_of, float* out, float* lse_out, int B, int T,
                                           int N, int heads, int dh, cudaStream_t stream) {
  GROVE_CHECK_ARG(q && k && v && out && B > 0 && N > 0 && heads > 0);
  if (T != 6 || dh != 16) { grove_set_error("t2i attention is built for T=6 tokens, 16-dim heads (got T=%d dh=%d)", T, dh); return GROVE_ERR_UNSUPPORTED; }
  t2i_attention_kernel<6, 16><<<dim3(B, heads), 128, 0, stream>>>(q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, src_of, out, lse_out, N, heads);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_decoder_i2t_attention(const void* qi, const float* kt, const float* vt, const int* src_of, void* out, int B, int T, int N,
                                           int heads, int dh, cudaStream_t stream) {
  GROVE_CHECK_ARG(qi && kt && vt && out && B > 0 && N > 0 && heads > 0 && B <= 65535);
  if (T != 6 || dh != 16) { grove_set_error("i2t attention is built for T=6 tokens, 16-dim heads (got T=%d dh=%d)", T, dh); return GROVE_ERR_UNSUPPORTED; }
  const int smem = 2 * T * heads * (dh + 1) * sizeof(float);
  i2t_attention_kernel<6, 16><<<dim3((N * heads + 256 * kI2tIter - 1) / (256 * kI2tIter), B), 256, smem, stream>>>((const __nv_bfloat16*)qi, kt, vt, src_of,
                                                                                       (__nv_bfloat16*)out, N, heads);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_decoder_keys_add_ln(const void* keys_in, const int* src_of, const float* delta, const float* g, const float* b, void* keys_out,
                                         int B, int N, int C, float eps, cudaStream_t stream) {
  GROVE_CHECK_ARG(keys_in && delta && g && b && keys_out && B > 0 && N > 0);
  if (C != 256) { grove_set_error("keys_add_ln is built for C=256 (got %d)", C); return GROVE_ERR_UNSUPPORTED; }
  const long long rows = (long long)B * N;
  keys_add_ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const __nv_bfloat16*)keys_in, src_of, delta, g, b, (__nv_bfloat16*)keys_out, B, N, eps);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_small_linear_f32(const float* x, const float* W, const float* b, const float* resid, float* y, int R, int N, int K, int act,
                                      cudaStream_t stream) {
  GROVE_CHECK_ARG(x && W && y && R > 0 && N > 0 && K > 0 && K % 4 == 0 && K <= 4096 && act >= 0 && act <= 3);
  const int smem = 8 * K * sizeof(float);
  static GrovePerDeviceOnce attr;
  if (attr.first_time())
    cudaFuncSetAttribute(small_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 * (int)sizeof(float));
  small_linear_kernel<<<dim3((N + 15) / 16, (R + 7) / 8), 256, smem, stream>>>(x, W, b, resid, y, R, N, K, act);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_token_self_attention(const float* q, const float* k, const float* v, float* out, int B, int T, int heads, int dh,
                                          cudaStream_t stream) {
  GROVE_CHECK_ARG(q && k && v && out && B > 0 && T > 0 && T <= 8 && heads > 0 && dh > 0);
  token_self_attention_kernel<<<B, 64, 0, stream>>>(q, k, v, out, T, heads, dh);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_add_layernorm_f32(const float* x, const float* r, const float* g, const float* b, float* y, const float* add2, float* y2,
                                       int R, int C, float eps, cudaStream_t stream) {
  GROVE_CHECK_ARG(x && g && b && y && R > 0 && C % 32 == 0 && C <= 1024);
  GROVE_CHECK_ARG((y2 == nullptr) == (add2 == nullptr));
  add_layernorm_f32_kernel<<<(R + 7) / 8, 256, 0, stream>>>(x, r, g, b, y, add2, y2, R, C, eps);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_decoder_heads_fwd(const float* queries, const float* att, const float* Wo, const float* bo, const float* ln_g, const float* ln_b,
                                       float eps, const float* W0, const float* b0, const float* W2, const float* b2, const float* Wt, const float* bt,
                                       float* records, float* hs_out, int B, int T, int tok, int C, int CI, cudaStream_t stream) {
  GROVE_CHECK_ARG(queries && att && Wo && bo && ln_g && ln_b && W0 && b0 && W2 && b2 && records && B > 0 && T > 0 && tok >= 0 && tok < T);
  GROVE_CHECK_ARG((Wt == nullptr) == (bt == nullptr));
  if (C != 256 || CI != 128) { grove_set_error("decoder heads are built for transformer_dim 256 / cross-attention width 128 (got %d / %d)", C, CI); return GROVE_ERR_UNSUPPORTED; }
  decoder_heads_kernel<<<B, 256, 0, stream>>>(queries, att, Wo, bo, ln_g, ln_b, eps, W0, b0, W2, b2, Wt, bt, records, hs_out, T, tok);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}
