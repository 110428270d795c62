// Backward-pass helpers of the grounding branch's training step (BASELINE config 4; SURVEY.md §8a row "4-bwd"):
// everything around the tensor-core contractions (which reuse gemm_tcgen05.cu with transposed / flipped weights):
//   * layout: transposition to bf16 (operands of the weight-gradient GEMMs), split-K reduction, casts;
//   * LayerNorm backward (encoder blocks: frozen affine; decoder: with d gamma / d beta);
//   * adapter gate backward (tanh(alpha) * relu(conv) + x, image_encoder.py:54), column sums for bias gradients;
//   * the token-side (6 tokens x 256, fp32) weight gradients and activation derivatives of the two-way transformer;
//   * the loss derivative (GIoU + L1 + BCE, GROVE.py:339-381).
// All HBM-bound: 16-byte accesses, one warp per row where a row reduction is needed, per-CTA partial sums + one atomicAdd
// per column for parameter gradients.
#include "common.cuh"
#include "grove_b200.h"

namespace grove {

// ---------------------------------------------------------------- transpose to bf16: in [R,C] (fp32 | bf16) -> out [C,R] bf16
template <bool IN_F32>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const void* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C) {
  __shared__ float tile[64][65];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    float a = 0.f, b = 0.f;
    if (r < R && c < C) {
      if (IN_F32) {
        const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(in) + (size_t)r * C + c);
        a = v.x; b = v.y;
      } else {
        const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(in) + (size_t)r * C + c));
        a = v.x; b = v.y;
      }
    }
    tile[ty + 8 * i][2 * tx] = a;
    tile[ty + 8 * i][2 * tx + 1] = b;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + 2 * tx;
    if (c < C && r < R) *reinterpret_cast<uint32_t*>(out + (size_t)c * R + r) = pack_bf16(tile[2 * tx][ty + 8 * i], tile[2 * tx + 1][ty + 8 * i]);
  }
}

// in [R,C] token-major (R = frames*G*G) -> out [3, C, R] bf16: plane d holds the rows shifted by dw = d-1 along the grid's w axis
// (zero where w+dw leaves [0,G)).  TMA cannot start a box at an odd element of the innermost dimension, so the conv weight-gradient
// GEMM reads the w-shift from the plane index and shifts h, t through the (outer) box coordinates.  G divides 64: no tile halo.
template <bool IN_F32>
__global__ void __launch_bounds__(256) transpose_shift3_kernel(const void* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C, int G) {
  __shared__ float tile[64][65];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    float a = 0.f, b = 0.f;
    if (r < R && c < C) {
      if (IN_F32) {
        const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(in) + (size_t)r * C + c);
        a = v.x; b = v.y;
      } else {
        const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(in) + (size_t)r * C + c));
        a = v.x; b = v.y;
      }
    }
    tile[ty + 8 * i][2 * tx] = a;
    tile[ty + 8 * i][2 * tx + 1] = b;
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cl = ty + 8 * i, c = c0 + cl, rl = 2 * tx, r = r0 + rl;
      if (c < C && r < R) {
        const int w0 = (r % G) + d - 1, w1 = w0 + 1;
        const float a = (w0 >= 0 && w0 < G) ? tile[rl + d - 1][cl] : 0.f;
        const float b = (w1 >= 0 && w1 < G) ? tile[rl + d][cl] : 0.f;
        *reinterpret_cast<uint32_t*>(out + ((size_t)d * C + c) * R + r) = pack_bf16(a, b);
      }
    }
}

// out[i] = (accumulate ? out[i] : 0) + scale * sum_s partials[s, i]
__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, long long n4, float* __restrict__ out, int accumulate, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = reinterpret_cast<const float4*>(part)[(long long)s * n4 + i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    if (accumulate) {
      const float4 o = reinterpret_cast<float4*>(out)[i];
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

// ---------------------------------------------------------------- LayerNorm backward, one warp per row (grid-stride over rows)
// XMODE 0: u = x (+ r), both fp32 [rows, D].   XMODE 1: u = bf16 keys[src_of[row / N] * N + row % N] + fp32 delta[row]  (norm4).
// dx_out = (dx_in ? dx_in : 0) + dLN(dy);  optional bf16 copy;  optional d gamma / d beta accumulated with atomics.
constexpr int kLnMaxV = 10;  // D <= 1280
// NV = D / 128 and PG (d gamma / d beta wanted) are compile-time: with run-time bounds every array was sized for D = 1280 and the
// parameter-gradient accumulators were always live -- 187 registers, one 8-warp CTA per SM, 280 us per call (~1 TB/s) on an HBM-bound op.
template <int XMODE, bool DY_F32, int NV, bool PG>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ x, const float* __restrict__ r, const int* __restrict__ src_of, int N,
                                                            const float* __restrict__ gamma, const void* __restrict__ dy,
                                                            const float* __restrict__ dx_in, float* __restrict__ dx_out,
                                                            __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int nv = NV;
  float4 ag[PG ? NV : 1], ab[PG ? NV : 1];
#pragma unroll
  for (int i = 0; i < (PG ? NV : 1); ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    float4 u[NV], g[NV];
    float s = 0.f;
    size_t kbase = 0;
    if (XMODE == 1) kbase = ((size_t)(src_of ? src_of[row / N] : row / N) * N + row % N) * D;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      {
        const int c4 = i * 32 + lane;
        if (XMODE == 0) {
          u[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + (size_t)row * D)[c4];
          if (r) {
            const float4 t = reinterpret_cast<const float4*>(r + (size_t)row * D)[c4];
            u[i].x += t.x; u[i].y += t.y; u[i].z += t.z; u[i].w += t.w;
          }
        } else {
          const uint2 kv = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + kbase)[c4];
          const float4 d = reinterpret_cast<const float4*>(r + (size_t)row * D)[c4];
          const float2 p0 = unpack_bf16(kv.x), p1 = unpack_bf16(kv.y);
          u[i] = make_float4(p0.x + d.x, p0.y + d.y, p1.x + d.z, p1.y + d.w);
        }
        s += u[i].x + u[i].y + u[i].z + u[i].w;
      }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      {
        u[i].x -= mean; u[i].y -= mean; u[i].z -= mean; u[i].w -= mean;
        q += u[i].x * u[i].x + u[i].y * u[i].y + u[i].z * u[i].z + u[i].w * u[i].w;
      }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      {
        const int c4 = i * 32 + lane;
        float4 d;
        if (DY_F32) {
          d = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + (size_t)row * D)[c4];
        } else {
          const uint2 dv = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)row * D)[c4];
          const float2 p0 = unpack_bf16(dv.x), p1 = unpack_bf16(dv.y);
          d = make_float4(p0.x, p0.y, p1.x, p1.y);
        }
        u[i].x *= rstd; u[i].y *= rstd; u[i].z *= rstd; u[i].w *= rstd;   // xhat
        if (PG) {
          ag[i].x += d.x * u[i].x; ag[i].y += d.y * u[i].y; ag[i].z += d.z * u[i].z; ag[i].w += d.w * u[i].w;
          ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
        }
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        sg += g[i].x + g[i].y + g[i].z + g[i].w;
        sgx += g[i].x * u[i].x + g[i].y * u[i].y + g[i].z * u[i].z + g[i].w * u[i].w;
      }
    const float mg = warp_sum(sg) / (float)D, mgx = warp_sum(sgx) / (float)D;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      {
        const int c4 = i * 32 + lane;
        float4 o = make_float4(rstd * (g[i].x - mg - u[i].x * mgx), rstd * (g[i].y - mg - u[i].y * mgx), rstd * (g[i].z - mg - u[i].z * mgx),
                               rstd * (g[i].w - mg - u[i].w * mgx));
        if (dx_in) {
          const float4 t = reinterpret_cast<const float4*>(dx_in + (size_t)row * D)[c4];
          o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
        }
        if (dx_out) reinterpret_cast<float4*>(dx_out + (size_t)row * D)[c4] = o;
        if (dx_bf16) reinterpret_cast<uint2*>(dx_bf16 + (size_t)row * D)[c4] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
  }
  if (PG) {
    __shared__ float red[8][32 * 4];
#pragma unroll
    for (int i = 0; i < (PG ? NV : 0); ++i) {
      const float4 a = ag[i], b = ab[i];
      for (int pass = 0; pass < 2; ++pass) {
        const float4 v = pass ? b : a;
        __syncthreads();
        red[wib][lane * 4 + 0] = v.x; red[wib][lane * 4 + 1] = v.y; red[wib][lane * 4 + 2] = v.z; red[wib][lane * 4 + 3] = v.w;
        __syncthreads();
        if (threadIdx.x < 128) {
          float t = 0.f;
          for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
          atomicAdd((pass ? dbeta : dgamma) + i * 128 + threadIdx.x, t);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- adapter gate backward (image_encoder.py:54)
// y = x + tanh(alpha) * relu(conv + b), with relu_out = relu(conv + b) saved in bf16 by the forward GEMM:
//   dyc = dy * tanh(alpha) * [relu_out > 0]  (bf16, operand of the Conv3d dgrad / wgrad GEMMs)
//   dbias[c] += sum_rows dyc ;  dalpha += (1 - tanh(alpha)^2) * sum dy * relu_out
__global__ void __launch_bounds__(256) adapter_gate_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ relu_out,
                                                               const float* __restrict__ alpha, __nv_bfloat16* __restrict__ dyc,
                                                               float* __restrict__ dbias, float* __restrict__ dalpha, long long rows, int D) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nv = D / 128;
  const float gate = tanhf(__ldg(alpha));
  float4 ab[kLnMaxV];
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float sa = 0.f;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i)
      if (i < nv) {
        const int c4 = i * 32 + lane;
        const float4 d = reinterpret_cast<const float4*>(dy + (size_t)row * D)[c4];
        const uint2 rv = reinterpret_cast<const uint2*>(relu_out + (size_t)row * D)[c4];
        const float2 r0 = unpack_bf16(rv.x), r1 = unpack_bf16(rv.y);
        sa += d.x * r0.x + d.y * r0.y + d.z * r1.x + d.w * r1.y;
        const float o0 = r0.x > 0.f ? d.x * gate : 0.f, o1 = r0.y > 0.f ? d.y * gate : 0.f;
        const float o2 = r1.x > 0.f ? d.z * gate : 0.f, o3 = r1.y > 0.f ? d.w * gate : 0.f;
        ab[i].x += o0; ab[i].y += o1; ab[i].z += o2; ab[i].w += o3;
        reinterpret_cast<uint2*>(dyc + (size_t)row * D)[c4] = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
      }
  }
  __shared__ float red[8][32 * 4];
  __shared__ float reda[8];
  sa = warp_sum(sa);
  if (lane == 0) reda[wib] = sa;
  for (int i = 0; i < nv; ++i) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kLnMaxV; ++j)
      if (j == i) b = ab[j];
    __syncthreads();
    red[wib][lane * 4 + 0] = b.x; red[wib][lane * 4 + 1] = b.y; red[wib][lane * 4 + 2] = b.z; red[wib][lane * 4 + 3] = b.w;
    __syncthreads();
    if (threadIdx.x < 128) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
      atomicAdd(dbias + i * 128 + threadIdx.x, t);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += reda[w];
    atomicAdd(dalpha, t * (1.f - gate * gate));
  }
}

// ---------------------------------------------------------------- column sums: out[c] += sum_r x[r, c]   (bias gradients)
template <bool IN_F32>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ x, float* __restrict__ out, long long R, int C, long long rows_per_cta) {
  const long long r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r)
      s += IN_F32 ? reinterpret_cast<const float*>(x)[r * C + c] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[r * C + c]);
    atomicAdd(out + c, s);
  }
}

// out[f, :] = sum over b in [off[f], off[f+1]) of x[b, :]    (phrases of a frame share layer-0 keys)
__global__ void segment_sum_kernel(const float* __restrict__ x, const int* __restrict__ off, float* __restrict__ out, long long n4, int accumulate) {
  const int f = blockIdx.y;
  const int b0 = off[f], b1 = off[f + 1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = accumulate ? reinterpret_cast<float4*>(out)[(long long)f * n4 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = b0; b < b1; ++b) {
      const float4 v = reinterpret_cast<const float4*>(x)[(long long)b * n4 + i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(out)[(long long)f * n4 + i] = a;
  }
}

// ---------------------------------------------------------------- token-side fp32 weight gradient: dW[N,K] += dY[R,N]^T . X[R,K]
// CTA = 16 outputs n x 64 inputs k; rows streamed through shared memory 32 at a time; thread = (n, 4 consecutive k).
__global__ void __launch_bounds__(256) small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw, int R, int N,
                                                          int K) {
  __shared__ float sdy[32][17];
  __shared__ float sx[32][64];
  const int n0 = blockIdx.y * 16, k0 = blockIdx.x * 64;
  const int tn = threadIdx.x >> 4, tk = (threadIdx.x & 15) * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r0 = 0; r0 < R; r0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 16; i += 256) {
      const int r = r0 + i / 16, n = n0 + i % 16;
      sdy[i / 16][i % 16] = (r < R && n < N) ? dy[(size_t)r * N + n] : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * 64; i += 256) {
      const int r = r0 + i / 64, k = k0 + i % 64;
      sx[i / 64][i % 64] = (r < R && k < K) ? x[(size_t)r * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const float d = sdy[rr][tn];
      const float4 xv = *reinterpret_cast<const float4*>(&sx[rr][tk]);
      acc[0] += d * xv.x; acc[1] += d * xv.y; acc[2] += d * xv.z; acc[3] += d * xv.w;
    }
  }
  const int n = n0 + tn;
  if (n < N)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (k0 + tk + j < K) dw[(size_t)n * K + k0 + tk + j] += acc[j];
}

// dx = dy * act'(.)  in place semantics allowed (dx may alias dy).  kind 2: ReLU from its output y; 3: sigmoid from its output y.
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n, int kind) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i];
    dx[i] = kind == 2 ? (v > 0.f ? dy[i] : 0.f) : dy[i] * v * (1.f - v);
  }
}

// self-attention among T <= 8 tokens, backward (transformer.py:155-161): thread = (head, query); dk/dv accumulated through smem
__global__ void token_self_attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                const float* __restrict__ dout, float* __restrict__ dq, float* __restrict__ dk,
                                                float* __restrict__ dv, int T, int heads, int dh) {
  extern __shared__ float sm[];  // dk [T*HD], dv [T*HD]
  const int b = blockIdx.x, HD = heads * dh;
  float* sdk = sm;
  float* sdv = sm + T * HD;
  for (int i = threadIdx.x; i < 2 * T * HD; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < heads * T; i += blockDim.x) {
    const int h = i / T, tq = i % T;
    const float* qp = q + ((size_t)b * T + tq) * HD + h * dh;
    const float* dop = dout + ((size_t)b * T + tq) * HD + h * dh;
    const float scale = rsqrtf((float)dh);
    float p[8], dp[8], mx = -INFINITY;
    for (int tk = 0; tk < T; ++tk) {
      const float* kp = k + ((size_t)b * T + tk) * HD + h * dh;
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a += qp[d] * kp[d];
      p[tk] = a * scale;
      mx = fmaxf(mx, p[tk]);
    }
    float l = 0.f;
    for (int tk = 0; tk < T; ++tk) { p[tk] = expf(p[tk] - mx); l += p[tk]; }
    float dsum = 0.f;
    for (int tk = 0; tk < T; ++tk) {
      p[tk] /= l;
      const float* vp = v + ((size_t)b * T + tk) * HD + h * dh;
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a += dop[d] * vp[d];
      dp[tk] = a;
      dsum += p[tk] * a;
    }
    for (int d = 0; d < dh; ++d) {
      float a = 0.f;
      for (int tk = 0; tk < T; ++tk) a += p[tk] * (dp[tk] - dsum) * k[((size_t)b * T + tk) * HD + h * dh + d];
      dq[((size_t)b * T + tq) * HD + h * dh + d] = a * scale;
    }
    for (int tk = 0; tk < T; ++tk) {
      const float ds = p[tk] * (dp[tk] - dsum) * scale;
      for (int d = 0; d < dh; ++d) {
        atomicAdd(&sdk[tk * HD + h * dh + d], ds * qp[d]);
        atomicAdd(&sdv[tk * HD + h * dh + d], p[tk] * dop[d]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * HD; i += blockDim.x) {
    dk[(size_t)b * T * HD + i] = sdk[i];
    dv[(size_t)b * T * HD + i] = sdv[i];
  }
}

// ---------------------------------------------------------------- loss derivative (GROVE.py:339-381, torchvision giou_loss.py)
// dboxes[b, 0..3] (cxcywh) and dlogits[b] of   wg * (sum GIoU + sum L1) / (n_gt + 1e-8) + wo * sum BCE / (n_pred + 1e-8)
// given through the two prefactors cg = upstream * wg / (n_gt + 1e-8), co = upstream * wo / (n_pred + 1e-8).
__global__ void box_losses_bwd_kernel(const float* __restrict__ boxes, const float* __restrict__ logits, const float* __restrict__ gt,
                                      const uint8_t* __restrict__ sel, const float* __restrict__ labels, float cg, float co,
                                      float* __restrict__ dboxes, float* __restrict__ dlogits, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float z = logits[b];
  dlogits[b] = co * (1.f / (1.f + expf(-z)) - labels[b]);
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  if (sel[b]) {
    const float eps = 1e-7f;
    const float cx = boxes[4 * b], cy = boxes[4 * b + 1], w = boxes[4 * b + 2], h = boxes[4 * b + 3];
    const float gcx = gt[4 * b], gcy = gt[4 * b + 1], gw = gt[4 * b + 2], gh = gt[4 * b + 3];
    const float x1 = cx - w / 2, y1 = cy - h / 2, x2 = cx + w / 2, y2 = cy + h / 2;
    const float x1g = gcx - gw / 2, y1g = gcy - gh / 2, x2g = gcx + gw / 2, y2g = gcy + gh / 2;
    const float xk1 = fmaxf(x1, x1g), yk1 = fmaxf(y1, y1g), xk2 = fminf(x2, x2g), yk2 = fminf(y2, y2g);
    const bool has = (yk2 > yk1) && (xk2 > xk1);
    const float inter = has ? (xk2 - xk1) * (yk2 - yk1) : 0.f;
    const float area_p = (x2 - x1) * (y2 - y1), area_g = (x2g - x1g) * (y2g - y1g);
    const float uni = area_p + area_g - inter;
    const float xc1 = fminf(x1, x1g), yc1 = fminf(y1, y1g), xc2 = fmaxf(x2, x2g), yc2 = fmaxf(y2, y2g);
    const float ac = (xc2 - xc1) * (yc2 - yc1);
    // loss = 1 - inter/(uni+eps) + (ac - uni)/(ac+eps)
    const float dl_dinter = -1.f / (uni + eps);
    const float dl_duni = inter / ((uni + eps) * (uni + eps)) - 1.f / (ac + eps);
    const float dl_dac = (uni + eps) / ((ac + eps) * (ac + eps));   // d/dac (ac-uni)/(ac+eps) = (uni+eps)/(ac+eps)^2
    // partials w.r.t. the predicted corners (torch.max / torch.min route the gradient to the selected operand; ties: to the prediction
    // for max(x1,x1g) when x1 >= x1g — matches torch's convention of splitting only on exact ties, which have measure zero here)
    float g_x1 = 0.f, g_y1 = 0.f, g_x2 = 0.f, g_y2 = 0.f;
    // inter = (xk2 - xk1) * (yk2 - yk1)
    const float dI = dl_dinter - dl_duni;  // uni = area_p + area_g - inter
    if (has) {
      const float iw = xk2 - xk1, ih = yk2 - yk1;
      if (x1 > x1g) g_x1 += dI * (-ih);
      if (x2 < x2g) g_x2 += dI * ih;
      if (y1 > y1g) g_y1 += dI * (-iw);
      if (y2 < y2g) g_y2 += dI * iw;
    }
    // area_p = (x2-x1)*(y2-y1) enters uni
    g_x1 += dl_duni * (-(y2 - y1)); g_x2 += dl_duni * (y2 - y1);
    g_y1 += dl_duni * (-(x2 - x1)); g_y2 += dl_duni * (x2 - x1);
    // enclosing box
    const float cw = xc2 - xc1, ch = yc2 - yc1;
    if (x1 < x1g) g_x1 += dl_dac * (-ch);
    if (x2 > x2g) g_x2 += dl_dac * ch;
    if (y1 < y1g) g_y1 += dl_dac * (-cw);
    if (y2 > y2g) g_y2 += dl_dac * cw;
    // corners -> cxcywh
    d[0] = g_x1 + g_x2; d[1] = g_y1 + g_y2; d[2] = 0.5f * (g_x2 - g_x1); d[3] = 0.5f * (g_y2 - g_y1);
    // L1 on cxcywh
    d[0] += (cx > gcx) ? 1.f : (cx < gcx ? -1.f : 0.f);
    d[1] += (cy > gcy) ? 1.f : (cy < gcy ? -1.f : 0.f);
    d[2] += (w > gw) ? 1.f : (w < gw ? -1.f : 0.f);
    d[3] += (h > gh) ? 1.f : (h < gh ? -1.f : 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] *= cg;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) dboxes[4 * b + j] = d[j];
}

}  // namespace grove
using namespace grove;

static inline int bw_grid(long long n, int block, int per_sm = 8) {
  long long g = (n + block - 1) / block, cap = (long long)kNumSMs * per_sm;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" int grove_transpose_to_bf16(const void* in, int in_is_f32, void* out, int R, int C, cudaStream_t stream) {
  GROVE_CHECK_ARG(in && out && R > 0 && C > 0 && R % 2 == 0 && C % 2 == 0);
  dim3 grid((C + 63) / 64, (R + 63) / 64);
  if (in_is_f32) transpose_to_bf16_kernel<true><<<grid, 256, 0, stream>>>(in, (__nv_bfloat16*)out, R, C);
  else transpose_to_bf16_kernel<false><<<grid, 256, 0, stream>>>(in, (__nv_bfloat16*)out, R, C);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_transpose_shift3_to_bf16(const void* in, int in_is_f32, void* out, int R, int C, int G, cudaStream_t stream) {
  GROVE_CHECK_ARG(in && out && R > 0 && C > 0 && R % 64 == 0 && C % 2 == 0 && G > 0 && 64 % G == 0);
  dim3 grid((C + 63) / 64, R / 64);
  if (in_is_f32) transpose_shift3_kernel<true><<<grid, 256, 0, stream>>>(in, (__nv_bfloat16*)out, R, C, G);
  else transpose_shift3_kernel<false><<<grid, 256, 0, stream>>>(in, (__nv_bfloat16*)out, R, C, G);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_reduce_partials_f32(const float* partials, int splits, long long n, float* out, int accumulate, float scale, cudaStream_t stream) {
  GROVE_CHECK_ARG(partials && out && splits > 0 && n > 0 && n % 4 == 0);
  reduce_partials_kernel<<<bw_grid(n / 4, 256), 256, 0, stream>>>(partials, splits, n / 4, out, accumulate, scale);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_layernorm_bwd(const void* x, const float* r, const int* src_of, int N, int x_is_keys_bf16, const float* gamma, const void* dy,
                                   int dy_is_f32, const float* dx_in, float* dx_out, void* dx_bf16, float* dgamma, float* dbeta, long long rows,
                                   int D, float eps, cudaStream_t stream) {
  GROVE_CHECK_ARG(x && gamma && dy && rows > 0 && D % 128 == 0 && D <= 128 * kLnMaxV);
  GROVE_CHECK_ARG((dx_out || dx_bf16) && ((dgamma == nullptr) == (dbeta == nullptr)));
  GROVE_CHECK_ARG(!x_is_keys_bf16 || (r && N > 0));
  const int grid = bw_grid(rows * 32, 256, 8);
  auto* o16 = (__nv_bfloat16*)dx_bf16;
#define LNB3(XM, DF, NVV, PGG) layernorm_bwd_kernel<XM, DF, NVV, PGG><<<grid, 256, 0, stream>>>(x, r, src_of, N, gamma, dy, dx_in, dx_out, o16, dgamma, dbeta, rows, D, eps)
#define LNB2(XM, DF, NVV) do { if (dgamma) LNB3(XM, DF, NVV, true); else LNB3(XM, DF, NVV, false); } while (0)
#define LNB(XM, DF) do { switch (D / 128) { case 2: LNB2(XM, DF, 2); break; case 6: LNB2(XM, DF, 6); break; case 8: LNB2(XM, DF, 8); break; \
                                            case 10: LNB2(XM, DF, 10); break; default: grove_set_error("grove_layernorm_bwd: D = %d is not built (256, 768, 1024, 1280)", D); return GROVE_ERR_UNSUPPORTED; } } while (0)
  if (x_is_keys_bf16) { if (dy_is_f32) LNB(1, true); else LNB(1, false); }
  else { if (dy_is_f32) LNB(0, true); else LNB(0, false); }
#undef LNB
#undef LNB2
#undef LNB3
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_adapter_gate_bwd(const float* dy, const void* relu_out, const float* alpha, void* dyc, float* dbias, float* dalpha,
                                      long long rows, int D, cudaStream_t stream) {
  GROVE_CHECK_ARG(dy && relu_out && alpha && dyc && dbias && dalpha && rows > 0 && D % 128 == 0 && D <= 128 * kLnMaxV);
  adapter_gate_bwd_kernel<<<bw_grid(rows * 32, 256, 4), 256, 0, stream>>>(dy, (const __nv_bfloat16*)relu_out, alpha, (__nv_bfloat16*)dyc, dbias,
                                                                           dalpha, rows, D);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_colsum(const void* x, int x_is_f32, float* out, long long R, int C, cudaStream_t stream) {
  GROVE_CHECK_ARG(x && out && R > 0 && C > 0);
  const int gx = (C + 255) / 256;
  long long gy = (kNumSMs * 4 + gx - 1) / gx;
  if (gy > (R + 63) / 64) gy = (R + 63) / 64;
  if (gy < 1) gy = 1;
  const long long rpc = (R + gy - 1) / gy;
  if (x_is_f32) colsum_kernel<true><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(x, out, R, C, rpc);
  else colsum_kernel<false><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(x, out, R, C, rpc);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_segment_sum_f32(const float* x, const int* offsets, float* out, int segments, long long n, int accumulate, cudaStream_t stream) {
  GROVE_CHECK_ARG(x && offsets && out && segments > 0 && segments <= 65535 && n > 0 && n % 4 == 0);
  int gx = bw_grid(n / 4, 256, 2);
  segment_sum_kernel<<<dim3(gx, segments), 256, 0, stream>>>(x, offsets, out, n / 4, accumulate);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_small_wgrad_f32(const float* dy, const float* x, float* dw, int R, int N, int K, cudaStream_t stream) {
  GROVE_CHECK_ARG(dy && x && dw && R > 0 && N > 0 && K > 0);
  small_wgrad_kernel<<<dim3((K + 63) / 64, (N + 15) / 16), 256, 0, stream>>>(dy, x, dw, R, N, K);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_act_bwd_f32(const float* dy, const float* y, float* dx, long long n, int kind, cudaStream_t stream) {
  GROVE_CHECK_ARG(dy && y && dx && n > 0 && (kind == 2 || kind == 3));
  act_bwd_kernel<<<bw_grid(n, 256), 256, 0, stream>>>(dy, y, dx, n, kind);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_token_self_attention_bwd(const float* q, const float* k, const float* v, const float* dout, float* dq, float* dk, float* dv,
                                              int B, int T, int heads, int dh, cudaStream_t stream) {
  GROVE_CHECK_ARG(q && k && v && dout && dq && dk && dv && B > 0 && T > 0 && T <= 8 && heads > 0 && dh > 0);
  const int smem = 2 * T * heads * dh * (int)sizeof(float);
  GROVE_CHECK_ARG(smem <= 48 * 1024);
  token_self_attention_bwd_kernel<<<B, 64, smem, stream>>>(q, k, v, dout, dq, dk, dv, T, heads, dh);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_box_losses_bwd(const float* boxes, const float* logits, const float* gt, const uint8_t* sel, const float* labels, float cg,
                                    float co, float* dboxes, float* dlogits, int B, cudaStream_t stream) {
  GROVE_CHECK_ARG(boxes && logits && gt && sel && labels && dboxes && dlogits && B > 0);
  box_losses_bwd_kernel<<<(B + 127) / 128, 128, 0, stream>>>(boxes, logits, gt, sel, labels, cg, co, dboxes, dlogits, B);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}
