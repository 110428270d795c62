// Backward of the two cross-attentions of the two-way box decoder (transformer.py:164-180, 99-104) — training step of the
// grounding branch (BASELINE config 4: the whole mask decoder is trainable, train.py:281-289).
// Same work split as the forward kernels in decoder_ops.cu: the 6-token side in fp32, the image side (N x 128 per
// instance) bf16 in HBM, read once per kernel; reductions over the N image tokens stay on chip.
#include "common.cuh"
#include "grove_b200.h"

namespace grove {

__device__ __forceinline__ void unpack16(const __nv_bfloat16* p, float (&f)[16]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 v = unpack_bf16(u[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}
__device__ __forceinline__ void pack16(__nv_bfloat16* p, const float (&f)[16]) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  reinterpret_cast<uint4*>(p)[1] = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
}

// ---------------------------------------------------------------- token -> image attention, backward
// grid (B, heads), block 128.  Inputs as the forward (+ its output `att`, the upstream gradient `datt`, the saved log2-domain
// log-sum-exp).  Outputs: dq fp32 [B,T,H*DH]; dk, dv bf16 [B,N,H*DH] per INSTANCE (layer 0's shared keys are summed per frame later).
template <int T, int DH>
__global__ void __launch_bounds__(128) t2i_attention_bwd_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                                const __nv_bfloat16* __restrict__ v, const int* __restrict__ src_of,
                                                                const float* __restrict__ att, const float* __restrict__ datt,
                                                                const float* __restrict__ lse, float* __restrict__ dq,
                                                                __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int N, int heads) {
  static_assert(DH == 16, "");
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int HD = heads * DH;
  const float scale = rsqrtf((float)DH), scale_log2 = scale * 1.4426950408889634f;
  __shared__ float qs[T][DH], dos[T][DH], dsum[T], lses[T];
  __shared__ float red[T][DH][4];
  if (tid < T * DH) {
    const size_t o = ((size_t)b * T + tid / DH) * HD + h * DH + tid % DH;
    qs[tid / DH][tid % DH] = q[o];
    dos[tid / DH][tid % DH] = datt[o];
  }
  if (tid < T) {
    float s = 0.f;
    for (int d = 0; d < DH; ++d) {
      const size_t o = ((size_t)b * T + tid) * HD + h * DH + d;
      s += att[o] * datt[o];
    }
    dsum[tid] = s;
    lses[tid] = lse[((size_t)b * T + tid) * heads + h];
  }
  __syncthreads();
  const size_t in_base = (size_t)(src_of ? src_of[b] : b) * N * HD + h * DH;
  const size_t out_base = (size_t)b * N * HD + h * DH;
  float dqa[T][DH];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int d = 0; d < DH; ++d) dqa[t][d] = 0.f;
  for (int n = tid; n < N; n += 128) {
    float kf[DH], vf[DH], dkf[DH], dvf[DH];
    unpack16(k + in_base + (size_t)n * HD, kf);
    unpack16(v + in_base + (size_t)n * HD, vf);
#pragma unroll
    for (int d = 0; d < DH; ++d) dkf[d] = dvf[d] = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s += qs[t][d] * kf[d]; dp += dos[t][d] * vf[d]; }
      const float p = exp2f(s * scale_log2 - lses[t]);
      const float ds = p * (dp - dsum[t]) * scale;
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        dvf[d] += p * dos[t][d];
        dkf[d] += ds * qs[t][d];
        dqa[t][d] += ds * kf[d];
      }
    }
    pack16(dk + out_base + (size_t)n * HD, dkf);
    pack16(dv + out_base + (size_t)n * HD, dvf);
  }
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      const float s = warp_sum(dqa[t][d]);
      if (lane == 0) red[t][d][warp] = s;
    }
  __syncthreads();
  if (tid < T * DH) {
    const int t = tid / DH, d = tid % DH;
    dq[((size_t)b * T + t) * HD + h * DH + d] = red[t][d][0] + red[t][d][1] + red[t][d][2] + red[t][d][3];
  }
}

// ---------------------------------------------------------------- image -> token attention, backward
// thread = (row n, head h) like the forward.  dout bf16 [B,N,H*DH] (gradient of the attention output, before out_proj).
// Outputs: dqi bf16 [B,N,H*DH] per instance; dkt, dvt fp32 [B,T,H*DH] (zero-initialised by the caller, accumulated with atomics).
template <int T, int DH>
__global__ void __launch_bounds__(256) i2t_attention_bwd_kernel(const __nv_bfloat16* __restrict__ qi, const float* __restrict__ kt,
                                                                const float* __restrict__ vt, const int* __restrict__ src_of,
                                                                const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqi,
                                                                float* __restrict__ dkt, float* __restrict__ dvt, int N, int heads) {
  static_assert(DH == 16, "");
  extern __shared__ float sm[];
  const int b = blockIdx.y;
  const int HD = heads * DH, HS = DH + 1;
  float* ks = sm;                          // [T][heads][HS]
  float* vs = ks + T * heads * HS;
  float* dks = vs + T * heads * HS;        // accumulators
  float* dvs = dks + T * heads * HS;
  const float scale = rsqrtf((float)DH), scale_log2 = scale * 1.4426950408889634f;
  for (int i = threadIdx.x; i < T * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD, o = (t * heads + c / DH) * HS + c % DH;
    ks[o] = kt[(size_t)b * T * HD + i];
    vs[o] = vt[(size_t)b * T * HD + i];
    dks[o] = 0.f;
    dvs[o] = 0.f;
  }
  __syncthreads();
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = gid / heads, h = gid % heads;
  const bool live = n < N;
  float qf[DH], dof[DH], p[T], ds[T];
#pragma unroll
  for (int d = 0; d < DH; ++d) qf[d] = dof[d] = 0.f;
  if (live) {
    unpack16(qi + ((size_t)(src_of ? src_of[b] : b) * N + n) * HD + h * DH, qf);
    unpack16(dout + ((size_t)b * N + n) * HD + h * DH, dof);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) a += qf[d] * ks[(t * heads + h) * HS + d];
    p[t] = a * scale_log2;
    mx = fmaxf(mx, p[t]);
  }
  float l = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) { p[t] = exp2f(p[t] - mx); l += p[t]; }
  float dsum = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    p[t] /= l;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) a += dof[d] * vs[(t * heads + h) * HS + d];
    ds[t] = a;
    dsum += p[t] * a;
  }
  float dq[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) dq[d] = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    ds[t] = live ? p[t] * (ds[t] - dsum) * scale : 0.f;
    if (!live) p[t] = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) dq[d] += ds[t] * ks[(t * heads + h) * HS + d];
  }
  if (live) pack16(dqi + ((size_t)b * N + n) * HD + h * DH, dq);
  // d kt[t,h,:] += ds[t] * qi ; d vt[t,h,:] += p[t] * dout — lanes with equal h (lane % heads) are reduced by shuffles first
  // (heads == 8: lanes l, l^8, l^16, l^24 share a head), then one shared-memory atomic per warp and value
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      float a = ds[t] * qf[d], c = p[t] * dof[d];
      if (heads == 8) {
        a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
        c += __shfl_xor_sync(0xffffffffu, c, 8); c += __shfl_xor_sync(0xffffffffu, c, 16);
        if ((threadIdx.x & 31) < 8) { atomicAdd(&dks[(t * heads + h) * HS + d], a); atomicAdd(&dvs[(t * heads + h) * HS + d], c); }
      } else {
        atomicAdd(&dks[(t * heads + h) * HS + d], a);
        atomicAdd(&dvs[(t * heads + h) * HS + d], c);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < T * HD; i += blockDim.x) {
    const int t = i / HD, c = i % HD, o = (t * heads + c / DH) * HS + c % DH;
    atomicAdd(dkt + (size_t)b * T * HD + i, dks[o]);
    atomicAdd(dvt + (size_t)b * T * HD + i, dvs[o]);
  }
}

// out[n, :] = sum_b x[b, n, :]  (bf16 in, fp32 out): the positional-encoding part of the k/q projection weight gradients
__global__ void batch_sum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int B, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < B; ++b) {
      const uint4 v = reinterpret_cast<const uint4*>(x)[(long long)b * n8 + i];
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(u[j]); a[2 * j] += f.x; a[2 * j + 1] += f.y; }
    }
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(a[0], a[1], a[2], a[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(a[4], a[5], a[6], a[7]);
  }
}

}  // namespace grove
using namespace grove;

extern "C" int grove_decoder_t2i_attention_bwd(const float* q, const void* k, const void* v, const int* src_of, const float* att, const float* datt,
                                               const float* lse, float* dq, void* dk, void* dv, int B, int T, int N, int heads, int dh,
                                               cudaStream_t stream) {
  GROVE_CHECK_ARG(q && k && v && att && datt && lse && dq && dk && dv && B > 0 && N > 0 && heads > 0);
  if (T != 6 || dh != 16) { grove_set_error("t2i attention backward is built for T=6, dh=16 (got T=%d dh=%d)", T, dh); return GROVE_ERR_UNSUPPORTED; }
  t2i_attention_bwd_kernel<6, 16><<<dim3(B, heads), 128, 0, stream>>>(q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, src_of, att, datt, lse, dq,
                                                                     (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, N, heads);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_decoder_i2t_attention_bwd(const void* qi, const float* kt, const float* vt, const int* src_of, const void* dout, void* dqi,
                                               float* dkt, float* dvt, int B, int T, int N, int heads, int dh, cudaStream_t stream) {
  GROVE_CHECK_ARG(qi && kt && vt && dout && dqi && dkt && dvt && B > 0 && N > 0 && heads > 0 && B <= 65535);
  if (T != 6 || dh != 16) { grove_set_error("i2t attention backward is built for T=6, dh=16 (got T=%d dh=%d)", T, dh); return GROVE_ERR_UNSUPPORTED; }
  GROVE_CHECK_ARG(256 % heads == 0);
  const int smem = 4 * T * heads * (dh + 1) * (int)sizeof(float);
  i2t_attention_bwd_kernel<6, 16><<<dim3((N * heads + 255) / 256, B), 256, smem, stream>>>((const __nv_bfloat16*)qi, kt, vt, src_of,
                                                                                           (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqi, dkt, dvt, N,
                                                                                           heads);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_batch_sum_bf16(const void* x, float* out, int B, long long n, cudaStream_t stream) {
  GROVE_CHECK_ARG(x && out && B > 0 && n > 0 && n % 8 == 0);
  long long g = (n / 8 + 255) / 256;
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  batch_sum_bf16_kernel<<<(unsigned)g, 256, 0, stream>>>((const __nv_bfloat16*)x, out, B, n / 8);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}
