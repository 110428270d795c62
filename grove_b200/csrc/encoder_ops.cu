// HBM-bound encoder helpers: patch im2col, LayerNorm, casts, layout transposes.  All vectorised (16 B per
// thread access), coalesced along the channel dimension; LayerNorm keeps the row in registers (one warp per row).
#include "common.cuh"
#include "grove_b200.h"

namespace grove {

// ---------------------------------------------------------------- im2col for the 16x16/stride-16 patch embed
// One thread moves one 16-pixel patch row (32 B): consecutive threads take consecutive gx, so a warp reads one
// contiguous 1 KB image-row segment.  Output row = patch, column k = c*256 + py*16 + px.
__global__ void im2col_patch16_kernel(const __nv_bfloat16* __restrict__ img, __nv_bfloat16* __restrict__ out, int V, int T, int H, int W) {
  const int GX = W / 16, GY = H / 16;
  const long long total = (long long)V * T * 3 * H * GX;  // (f, c, y, gx)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % GX);
    long long r = i / GX;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % 3);
    const int f = (int)(r / 3);
    const int v = f / T, t = f % T;
    const __nv_bfloat16* src = img + ((((long long)v * 3 + c) * T + t) * H + y) * W + gx * 16;
    const int gy = y / 16, py = y % 16;
    __nv_bfloat16* dst = out + (((long long)f * GY + gy) * GX + gx) * 768 + c * 256 + py * 16;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(src) + 1);
    reinterpret_cast<uint4*>(dst)[0] = a;
    reinterpret_cast<uint4*>(dst)[1] = b;
  }
}

// ---------------------------------------------------------------- LayerNorm, one warp per row
template <bool OUT_F32, bool IN_BF16 = false>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, void* __restrict__ y, int rows, int D, float eps) {
  constexpr int MAXV = 10;  // D <= 1280: 10 float4 per lane
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nv = D / 128;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      if (IN_BF16) {       // bf16 residual stream: 4 values = 8 bytes per lane
        const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)warp * D)[i * 32 + lane];
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        v[i] = make_float4(a.x, a.y, b.x, b.y);
      } else {
        v[i] = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + (size_t)warp * D)[i * 32 + lane];
      }
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (OUT_F32) {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)warp * D)[i * 32 + lane] = make_float4(o0, o1, o2, o3);
      } else {
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + (size_t)warp * D)[i * 32 + lane] =
            make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
      }
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
    reinterpret_cast<uint4*>(y)[i] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

// [F, R, Ccols] -> [F, Ccols, R] for 16-bit elements through a padded 32x32 smem tile
__global__ void transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int R, int C) {
  __shared__ uint16_t tile[32][34];
  const size_t base = (size_t)blockIdx.z * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[base + (size_t)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[base + (size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

}  // namespace grove
using namespace grove;

static inline int grid_for(long long n, int block, int per_sm = 8) {
  long long g = (n + block - 1) / block;
  long long cap = (long long)kNumSMs * per_sm;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" int grove_im2col_patch16(const void* images, void* patches, int V, int T, int H, int W, cudaStream_t stream) {
  GROVE_CHECK_ARG(images && patches && V > 0 && T > 0 && H % 16 == 0 && W % 16 == 0);
  const long long total = (long long)V * T * 3 * H * (W / 16);
  im2col_patch16_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(images),
                                                                   reinterpret_cast<__nv_bfloat16*>(patches), V, T, H, W);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_layernorm(const float* x, const float* gamma, const float* beta, void* y, int y_f32, int rows, int D, float eps,
                               cudaStream_t stream) {
  GROVE_CHECK_ARG(x && gamma && beta && y && rows > 0);
  GROVE_CHECK_ARG(D % 128 == 0 && D <= 1280);
  const int blocks = (rows + 7) / 8;
  if (y_f32) layernorm_kernel<true><<<blocks, 256, 0, stream>>>(x, gamma, beta, y, rows, D, eps);
  else       layernorm_kernel<false><<<blocks, 256, 0, stream>>>(x, gamma, beta, y, rows, D, eps);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_layernorm_bf16in(const void* x, const float* gamma, const float* beta, void* y, int y_f32, int rows, int D, float eps,
                                      cudaStream_t stream) {
  GROVE_CHECK_ARG(x && gamma && beta && y && rows > 0);
  GROVE_CHECK_ARG(D % 128 == 0 && D <= 1280);
  const int blocks = (rows + 7) / 8;
  if (y_f32) layernorm_kernel<true, true><<<blocks, 256, 0, stream>>>(x, gamma, beta, y, rows, D, eps);
  else       layernorm_kernel<false, true><<<blocks, 256, 0, stream>>>(x, gamma, beta, y, rows, D, eps);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_cast_f32_bf16(const float* x, void* y, long long n, cudaStream_t stream) {
  GROVE_CHECK_ARG(x && y && n > 0 && n % 8 == 0);
  cast_f32_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n / 8);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

static int transpose16(const void* in, void* out, int F, int R, int C, cudaStream_t stream) {
  GROVE_CHECK_ARG(in && out && F > 0 && R > 0 && C > 0 && F <= 65535);
  dim3 grid((C + 31) / 32, (R + 31) / 32, F), block(32, 8);
  transpose16_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<const uint16_t*>(in), reinterpret_cast<uint16_t*>(out), R, C);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}
extern "C" int grove_tokens_to_nchw_bf16(const void* tok, void* nchw, int F, int N, int C, cudaStream_t stream) {
  return transpose16(tok, nchw, F, N, C, stream);
}
extern "C" int grove_nchw_to_tokens_bf16(const void* nchw, void* tok, int F, int N, int C, cudaStream_t stream) {
  return transpose16(nchw, tok, F, C, N, stream);
}

// ------------------------------------------------------------------ AdaptiveAvgPooling3D over token-major video features
// (model/llava/model/multimodal_encoder/pooling.py:6-25: '(b t) (h w) c -> b c t h w', nn.AdaptiveAvgPool3d((OT, OH, OW)),
// 'b c t h w -> b (t h w) c').  The two rearranges are folded into the indexing: one thread owns 8 (bf16) / 4 (fp32) channels of one
// output token, sums its window in fp32 (window = [floor(i*I/O), ceil((i+1)*I/O)) per axis, as ATen's adaptive pooling) and divides by
// the window size.  HBM-bound: every input element is read once or twice (overlapping windows), 16-byte accesses along c.
namespace grove {
template <bool F32>
__global__ void __launch_bounds__(256) adaptive_avgpool3d_tokens_kernel(const void* __restrict__ x, void* __restrict__ out, int B, int T, int H, int W,
                                                                        int C, int OT, int OH, int OW) {
  constexpr int VEC = F32 ? 4 : 8;
  const int cv = C / VEC;
  const long long total = (long long)B * OT * OH * OW * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * VEC;
    long long r = i / cv;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH); r /= OH;
    const int ot = (int)(r % OT);
    const int b = (int)(r / OT);
    const int t0 = (ot * T) / OT, t1 = ((ot + 1) * T + OT - 1) / OT;
    const int h0 = (oh * H) / OH, h1 = ((oh + 1) * H + OH - 1) / OH;
    const int w0 = (ow * W) / OW, w1 = ((ow + 1) * W + OW - 1) / OW;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    for (int t = t0; t < t1; ++t)
      for (int h = h0; h < h1; ++h)
        for (int w = w0; w < w1; ++w) {
          const size_t off = (((size_t)(b * T + t) * H + h) * W + w) * C + c0;
          if (F32) {
            const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + off);
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
          } else {
            const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + off);
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16(u[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
          }
        }
    const float inv = 1.0f / (float)((t1 - t0) * (h1 - h0) * (w1 - w0));
    const size_t o = ((((size_t)b * OT + ot) * OH + oh) * OW + ow) * C + c0;
    if (F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + o) =
          make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv), pack_bf16(acc[4] * inv, acc[5] * inv),
                     pack_bf16(acc[6] * inv, acc[7] * inv));
    }
  }
}
}  // namespace grove

extern "C" int grove_adaptive_avgpool3d_tokens(const void* x, void* out, int is_f32, int B, int T, int H, int W, int C, int OT, int OH, int OW,
                                               cudaStream_t stream) {
  GROVE_CHECK_ARG(x && out && B > 0 && T > 0 && H > 0 && W > 0 && OT > 0 && OH > 0 && OW > 0);
  GROVE_CHECK_ARG(OT <= T && OH <= H && OW <= W && C % 8 == 0);
  GROVE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0);
  const long long total = (long long)B * OT * OH * OW * (C / (is_f32 ? 4 : 8));
  const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  if (is_f32) grove::adaptive_avgpool3d_tokens_kernel<true><<<grid, 256, 0, stream>>>(x, out, B, T, H, W, C, OT, OH, OW);
  else grove::adaptive_avgpool3d_tokens_kernel<false><<<grid, 256, 0, stream>>>(x, out, B, T, H, W, C, OT, OH, OW);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

