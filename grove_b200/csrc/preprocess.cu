// Frame pre-processing fused into the patch embed's operand: decoded RGB frames (uint8, HWC) ->
// ResizeLongestSide.apply_image (model/SAM/utils/transforms.py:27-34: PIL bilinear, restated from Pillow's Resample.c and
// bit-exact with it) -> grounding_enc_processor (HowTo100M.py:168-178: (x - mean) / std in fp32, zero pad right / bottom) ->
// .bfloat16() (train.py:751-753) -> 16x16 patch rows of the PatchEmbed GEMM (image_encoder.py:484-491).
// The normalised fp32 / bf16 image never exists in HBM: the vertical resampling pass writes the GEMM's A operand directly
// (6 MB per 1024^2 frame of traffic and two host-side passes removed).  HBM-bound byte work: one thread per 16-pixel patch row.
#include "common.cuh"
#include "grove_b200.h"

namespace grove {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow's 8-bit resampling fixed point

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: in [F, h, w_in, 3] -> out [F, h, w_out, 3]; bounds [w_out, 2] = (xmin, count), coeffs [w_out, ksize]
__global__ void resize_rows_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int* __restrict__ bounds,
                                      const int* __restrict__ coeffs, int ksize, long long rows, int w_in, int w_out) {
  const long long total = rows * w_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % w_out);
    const long long row = i / w_out;
    const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
    const uint8_t* src = in + (row * w_in + xmin) * 3;
    const int* k = coeffs + (size_t)xx * ksize;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < cnt; ++x) {
      const int kk = __ldg(k + x);
      s0 += (int)src[3 * x] * kk;
      s1 += (int)src[3 * x + 1] * kk;
      s2 += (int)src[3 * x + 2] * kk;
    }
    uint8_t* dst = out + (row * w_out + xx) * 3;
    dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
  }
}

// vertical pass + normalise + pad + patchify: in [F, h_in, w, 3] u8 -> patches [(F * G * G), 768] bf16, k = c*256 + py*16 + px.
// thread = (frame, output row y < img, patch column gx).  Rows y >= h_out and columns x >= w are the zero padding.
__global__ void frames_to_patches_u8_kernel(const uint8_t* __restrict__ in, const int* __restrict__ vbounds, const int* __restrict__ vcoeffs,
                                            int vksize, __nv_bfloat16* __restrict__ patches, int F, int h_in, int w, int h_out, int img,
                                            float m0, float m1, float m2, float s0, float s1, float s2) {
  const int G = img / 16;
  const long long total = (long long)F * img * G;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % G);
    const int y = (int)((i / G) % img);
    const int f = (int)(i / ((long long)G * img));
    __nv_bfloat16* dst = patches + (((size_t)f * G + y / 16) * G + gx) * 768 + (y % 16) * 16;
    uint32_t o[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) o[c][j] = 0u;
    if (y < h_out) {
      const int ymin = __ldg(vbounds + 2 * y), cnt = __ldg(vbounds + 2 * y + 1);
      const int* k = vcoeffs + (size_t)y * vksize;
      const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
#pragma unroll 1
      for (int px = 0; px < 16; px += 2) {
        float v[3][2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int x = gx * 16 + px + e;
          if (x < w) {
            int acc[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
            const uint8_t* src = in + (((size_t)f * h_in + ymin) * w + x) * 3;
            for (int r = 0; r < cnt; ++r) {
              const int kk = __ldg(k + r);
              acc[0] += (int)src[0] * kk; acc[1] += (int)src[1] * kk; acc[2] += (int)src[2] * kk;
              src += (size_t)w * 3;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c][e] = __fdiv_rn(__fsub_rn((float)clip8(acc[c]), mean[c]), sd[c]);
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c][e] = 0.f;
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c][px / 2] = pack_bf16(v[c][0], v[c][1]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint4* d = reinterpret_cast<uint4*>(dst + c * 256);
      d[0] = make_uint4(o[c][0], o[c][1], o[c][2], o[c][3]);
      d[1] = make_uint4(o[c][4], o[c][5], o[c][6], o[c][7]);
    }
  }
}

}  // namespace grove
using namespace grove;

extern "C" int grove_resize_rows_u8(const uint8_t* in, uint8_t* out, const int* bounds, const int* coeffs, int ksize, long long rows, int w_in,
                                    int w_out, cudaStream_t stream) {
  GROVE_CHECK_ARG(in && out && bounds && coeffs && ksize > 0 && rows > 0 && w_in > 0 && w_out > 0);
  long long g = (rows * w_out + 255) / 256;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  resize_rows_u8_kernel<<<(unsigned)g, 256, 0, stream>>>(in, out, bounds, coeffs, ksize, rows, w_in, w_out);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

extern "C" int grove_frames_to_patches_u8(const uint8_t* in, const int* vbounds, const int* vcoeffs, int vksize, void* patches, int F, int h_in,
                                          int w, int h_out, int img, const float* mean3, const float* std3, cudaStream_t stream) {
  GROVE_CHECK_ARG(in && vbounds && vcoeffs && patches && mean3 && std3 && vksize > 0 && F > 0 && h_in > 0 && w > 0);
  GROVE_CHECK_ARG(img % 16 == 0 && h_out > 0 && h_out <= img && w <= img);
  long long g = ((long long)F * img * (img / 16) + 127) / 128;
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  frames_to_patches_u8_kernel<<<(unsigned)g, 128, 0, stream>>>(in, vbounds, vcoeffs, vksize, (__nv_bfloat16*)patches, F, h_in, w, h_out, img,
                                                               mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}
