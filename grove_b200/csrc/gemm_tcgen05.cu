// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (UMMA 128 x BN x 16, fp32 accumulators double-buffered in TMEM) -> tcgen05.ld epilogue with fused
// bias / GELU / ReLU / tanh-gate / fp32 residual.  One kernel serves
//   * every Linear of the SAM ViT blocks (K4, K9, K10 of SURVEY.md §2.3), the patch-embed GEMM (K1),
//     the neck 1x1 conv (K12), text_hidden_fcs (K13) and the box decoder's image-side projections (K17),
//   * the Conv3d spatio-temporal adapter (K11) and the neck 3x3 conv (K12) as *implicit* GEMMs: the A
//     operand is fetched with a 5-D TMA box at shifted (w,h,t) coordinates, and TMA's out-of-bounds
//     zero fill implements the 'same' zero padding in t, h and w.
//
// out[m, n] = resid[m (mod resid_mod), n] + gate * act( sum_k A[m, k] * W[n, k] + bias[n] )
#include <cuda.h>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "grove_b200.h"

namespace grove {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue

struct GemmParams {
  int M, N;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int conv;            // 0: A is a plain [M,K] matrix; 1: implicit-GEMM taps over [V,T,G,G,C]
  int G, rows_per_tile, tiles_per_frame, T, kt, kc_blocks;
  const float* bias;   // [N] or null
  const float* resid;  // fp32 [*,N] or null
  int resid_mod;       // >0: residual row = m % resid_mod (abs-pos embedding broadcast over frames)
  const float* gate_alpha;  // non-null: gate = tanh(*gate_alpha)   (adapter, image_encoder.py:54)
  int act;             // 0 none, 1 exact GELU, 2 ReLU
  void* out;           // [M,N] fp32 or bf16
  int out_f32;
  __nv_bfloat16* out2; // optional extra bf16 copy of the output (feeds the next tensor-core op)
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two: 256 or 512)
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
        int f = 0, h0 = 0;
        if (p.conv) {
          f = m_blk / p.tiles_per_frame;
          h0 = (m_blk % p.tiles_per_frame) * p.rows_per_tile;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (!p.conv) {
            tma_load_2d(sa, &tmap_a, full_bar(s), kb * BK, m_blk * BM);
          } else {
            const int tap = kb / p.kc_blocks, c0 = (kb % p.kc_blocks) * BK;
            int dt = 0, dh, dw;
            if (p.kt == 3) { dt = tap / 9 - 1; dh = (tap / 3) % 3 - 1; dw = tap % 3 - 1; }
            else           { dh = tap / 3 - 1; dw = tap % 3 - 1; }
            tma_load_5d(sa, &tmap_a, full_bar(s), c0, dw, h0 + dh, (f % p.T) + dt, f / p.T);
          }
          tma_load_2d(sb, &tmap_b, full_bar(s), kb * BK, n_blk * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    uint32_t it = 0, tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const uint32_t acc = tile_it & 1u, acc_ph = (tile_it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_k_blocks; ++kb, ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = (it / Cfg::kStages) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_base + s * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32);
            const uint64_t db = umma_desc_sw128(sb + k * 32);
            tc_mma_f16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          tc_commit(empty_bar(s));                                   // smem slot free once these MMAs retire
          if (kb == p.num_k_blocks - 1) tc_commit(tfull_bar(acc));   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const float gate = p.gate_alpha ? tanhf(__ldg(p.gate_alpha)) : 1.0f;
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_it) {
      const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
      const uint32_t acc = tile_it & 1u, acc_ph = (tile_it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const size_t out_off = (size_t)row * p.N;
      const float* resid_row = nullptr;
      if (p.resid) resid_row = p.resid + (size_t)(p.resid_mod > 0 ? row % p.resid_mod : row) * p.N;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + ch * 32 + ((uint32_t)(quad * 32) << 16), r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + ch * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (p.gate_alpha) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= gate;
        }
        if (row_ok) {
          if (resid_row) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(resid_row + col0 + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + out_off + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_off + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                            pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
          }
          if (p.out2) {
            __nv_bfloat16* o = p.out2 + out_off + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                            pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------ host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 tensor, innermost dimension contiguous; dims/box listed innermost first.  128-byte swizzle, zero OOB fill.
static int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { grove_set_error("cuTensorMapEncodeTiled entry point not available"); return GROVE_ERR_CUDA; }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { grove_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank); return GROVE_ERR_CUDA; }
  return GROVE_OK;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = kNumSMs;
  }
  return g_num_sms;
}

template <int BN>
static int launch_gemm(const GemmParams& p, const CUtensorMap& ta, const CUtensorMap& tb, int max_ctas, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { grove_set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e)); return GROVE_ERR_CUDA; }
    attr_set = true;
  }
  int grid = p.num_m_blocks * p.num_n_blocks;
  int cap = max_ctas > 0 ? max_ctas : num_sms();
  if (grid > cap) grid = cap;
  gemm_bf16_tcgen05_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(p, ta, tb);
  grove_count_launch();
  GROVE_CHECK_LAUNCH();
  return GROVE_OK;
}

}  // namespace grove

using namespace grove;

static int fill_epilogue(GemmParams& p, const grove_gemm_epilogue* e, int M, int N) {
  p.bias = e ? e->bias : nullptr;
  p.resid = e ? e->resid : nullptr;
  p.resid_mod = e ? e->resid_row_mod : 0;
  p.gate_alpha = e ? e->gate_alpha : nullptr;
  p.act = e ? e->act : 0;
  p.out_f32 = e ? e->out_f32 : 0;
  p.out2 = e ? reinterpret_cast<__nv_bfloat16*>(e->out2_bf16) : nullptr;
  GROVE_CHECK_ARG(p.act >= 0 && p.act <= 2);
  GROVE_CHECK_ARG(((uintptr_t)p.bias & 15) == 0 && ((uintptr_t)p.resid & 15) == 0 && ((uintptr_t)p.out2 & 15) == 0);
  (void)M; (void)N;
  return GROVE_OK;
}

extern "C" int grove_gemm_bf16(const void* A, const void* W, void* out, int M, int N, int K, const grove_gemm_epilogue* epi,
                               cudaStream_t stream) {
  GROVE_CHECK_ARG(A && W && out && M > 0 && N > 0 && K > 0);
  GROVE_CHECK_ARG(N % 128 == 0 && K % 8 == 0);
  GROVE_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)out & 15) == 0);
  const int BN = (N % 256 == 0) ? 256 : 128;
  GemmParams p{};
  p.M = M; p.N = N;
  p.num_m_blocks = (M + BM - 1) / BM;
  p.num_n_blocks = N / BN;
  p.num_k_blocks = (K + BK - 1) / BK;
  p.conv = 0;
  p.out = out;
  int rc = fill_epilogue(p, epi, M, N);
  if (rc) return rc;
  CUtensorMap ta, tb;
  uint64_t da[2] = {(uint64_t)K, (uint64_t)M};
  uint32_t ba[2] = {BK, BM};
  if ((rc = make_tmap_bf16(&ta, A, 2, da, ba))) return rc;
  uint64_t db[2] = {(uint64_t)K, (uint64_t)N};
  uint32_t bb[2] = {BK, (uint32_t)BN};
  if ((rc = make_tmap_bf16(&tb, W, 2, db, bb))) return rc;
  const int cap = epi ? epi->max_ctas : 0;
  return BN == 256 ? launch_gemm<256>(p, ta, tb, cap, stream) : launch_gemm<128>(p, ta, tb, cap, stream);
}

extern "C" int grove_conv_gemm_bf16(const void* X, const void* Wp, void* out, int V, int T, int G, int C, int N, int kt,
                                    const grove_gemm_epilogue* epi, cudaStream_t stream) {
  GROVE_CHECK_ARG(X && Wp && out && V > 0 && T > 0 && G > 0 && C > 0 && N > 0);
  GROVE_CHECK_ARG(kt == 1 || kt == 3);
  GROVE_CHECK_ARG(G <= 128 && 128 % G == 0 && (G * G) % 128 == 0);  // a 128-token M tile is whole grid rows of one frame
  GROVE_CHECK_ARG(C % 64 == 0 && N % 128 == 0);
  GROVE_CHECK_ARG(((uintptr_t)X & 15) == 0 && ((uintptr_t)Wp & 15) == 0 && ((uintptr_t)out & 15) == 0);
  const int BN = (N % 256 == 0) ? 256 : 128;
  const int ntaps = kt * 9;
  GemmParams p{};
  p.M = V * T * G * G; p.N = N;
  p.num_m_blocks = p.M / BM;
  p.num_n_blocks = N / BN;
  p.kc_blocks = C / BK;
  p.num_k_blocks = ntaps * p.kc_blocks;
  p.conv = 1; p.G = G; p.rows_per_tile = BM / G; p.tiles_per_frame = G * G / BM; p.T = T; p.kt = kt;
  p.out = out;
  int rc = fill_epilogue(p, epi, p.M, N);
  if (rc) return rc;
  CUtensorMap ta, tb;
  uint64_t da[5] = {(uint64_t)C, (uint64_t)G, (uint64_t)G, (uint64_t)T, (uint64_t)V};
  uint32_t ba[5] = {BK, (uint32_t)G, (uint32_t)(BM / G), 1, 1};
  if ((rc = make_tmap_bf16(&ta, X, 5, da, ba))) return rc;
  uint64_t db[2] = {(uint64_t)ntaps * C, (uint64_t)N};
  uint32_t bb[2] = {BK, (uint32_t)BN};
  if ((rc = make_tmap_bf16(&tb, Wp, 2, db, bb))) return rc;
  const int cap = epi ? epi->max_ctas : 0;
  return BN == 256 ? launch_gemm<256>(p, ta, tb, cap, stream) : launch_gemm<128>(p, ta, tb, cap, stream);
}
