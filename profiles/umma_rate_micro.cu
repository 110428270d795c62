// Tensor-pipe rate of the global-attention kernel's UMMA instruction mix, in isolation and next to softmax-like TMEM traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I grove_b200/csrc -I include -o /tmp/umma_rate profiles/umma_rate_micro.cu && /tmp/umma_rate
// One CTA per SM.  Warp 0 issues, per "block", 4 x (128x128x16, A from tensor memory: Q.K^T) + 8 x (128x64x16, A from tensor memory, B
// MN-major: P.V) -- nominal 4 x 64 + 8 x 32 = 512 clk -- and commits; warps 2-9 optionally run the softmax's tensor-memory traffic in a loop
// (two tcgen05.ld.x32 of S, two tcgen05.st.x16 of P per iteration and thread).  Prints clk per block for: MMA alone (TS), MMA alone with
// Q from shared memory (SS), MMA + traffic.
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
#include "tmem_ldst.cuh"
using namespace grove;

template <int MODE>   // 0: TS alone, 1: SS Q.K^T alone, 2: TS + TMEM traffic from 8 warps
__global__ void __launch_bounds__(320, 1) umma_rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = s0, sK = s0 + 16384, sV = s0 + 32768, bar0 = s0 + 49152, slot = bar0 + 64;   // bar0, bar0 + 8: two commit barriers
  volatile int* stop = reinterpret_cast<volatile int*>(smem_raw + (s0 - smem_u32(smem_raw)) + 49152 + 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t a = s0 + threadIdx.x * 16; a < s0 + 49152; a += 320 * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0x3c003c00u) : "memory");
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); *stop = 0; fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tb;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tb) : "r"(slot));
  const uint32_t tS = tb, tO = tb + 256, tQ = tb + 336, tP = tb + 384;
  if (warp == 0) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128), idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);
    long long t0 = 0;
    for (int it = 0; it < iters + 8; ++it) {
      if (it == 8) t0 = clock64();
      if (it >= 2) mbar_wait(bar0 + 8 * (it & 1), ((it - 2) >> 1) & 1u);      // two blocks in flight (S is double-buffered in the kernel)
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = (it & 1) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (MODE == 1) tc_mma_f16(tS + sb, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(sK + k * 32), idesc_s, k != 0);
          else tc_mma_f16_ts(tS + sb, tQ + k * 8, umma_desc_sw128(sK + k * 32), idesc_s, k != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) tc_mma_f16_ts(tO, tP + (it & 1) * 64 + kk * 8, umma_desc_sw128(sV + kk * 2048), idesc_o, 1);
        tc_commit(bar0 + 8 * (it & 1));
      }
      __syncwarp();
    }
    mbar_wait(bar0 + 8 * ((iters + 6) & 1), ((iters + 6) >> 1) & 1u);
    mbar_wait(bar0 + 8 * ((iters + 7) & 1), ((iters + 7) >> 1) & 1u);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = (t1 - t0) / iters;
    *stop = 1;
  } else if (warp >= 2 && MODE == 2) {
    const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
    const int hs = (warp - 2) >> 2;
    uint32_t acc = 0;
    while (!*stop) {
      uint32_t r0[32], r1[32], p0[16], p1[16];
      tmem_ld_x32(tS + hs * 32 + tlane, r0);
      tmem_ld_x32(tS + (hs + 2) * 32 + tlane, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) { p0[j] = r0[2 * j] ^ r0[2 * j + 1]; p1[j] = r1[2 * j] ^ r1[2 * j + 1]; acc += p0[j] + p1[j]; }
      tmem_st_x16(tP + hs * 16 + tlane, p0);
      tmem_st_x16(tP + (hs + 2) * 16 + tlane, p1);
      tmem_st_wait();
      __nanosleep(200);                              // the softmax spends ~1300 clk per block between its TMEM accesses
    }
    if (acc == 0x12345678u) out[1000] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

template <int MODE>
static void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 2048 * sizeof(long long));
  const int smem = 49152 + 1024 + 512;
  cudaFuncSetAttribute(umma_rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_rate_kernel<MODE><<<148, 320, smem>>>(d, 2000);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mn = h[0], mx = h[0], sum = 0;
  for (int i = 0; i < 148; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
  printf("%-44s %s  clk per block (4 QK^T + 8 PV UMMAs, nominal 512): mean %lld  min %lld  max %lld\n", name, cudaGetErrorString(e), sum / 148, mn, mx);
  cudaFree(d);
}

int main() {
  run<0>("UMMA alone, Q from tensor memory (TS)");
  run<1>("UMMA alone, Q from shared memory (SS)");
  run<2>("UMMA (TS) + softmax-like TMEM ld/st traffic");
  return 0;
}
