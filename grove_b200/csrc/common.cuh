// Shared device/host helpers for the grove_b200 CUDA library (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define GROVE_OK 0
#define GROVE_ERR_ARG 1
#define GROVE_ERR_CUDA 2
#define GROVE_ERR_UNSUPPORTED 3

#define GROVE_CHECK_ARG(cond)                                            \
  do {                                                                   \
    if (!(cond)) {                                                       \
      grove_set_error("%s:%d: argument check failed: %s", __FILE__, __LINE__, #cond); \
      return GROVE_ERR_ARG;                                              \
    }                                                                    \
  } while (0)

#define GROVE_CHECK_LAUNCH()                                             \
  do {                                                                   \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) {                                            \
      grove_set_error("%s:%d: CUDA launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return GROVE_ERR_CUDA;                                             \
    }                                                                    \
  } while (0)

void grove_set_error(const char* fmt, ...);
void grove_count_launch(int n = 1);

// Function attributes (max dynamic shared memory) and the SM count belong to a DEVICE, not to the process: one flag bit per device
// ordinal, set atomically, so the first launch on every GPU of a multi-device process configures that GPU's copy of the kernel.
struct GrovePerDeviceOnce {
  unsigned long long bits = 0;
  bool first_time() {   // true exactly once per device (racing callers may both see true: setting an attribute twice is harmless)
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    const unsigned long long prev = __atomic_fetch_or(&bits, bit, __ATOMIC_ACQ_REL);
    return !(prev & bit);
  }
  void reset_current() {
    int dev = 0;
    cudaGetDevice(&dev);
    __atomic_fetch_and(&bits, ~(1ull << (dev & 63)), __ATOMIC_ACQ_REL);
  }
};

namespace grove {

// GROVE_PDL=1 launches the GEMM and the two forward attention kernels with programmatic dependent launch (their prologue then overlaps the
// predecessor's last wave).  Same-box A/B: kernel-by-kernel launches 13.5 -> 13.2 ms per step (+2.3 %), but no effect inside the captured
// whole-step graph (the default path), so it is opt-in.
inline bool grove_pdl_enabled() {
  static const bool v = []() { const char* e = getenv("GROVE_PDL"); return e && e[0] == '1'; }();
  return v;
}
// kernel<<<grid, block, smem, st>>>(args...) with the PDL attribute (the kernel must call pdl_wait() before touching global memory)
template <class... KArgs, class... Args>
inline cudaError_t grove_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = a;
  cfg.numAttrs = grove_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kNumSMs = 148;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

// exact (erf) GELU — nn.GELU() default, model/SAM/modeling/common.py:18
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU for fused epilogues: x * Phi(x) with Phi(x) = 1 / (1 + exp(-u)), u = x (a + b x^2 + c x^4) -- a minimax fit of
// atanh(erf(x / sqrt 2)) (scratch fit: max |error| of the GELU value 2.5e-5 over all x, 100x below the bf16 rounding of the
// output it feeds).  9 instructions, 2 MUFU (ex2, rcp): the erf form (rcp + ex2 + 8 FMA + sign handling) made the fc1 epilogue
// issue-bound.  x^2 is clamped at 50 where the fit has saturated (|x| > 7: Phi = 0 or 1 to 1e-11).
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 50.0f);
  float q = fmaf(x2, 0.0010142630f, -0.10677572f);      // -log2(e) * (c x^2 + b)
  q = fmaf(q, x2, -2.3011214f);                         // ... * x^2 - log2(e) * a
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * x));
  return __fdividef(x, 1.0f + e);
}

// the same on a pair with Blackwell's packed fp32 instructions (FMUL2 / FFMA2 / FADD2): 13 issue slots per two values instead of 18
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  float2 x2 = __fmul2_rn(x, x);
  x2.x = fminf(x2.x, 50.0f);
  x2.y = fminf(x2.y, 50.0f);
  float2 q = __ffma2_rn(x2, make_float2(0.0010142630f, 0.0010142630f), make_float2(-0.10677572f, -0.10677572f));
  q = __ffma2_rn(q, x2, make_float2(-2.3011214f, -2.3011214f));
  const float2 t = __fmul2_rn(q, x);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  return make_float2(__fdividef(x.x, d.x), __fdividef(x.y, d.y));
}

// d/dx of the exact GELU: Phi(x) + x * phi(x).  gelu_grad is the erf form (token-side kernels); gelu_grad_fast serves the GEMM epilogue of
// the fc2 input gradient (M x 3072 elements per block): Phi from the same sigmoid-polynomial fit as gelu_fast (max |error| 5e-5, far
// below the bf16 rounding of the gradient it scales), phi with one ex2 -- 3 MUFU + 9 FMA-pipe instructions instead of erff + expf.
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float xx = x * x;
  const float x2 = fminf(xx, 50.0f);
  float q = fmaf(x2, 0.0010142630f, -0.10677572f);
  q = fmaf(q, x2, -2.3011214f);
  float e, g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(-0.72134752044f * xx));      // exp(-x^2 / 2)
  return __fdividef(1.0f, 1.0f + e) + x * 0.3989422804014327f * g;
}

// d/dx of gelu_fast on a pair (packed fp32): with s = sigmoid(u), u = x (a + b x^2 + c x^4), the derivative is s + x s (1 - s) u'(x) -- the
// exact derivative of the function the forward epilogue applies, 2 MUFU (ex2, rcp) and ~8.5 issue slots per value.  The three-MUFU
// gelu_grad_fast made the fc2 input-gradient GEMM epilogue MUFU-bound: 3 x 128 x 256 per tile at 16 / clk = the tile's whole MMA time.
__device__ __forceinline__ float2 gelu_grad_fast2(float2 x) {
  const float2 xx = __fmul2_rn(x, x);
  const float2 x2 = make_float2(fminf(xx.x, 50.0f), fminf(xx.y, 50.0f));
  float2 q = __ffma2_rn(x2, make_float2(0.0010142630f, 0.0010142630f), make_float2(-0.10677572f, -0.10677572f));
  q = __ffma2_rn(q, x2, make_float2(-2.3011214f, -2.3011214f));
  const float2 t = __fmul2_rn(q, x);                    // -log2(e) u
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  float2 s;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s.y) : "f"(d.y));
  // u'(x) = -ln 2 * (A + 3 B x^2 + 5 C x^4) with the log2-scaled coefficients above
  float2 w = __ffma2_rn(x2, make_float2(-0.0035151677f, -0.0035151677f), make_float2(0.22203387f, 0.22203387f));
  w = __ffma2_rn(w, x2, make_float2(1.5950158f, 1.5950158f));
  const float2 wx = __fmul2_rn(w, x);
  const float2 sm = __ffma2_rn(make_float2(-s.x, -s.y), s, s);      // s (1 - s)
  return __ffma2_rn(wx, sm, s);
}

// 2^x for x <= 0 on the FMA pipe (Cody-Waite split with the 1.5 * 2^23 rounding constant, degree-3 minimax of 2^f on [-0.5, 0.5],
// exponent patched in with an integer add): relative error 2.8e-4, well inside the bf16 rounding of P.  The softmax is bound by
// the 16 ex2/clk/SM MUFU rate, so one exponential in four is computed here instead (the FlashAttention-4 trick).
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.0565415754f, 0.242068237f);
  p = fmaf(p, f, 0.692983806f);
  p = fmaf(p, f, 0.999953968f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// The same polynomial on two values at once with Blackwell's packed fp32 pipe (FFMA2 / FADD2): 10 issue slots per PAIR.  The softmax warps
// are issue-bound (one instruction per clock per SM sub-partition), so everything per-element in the loops below is written on float2:
// scale*S + rel_w is one FFMA2 per two scores, the running maximum one FMNMX3, the offset add and the row-sum one FADD2 each.
__device__ __forceinline__ float2 ex2_fma2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 kMagic = make_float2(12582912.0f, 12582912.0f);
  const float2 t = __fadd2_rn(x, kMagic);
  const float2 u = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(u, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(f, make_float2(0.0565415754f, 0.0565415754f), make_float2(0.242068237f, 0.242068237f));
  p = __ffma2_rn(p, f, make_float2(0.692983806f, 0.692983806f));
  p = __ffma2_rn(p, f, make_float2(0.999953968f, 0.999953968f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start (block
// scheduling, barrier init, TMEM allocation, descriptor prefetch) while its predecessor in the stream / graph is still draining its last
// wave; pdl_wait() blocks until the predecessor grid has completed and its memory is visible, so it must precede EVERY global-memory access
// of the kernel.  pdl_launch_dependents() lets the successor's blocks be scheduled as soon as this grid's blocks have all started.
// Both are no-ops for kernels launched without the attribute / without a programmatic successor.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// One leader lane of a fully converged warp (elect.sync).  tcgen05.mma / tcgen05.commit / TMA are uniform-datapath instructions:
// issued under `if (lane == 0)` ptxas wraps EVERY one of them in a serialising ELECT ... BRA.U.ANY loop with R2UR moves (~60-100 clk
// per instruction, measured), which made the single MMA-issuing warp the bottleneck of the attention kernels; under an elect.sync
// predicate the instruction is issued directly.  The election is deterministic for a given member mask, so the same lane issues
// the MMAs and the commits that track them.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Wait with a watchdog: a protocol bug must trap (launch error reported to the caller) instead of
// hanging the GPU.  try_wait suspends in hardware, so the poll loop is cheap.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef GROVE_MBAR_SPIN   // A/B variant without the hardware-sleep hint: measured identical, kept for re-measurement
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins == 0x10000000u) __trap();
  }
#else
  uint32_t done = 0, spins = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)   // suspend-time hint: park the thread in hardware instead of spinning
        : "memory");
    if (done) break;
    if ((++spins & 0xfffu) == 0) {
      uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();  // 4 s
    }
  }
#endif
}

// Wait for TWO barriers with both polls in flight at once.  An mbarrier.try_wait costs ~90 clk even when the phase has already
// completed; a single MMA-issuing warp that waits for "operand landed" and "accumulator drained" one after the other pays that
// latency twice per tile, which was the pacing term of the attention kernels (measured with in-kernel clock probes).
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
  uint32_t da, db;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "selp.u32 %1, 1, 0, q;\n\t}"
      : "=r"(da), "=r"(db)
      : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b)
      : "memory");
  if (!da) mbar_wait(bar_a, par_a);
  if (!db) mbar_wait(bar_b, par_b);
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 in, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}


// ------------------------------------------------------------------ 2-CTA (cta_group::2) variants and cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA of the pair
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank)
      : "memory");
}
// The same without release semantics, for hand-offs that publish NO memory: an epilogue warp telling the MMA warp of the pair's leader that
// its tcgen05.ld reads of an accumulator have completed (tcgen05.wait::ld + tcgen05.fence::before_thread_sync order those).  The .release
// form compiles to MEMBAR.ALL.CTA + ERRBAR, which waits for every global store and prefetch load the warp has in flight -- ~12 % of the
// epilogue warps' stall samples on the K = 768 residual GEMM (profiles/r2_gemm_proj_summary.txt).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cta(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit -> arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_cg2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand tile stored by TMA with SWIZZLE_128B
// (rows of 128 bytes, 8-row / 1024-byte swizzle atoms): start>>4 | SBO(1024B)>>4 @32 | version 1 @46 | layout 2 @61.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Same for an operand tile with 32-byte rows stored with SWIZZLE_32B (the 16-wide tail of an 80-wide head): 8-row atoms of 256 B.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M x 16 bf16, K-major) is read from tensor memory — 128 lanes x 8 columns,
// each 32-bit column packing elements (2c, 2c+1).  Used for P.V with P written by tcgen05.st straight from the softmax registers.
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Instruction descriptor for kind::f16, A/B = bf16 K-major, D = fp32, shape M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace grove
