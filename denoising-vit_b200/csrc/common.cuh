// Common device/host helpers for the dvt_b200 kernels (sm_100a only).
//
// Thin inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld)
// and the fences between the generic, async and tensor proxies.  Every blocking wait in this file is
// bounded by a clock watchdog: a kernel that would dead-lock records a code in g_dvt_dev_error and traps
// instead of hanging the GPU.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>

#ifndef DVT_WATCHDOG_CYCLES
#define DVT_WATCHDOG_CYCLES 4000000000ll  // ~2 s at 1.9 GHz
#endif

namespace dvt {

// ------------------------------------------------------------------------------------------------
// host-side error plumbing (api.cu owns the storage)
// ------------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define DVT_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return ::dvt::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define DVT_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      ::dvt::set_last_error(__VA_ARGS__);         \
      return DVT_ERR_INVALID;                     \
    }                                             \
  } while (0)

enum { DVT_OK = 0, DVT_ERR_INVALID = 1, DVT_ERR_CUDA = 2, DVT_ERR_DEVICE = 3 };

// device-side error word: 0 = fine; otherwise (code << 16 | detail)
// (defined here: the library is built as ONE translation unit, see dvt_b200_all.cu)
__device__ unsigned int g_dvt_dev_error = 0;

// ------------------------------------------------------------------------------------------------
// small device utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void dev_fail(unsigned code, unsigned detail) {
  atomicCAS(&g_dvt_dev_error, 0u, (code << 16) | (detail & 0xffffu));
  __threadfence_system();
  __trap();
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// (A suspend-time hint on try_wait -- `mbarrier.try_wait ..., hint_ns` -- was measured in round 2: no effect on the attention
// kernel, 0.365 -> 0.368 ms, and the fit's step chain got slower, so the plain form stays.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: traps (instead of hanging) when the phase never completes.
// Wait of a role that is not latency critical (a TMA producer waiting for a free slot): backs off with nanosleep so that
// its spin loop does not take issue slots from the math warps on the same scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (clock64() - t0 > DVT_WATCHDOG_CYCLES) dev_fail(0xDEADu, tag);
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > DVT_WATCHDOG_CYCLES) dev_fail(0xDEADu, tag);
  }
}

// ------------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): a kernel launched with launch_k(pdl = true, ...) may start -- block scheduling,
// barrier / TMEM set-up -- while the previous kernel of its stream is still running; pdl_wait() blocks until that kernel
// has completed and its writes are visible, pdl_trigger() lets the NEXT kernel of the stream start its own prologue.
// Both are no-ops in a kernel that was launched without the attribute.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// prio_drop > 0: the kernel runs that many levels BELOW the highest stream priority whatever its stream's priority is
// (kernels off the critical path that would otherwise take the SMs the next critical kernel is waiting for).
struct LaunchOpt {
  bool pdl = false;
  int prio_drop = 0;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kx(LaunchOpt o, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (o.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (o.prio_drop > 0) {
    static int lo = 0, hi = 0;
    static const cudaError_t range_rc = cudaDeviceGetStreamPriorityRange(&lo, &hi);  // (numerically lower = higher priority)
    if (range_rc != cudaSuccess) return range_rc;
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = hi + o.prio_drop < lo ? hi + o.prio_drop : lo;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  return launch_kx(LaunchOpt{pdl, 0}, kern, grid, block, smem, st, static_cast<Args&&>(args)...);
}

// ------------------------------------------------------------------------------------------------
// proxy fences
// ------------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA store / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA (tiled mode).  Coordinates are innermost-first.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 1-D bulk copy global -> shared (no tensor map): size and both addresses multiples of 16 bytes
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ------------------------------------------------------------------------------------------------
// Whole-warp calls (.sync.aligned).  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1").  layout_type: 0 none, 2 SW128, 4 SW64, 6 SW32.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;
  return d;
}

// Instruction descriptor, kind::f16 / kind::tf32, fp32 accumulate.
//   fmt: 0 = f16, 1 = bf16, 2 = tf32.  a_mn / b_mn: 1 = MN-major operand, 0 = K-major.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; single-thread issue.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from tensor memory (lane = row of the M128 tile, each 32-bit
// column holds two consecutive K elements of a 16-bit type; K-major by construction).
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers, 32 lanes x 32 consecutive 32-bit columns (thread i of the warp gets lane base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM, 32 lanes x 32 columns
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(r[0]),
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
      "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
      "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// math helpers
// ------------------------------------------------------------------------------------------------
// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: one issue slot for two independent fp32 operations, each rounded
// exactly like its scalar form).  The 64-bit moves are register-pair bookkeeping and disappear in SASS.
__device__ __forceinline__ uint64_t f2_pack(float2 a) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(uint64_t r) {
  float2 a;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// erf with ONE transcendental: erfc(t) = 2^q(t) for t = |x|, q a degree-6 polynomial without constant term fitted to
// log2(erfc(t)) on [0, 4.2] (weighted so that the absolute error of erf is minimised: max |err| 3.3e-7 in fp32, evaluated
// against scipy.special.erf on 2e5 points); erf(x) = sign(x) (1 - 2^q).  t is clamped to 6 (erfc(6) = 2e-17 rounds 1 - e
// to 1; the polynomial turns upward far outside its fitting range).  Round 1 used Abramowitz-Stegun 7.1.26 (a reciprocal
// AND an exponential per element): in the fc1 + GELU epilogue the MUFU pipe -- 2 ops x 32 768 elements per tile -- cost
// 2/3 of the tile's MMA time.
__device__ __forceinline__ float fast_erf(float x) {
  const float t = fminf(fabsf(x), 6.0f);
  float q = 1.580459628734477e-4f;
  q = fmaf(q, t, -3.742739173536956e-3f);
  q = fmaf(q, t, 3.1032528216293022e-2f);
  q = fmaf(q, t, -1.498016394645605e-1f);
  q = fmaf(q, t, -9.181337298819333e-1f);
  q = fmaf(q, t, -1.6279281218285298f);
  q *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return copysignf(1.0f - e, x);
}
// GELU(x) = x Phi(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2): with erfc = 2^q as in fast_erf (same polynomial, written in
// a = |x| with the -1 of the factor 0.5 folded into the exponent) this is 11 instructions per element instead of 18 -- the
// fc1 epilogue is issue bound (ncu: issue 50 %, tensor 51 %).  Max abs error vs the exact erf GELU 5.2e-7.
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fminf(fabsf(x), 8.485281374f);
  float q = 1.9755745359180961e-05f;
  q = fmaf(q, a, -6.6162906245512902e-04f);
  q = fmaf(q, a, 7.7581320540732555e-03f);
  q = fmaf(q, a, -5.2962877549126521e-02f);
  q = fmaf(q, a, -4.5906686494096666e-01f);
  q = fmaf(q, a, -1.1511190142292334f);
  q = fmaf(q, a, -1.0f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return fmaf(-a, e, fmaxf(x, 0.0f));
}

// Two GELUs in packed fp32 pairs: the polynomial is evaluated in n = -min(|x|, 8.485) (one FMNMX with source modifiers; the
// odd coefficients change sign, every intermediate is the exact negation or copy of the scalar form's, so the result is
// bit-identical to gelu_erf), 7 FFMA2 + 2 FMNMX + 2 FMNMX + 2 MUFU per pair: 6.5 instead of 10 issue slots per element.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 n = make_float2(fmaxf(-fabsf(x.x), -8.485281374f), fmaxf(-fabsf(x.y), -8.485281374f));
  float2 q = make_float2(1.9755745359180961e-05f, 1.9755745359180961e-05f);
  q = ffma2(q, n, make_float2(6.6162906245512902e-04f, 6.6162906245512902e-04f));
  q = ffma2(q, n, make_float2(7.7581320540732555e-03f, 7.7581320540732555e-03f));
  q = ffma2(q, n, make_float2(5.2962877549126521e-02f, 5.2962877549126521e-02f));
  q = ffma2(q, n, make_float2(-4.5906686494096666e-01f, -4.5906686494096666e-01f));
  q = ffma2(q, n, make_float2(1.1511190142292334f, 1.1511190142292334f));
  q = ffma2(q, n, make_float2(-1.0f, -1.0f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
  return ffma2(n, e, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

// d/dx gelu_erf(x) = Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_grad(float x) {
  return fmaf(x * 0.3989422804014327f, __expf(-0.5f * x * x), 0.5f * (1.0f + fast_erf(x * 0.70710678118654752f)));
}

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

// ------------------------------------------------------------------------------------------------
// host: TMA descriptor creation (driver entry point resolved at run time; no link against libcuda)
// ------------------------------------------------------------------------------------------------
enum TmapDtype { TMAP_BF16 = 0, TMAP_F32 = 1 };
// 2-D row-major tensor [rows, cols] with `row_pitch_bytes`; box = [box_rows, box_cols]; 128B swizzle.
int make_tmap_2d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t rows, uint64_t cols,
                 uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols);
// 3-D tensor [d2, d1, d0(contiguous)] with byte strides for d1 and d2; box = [1, box1, box0].
int make_tmap_3d(CUtensorMap* out, const void* base, TmapDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2 = 1,
                 bool swizzle_atom32 = false /* CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: MN-major TF32 operands */);

int num_sms();

// host-side tally of kernels launched by this library (graph replays add the node count of the graph)
void count_launch(long long n = 1);
// Set by vit_forward around its block loop: LayerNorm / attention launches then use programmatic dependent launch (each
// process drives one GPU from one thread).
extern bool g_vit_pdl;
long long launch_count();

}  // namespace dvt
