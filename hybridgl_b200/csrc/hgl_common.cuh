// Shared device/host helpers for libhgl (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hgl.h"

namespace hgl {

// ---- error plumbing (host) --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int launch_status(const char* what);   // cudaGetLastError -> HGL_ECUDA / HGL_OK

#define HGL_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::hgl::set_error(__VA_ARGS__);           \
      return HGL_EINVAL;                       \
    }                                          \
  } while (0)

// Tuning hooks (environment variables) exist only in builds with -DHGL_TUNING; the shipped library ignores the environment.
#ifdef HGL_TUNING
#include <stdlib.h>
static inline int tuning_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
static inline bool tuning_flag(const char* name) { return getenv(name) != nullptr; }
#else
static inline int tuning_int(const char*, int dflt) { return dflt; }
static inline bool tuning_flag(const char*) { return false; }
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
int sm_count();
// Opt a kernel in to `bytes` of dynamic shared memory (cudaFuncAttributeMaxDynamicSharedMemorySize) ONCE per (kernel, device)
// and size: later launches that need no more than what was already granted cost a table lookup, not a driver call.
// Returns HGL_OK / HGL_ECUDA (error text set).
int ensure_dyn_smem(const void* kernel, size_t bytes, const char* what);

// The stream's priority as an explicit launch attribute: a kernel launched with it keeps that priority as a NODE of a captured
// graph too (the small latency-bound kernels of the scoring chain must win SM slots from the bandwidth-bound prep kernel as
// its CTAs retire, whether the pass is launched eagerly or replayed).
static inline cudaLaunchAttribute priority_attr(cudaStream_t st) {
  cudaLaunchAttribute a;
  int prio = 0;
  if (cudaStreamGetPriority(st, &prio) != cudaSuccess) { prio = 0; (void)cudaGetLastError(); }
  a.id = cudaLaunchAttributePriority;
  a.val.priority = prio;
  return a;
}

// ---- device helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// streaming stores (outputs are consumed by a later kernel, never re-read here)
__device__ __forceinline__ void stg_stream(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(uint2* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }

// ---- mbarrier + 1-D bulk async copy (TMA, SASS UBLKCP) ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace hgl
