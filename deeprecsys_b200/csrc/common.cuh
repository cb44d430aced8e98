// common.cuh — shared device/host helpers for the DeepRecSys B200 backend.
// sm_100a only: mbarrier + cp.async.bulk (1-D TMA) + cache-hinted vector loads.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace drs {

constexpr int kWarp = 32;

// Function attributes (opt-in dynamic shared memory) are per device.  `mask` is one static per
// call site: returns true the first time the site is reached on the CURRENT device (bit =
// device ordinal), so a second engine on another GPU of the same process gets its own opt-in.
inline bool first_use_on_device(std::atomic<uint64_t>& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
  const uint64_t bit = uint64_t(1) << dev;
  return (mask.fetch_or(bit) & bit) == 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals)
               : "memory");
}

// make the init visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(bytes)
      : "memory");
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

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP). src/dst 16-B
// aligned, bytes a multiple of 16; completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------ global loads
// Embedding rows are touched once per query (uniform ids over GBs of table):
// read-only path, do not allocate in L1.
// DRS_ROW_EVICT_FIRST (probe builds, tools/k1_probe.cu): additionally mark the row's L2 line
// evict-first, so that streamed rows do not push ids / activations / weights out of L2.
#ifndef DRS_ROW_EVICT_FIRST
#define DRS_ROW_EVICT_FIRST 0
#endif
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
#if DRS_ROW_EVICT_FIRST
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
#else
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
#endif
  return v;
}

__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) {
  return (a + b - 1) / b;
}

}  // namespace drs
