// hbm_probe.cu — what can HBM deliver on this GPU for the access patterns of the
// SparseLengthsSum gather?  Standalone measurement tool (not part of the product path):
//   seq_read    : every byte of a 4 GiB buffer once, 16-B loads, read-only
//   copy        : the same buffer copied (read + write bytes) — MEASURED_PEAKS.json's method
//   random rows : R-byte rows (R = 128, 256, 512) at uniformly random row ids from a 4 GiB
//                 table, one lane group per row, U independent row loads in flight per group,
//                 ids streamed from a 64 MiB id array (counted in the bytes)
// The random-row figure is the practical ceiling for K1 (rows of D=32 fp32 = 128 B).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/hbm_probe tools/hbm_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_nc(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__global__ void k_seq_read(const float4* __restrict__ src, size_t n, float* sink) {
  float4 acc = make_float4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    float4 a = ldg_nc(src + i), b = ldg_nc(src + i + stride), c = ldg_nc(src + i + 2 * stride),
           d = ldg_nc(src + i + 3 * stride);
    acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y;
    acc.z += a.z + b.z + c.z + d.z; acc.w += a.w + b.w + c.w + d.w;
  }
  for (; i < n; i += stride) { float4 a = ldg_nc(src + i); acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w; }
  if (acc.x + acc.y + acc.z + acc.w == 1234.5678f) *sink = acc.x;
}

__global__ void k_copy(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

// LPR lanes cover one row (LPR * 16 bytes); each lane group walks its own id list, U loads in flight.
template <int LPR, int U>
__global__ void __launch_bounds__(256) k_random_rows(const float4* __restrict__ table,
                                                     const uint32_t* __restrict__ ids,
                                                     size_t n_ids, float* sink) {
  const int lane = threadIdx.x % LPR;
  const size_t group = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const size_t ngroups = ((size_t)gridDim.x * blockDim.x) / LPR;
  const size_t per = n_ids / ngroups;                 // contiguous id span per group
  const uint32_t* my = ids + group * per;
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t i = 0; i + U <= per; i += U) {
    uint32_t id[U];
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) id[u] = __ldg(my + i + u);
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg_nc(table + (size_t)id[u] * LPR + lane);
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 1234.5678f) *sink = acc.x;
}

template <typename F>
static float time_ms(F launch, int reps, float* best) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float tot = 0, bst = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    tot += ms; if (ms < bst) bst = ms;
  }
  CK(cudaGetLastError());
  *best = bst;
  return tot / reps;
}

template <int LPR, int U>
static void run_random(const float4* table, size_t table_bytes, const uint32_t* ids, size_t n_ids,
                       float* sink, int ctas_per_sm, int n_sm) {
  const int grid = n_sm * ctas_per_sm;
  float best;
  float ms = time_ms([&]() { k_random_rows<LPR, U><<<grid, 256>>>(table, ids, n_ids, sink); }, 10, &best);
  const size_t ngroups = (size_t)grid * 256 / LPR;
  const size_t used = (n_ids / ngroups) / U * U * ngroups;
  const double bytes = (double)used * (LPR * 16 + 4);
  printf("{\"probe\": \"random_rows\", \"row_bytes\": %d, \"unroll\": %d, \"ctas_per_sm\": %d, "
         "\"rows\": %zu, \"avg_GBps\": %.1f, \"best_GBps\": %.1f}\n",
         LPR * 16, U, ctas_per_sm, used, bytes / ms / 1e6, bytes / best / 1e6);
  (void)table_bytes;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int n_sm = prop.multiProcessorCount;
  const size_t bytes = (size_t)4 << 30;
  float4 *src, *dst;
  float* sink;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(src, 1, bytes)); CK(cudaMemset(dst, 0, bytes));
  const size_t n16 = bytes / 16;
  float best;
  float ms = time_ms([&]() { k_seq_read<<<n_sm * 8, 256>>>(src, n16, sink); }, 10, &best);
  printf("{\"probe\": \"seq_read\", \"bytes\": %zu, \"avg_GBps\": %.1f, \"best_GBps\": %.1f}\n", bytes,
         bytes / ms / 1e6, bytes / best / 1e6);
  ms = time_ms([&]() { k_copy<<<n_sm * 8, 256>>>(src, dst, n16); }, 10, &best);
  printf("{\"probe\": \"copy\", \"bytes_rw\": %zu, \"avg_GBps\": %.1f, \"best_GBps\": %.1f}\n", 2 * bytes,
         2.0 * bytes / ms / 1e6, 2.0 * bytes / best / 1e6);
  ms = time_ms([&]() { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice)); }, 10, &best);
  printf("{\"probe\": \"cudaMemcpyD2D\", \"bytes_rw\": %zu, \"avg_GBps\": %.1f, \"best_GBps\": %.1f}\n",
         2 * bytes, 2.0 * bytes / ms / 1e6, 2.0 * bytes / best / 1e6);

  // random row ids (xorshift), 16 Mi ids = 64 MiB
  const size_t n_ids = (size_t)16 << 20;
  std::vector<uint32_t> h(n_ids);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  uint32_t* ids;
  CK(cudaMalloc(&ids, n_ids * 4));
  for (int rb : {128, 256, 512}) {
    const uint32_t n_rows = (uint32_t)(bytes / rb);
    for (size_t i = 0; i < n_ids; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      h[i] = (uint32_t)((s >> 20) % n_rows);
    }
    CK(cudaMemcpy(ids, h.data(), n_ids * 4, cudaMemcpyHostToDevice));
    for (int c : {2, 4, 8}) {
      if (rb == 128) {
        run_random<8, 4>(src, bytes, ids, n_ids, sink, c, n_sm);
        run_random<8, 8>(src, bytes, ids, n_ids, sink, c, n_sm);
        run_random<8, 16>(src, bytes, ids, n_ids, sink, c, n_sm);
      } else if (rb == 256) {
        run_random<16, 8>(src, bytes, ids, n_ids, sink, c, n_sm);
      } else {
        run_random<32, 8>(src, bytes, ids, n_ids, sink, c, n_sm);
      }
    }
  }
  return 0;
}
