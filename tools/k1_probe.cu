// k1_probe.cu — variants of the K1 SparseLengthsSum gather on the RMC1 shape, timed side by side.
// Standalone measurement tool (not part of the product path).  Every variant pools T tables x B
// samples x L sorted ids (stratified-uniform over the table, like np.unique of a uniform draw)
// of [N, D = 32] fp32 tables, adds in index order, and writes out[B][T*D]; the outputs of all
// variants must be BIT-IDENTICAL to the production kernel's (checked).
//
//   prod      k_sls_gather_persist<8,false> from csrc/sls_gather.cuh (the shipped kernel),
//             C persistent CTAs per SM
//   stream<S> rows land in SHARED memory instead of registers: every lane keeps S 16-byte
//             cp.async (LDGSTS) copies in flight in its own ring slot column and adds them in
//             order as they complete; the row stream of a lane group runs across segment
//             boundaries, so the queue never drains (registers stop being the landing space:
//             ~40 regs/thread, depth limited by shared memory only)
//
// Build (tools/build_probes.sh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I deeprecsys_b200/csrc \
//        -o tools/bin/k1_probe tools/k1_probe.cu
// Run:  tools/bin/k1_probe [B=16384] [reps=10] [T=8] [N=4000000] [L=80]   -> JSON lines
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include "sls_gather.cuh"

using namespace drs;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __host__ inline uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// small dyadic values: exact sums, so "bit-identical" is also "correct"
__global__ void k_fill_table(float* w, int64_t n, int D, int t) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * D;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D;
    const int c = (int)(i - r * D);
    w[i] = (float)((mix((uint32_t)r) & 511u) + 512u * (uint32_t)(c + 32 * t)) * 1.1920929e-7f;
  }
}

// ids of segment (t, b): L strata of the table, one uniformly placed id per stratum -> sorted, unique
__global__ void k_fill_ids(int32_t* idx, int B, int L, int64_t N, int t) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)B * L;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % L);
    const uint32_t h = mix((uint32_t)i * 2654435761u + 97u * (uint32_t)t + 1u);
    const double u = (h >> 8) * (1.0 / 16777216.0);
    int64_t id = (int64_t)(((double)j + u) * (double)N / (double)L);
    if (id >= N) id = N - 1;
    idx[i] = (int32_t)id;
  }
}

__global__ void k_fill_off(int32_t* off, int B, int L) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= B; i += gridDim.x * blockDim.x)
    off[i] = i * L;
}

// ---------------------------------------------------------------- stream variant
struct StreamParams {
  const float* rows[16];
  const int32_t* idx[16];
  int T, B, L, D;
  float* out;
  int64_t ld_out;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// 8 lanes per 128-byte row; requires L % 8 == 0 and D == 32
template <int S>
__global__ void __launch_bounds__(256) k_gather_stream(const StreamParams p) {
  extern __shared__ __align__(16) float4 ring[];            // [S][256]
  const int tid = threadIdx.x, lane8 = tid & 7;
  const int gid = (blockIdx.x * 256 + tid) >> 3;
  const int G = gridDim.x * 32;
  const int nseg = p.B * p.T;
  const int bps = p.L >> 3;                                  // id blocks per segment
  int nmine = gid < nseg ? (nseg - gid + G - 1) / G : 0;
  // every lane of the warp runs the same number of iterations (shuffles below): the warp's max
  for (int o = 16; o >= 8; o >>= 1) nmine = max(nmine, __shfl_xor_sync(0xffffffffu, nmine, o));
  const int nrows = nmine * p.L;
  const int nblocks = nmine * bps;

  auto load_block = [&](int q) -> int32_t {                  // ids of rows 8q..8q+7 of the stream
    if (q >= nblocks) return 0;
    const int seg = gid + (q / bps) * G;
    if (seg >= nseg) return 0;
    const int t = seg % p.T, b = seg / p.T;
    return __ldg(p.idx[t] + (int64_t)b * p.L + (q % bps) * 8 + lane8);
  };
  int32_t blk_cur = load_block(0), blk_next = load_block(1);
  int qi = 0;                                                // block of the issue pointer
  int seg_i = gid, r_i = 0;

  auto issue = [&](int i) {                                  // row i of the stream -> slot i % S
    const int32_t id = __shfl_sync(0xffffffffu, blk_cur, r_i & 7, 8);
    if (i < nrows && seg_i < nseg) {
      const float* src = p.rows[seg_i % p.T] + (int64_t)id * p.D + lane8 * 4;
      cp_async16(&ring[(i % S) * 256 + tid], src);
    }
    cp_async_commit();
    ++r_i;
    if ((r_i & 7) == 0) { blk_cur = blk_next; ++qi; blk_next = load_block(qi + 1); }
    if (r_i == p.L) { r_i = 0; seg_i += G; }
  };

#pragma unroll 1
  for (int i = 0; i < S; ++i) issue(i);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int seg_c = gid, r_c = 0;
#pragma unroll 1
  for (int i = 0; i < nrows; ++i) {
    cp_async_wait<S - 1>();
    if (seg_c < nseg) {
      const float4 v = ring[(i % S) * 256 + tid];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    issue(i + S);
    if (++r_c == p.L) {
      if (seg_c < nseg) {
        const int t = seg_c % p.T, b = seg_c / p.T;
        *reinterpret_cast<float4*>(p.out + (int64_t)b * p.ld_out + t * p.D + lane8 * 4) = acc;
      }
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      r_c = 0; seg_c += G;
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------- co-runner
// Stand-in for the tensor-core MLP kernels that share the chip with the gather in the real step
// (k_mlp_chain: 128 CTAs x 192 threads, ~160 registers, ~160 KB shared memory, 15 us + 33 us per
// 16,384-sample step): same footprint, spins on FMAs for `cycles` clocks.
template <int R>
__global__ void __launch_bounds__(192, 1) k_dummy_mlp(float* sink, long long cycles) {
  extern __shared__ float dsm[];
  float a[R];
#pragma unroll
  for (int i = 0; i < R; ++i) a[i] = threadIdx.x * 0.001f + i;
  dsm[threadIdx.x] = a[0];
  const long long t0 = clock64();
  do {
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
  } while (clock64() - t0 < cycles);
  float sum = dsm[(threadIdx.x + 1) % 192];
#pragma unroll
  for (int i = 0; i < R; ++i) sum += a[i];
  if (sum == 1234.5f) *sink = sum;
}

// ---------------------------------------------------------------- harness
struct Ctx {
  int T, B, L, D;
  int64_t N;
  std::vector<float*> rows;
  std::vector<int32_t*> idx, off;
  SlsTableDesc* descs;
  float* out;
  float* ref;
  int* err;
  int n_sm;
  double bytes;
};

static float time_launches(int reps, const std::function<void()>& launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

// K1 launches on stream A while stand-in MLP kernels (2 per K1 launch, like the bottom and the
// top stack of a step) run on stream B: wall time from a common start to the end of both.
struct Corun { float k1_ms, all_ms; };
static Corun time_corun(int reps, const std::function<void(cudaStream_t)>& launch, float* sink,
                        int mlp_regs_class, size_t mlp_smem, double sm_ghz) {
  cudaStream_t sa, sb;
  CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
  cudaEvent_t e0, ea, eb;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&ea)); CK(cudaEventCreate(&eb));
  auto mlp = [&](double us) {
    const long long cyc = (long long)(us * 1e3 * sm_ghz);
    if (mlp_regs_class == 0) k_dummy_mlp<150><<<128, 192, mlp_smem, sb>>>(sink, cyc);
    else k_dummy_mlp<70><<<128, 192, mlp_smem, sb>>>(sink, cyc);
  };
  CK(cudaFuncSetAttribute(k_dummy_mlp<150>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_dummy_mlp<70>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int i = 0; i < 2; ++i) { launch(sa); mlp(15); mlp(33); }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0, sa));
  CK(cudaStreamWaitEvent(sb, e0, 0));
  for (int i = 0; i < reps; ++i) { launch(sa); mlp(15); mlp(33); }
  CK(cudaEventRecord(ea, sa));
  CK(cudaEventRecord(eb, sb));
  CK(cudaStreamWaitEvent(sa, eb, 0));
  cudaEvent_t eall;
  CK(cudaEventCreate(&eall));
  CK(cudaEventRecord(eall, sa));
  CK(cudaEventSynchronize(eall));
  Corun r;
  CK(cudaEventElapsedTime(&r.k1_ms, e0, ea));
  CK(cudaEventElapsedTime(&r.all_ms, e0, eall));
  r.k1_ms /= reps; r.all_ms /= reps;
  CK(cudaGetLastError());
  cudaStreamDestroy(sa); cudaStreamDestroy(sb);
  return r;
}

static void report_corun(const Ctx& c, const char* name, int ctas_per_sm, int depth,
                         const char* mlp, Corun r) {
  printf("{\"corun\": \"%s\", \"ctas_per_sm\": %d, \"depth\": %d, \"mlp_standin\": \"%s\", "
         "\"k1_stream_ms_per_step\": %.4f, \"step_ms\": %.4f, \"step_GBps\": %.1f}\n",
         name, ctas_per_sm, depth, mlp, r.k1_ms, r.all_ms, c.bytes / (r.all_ms * 1e-3) / 1e9);
  fflush(stdout);
}

static bool same_as_ref(const Ctx& c) {
  const size_t n = (size_t)c.B * c.T * c.D;
  std::vector<float> a(n), b(n);
  CK(cudaMemcpy(a.data(), c.out, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), c.ref, n * 4, cudaMemcpyDeviceToHost));
  return memcmp(a.data(), b.data(), n * 4) == 0;
}

static void report(const Ctx& c, const char* name, int ctas_per_sm, int depth, size_t smem,
                   float ms, bool same) {
  printf("{\"variant\": \"%s\", \"ctas_per_sm\": %d, \"depth\": %d, \"smem_per_cta\": %zu, "
         "\"unroll\": %d, \"ms\": %.4f, \"GBps\": %.1f, \"M_samples_per_s\": %.2f, "
         "\"bit_identical\": %s}\n",
         name, ctas_per_sm, depth, smem, kSlsUnroll, ms, c.bytes / (ms * 1e-3) / 1e9,
         c.B / (ms * 1e-3) / 1e6, same ? "true" : "false");
  fflush(stdout);
}

static void run_prod(Ctx& c, int reps, int ctas_per_sm, bool is_ref) {
  SlsLaunch p{};
  p.tables = c.descs; p.T = c.T; p.B = c.B; p.D = c.D;
  p.out = is_ref ? c.ref : c.out; p.ld_out = (int64_t)c.T * c.D;
  const int groups = kSlsThreads / 8;
  p.SB = groups / c.T < 1 ? 1 : groups / c.T;
  p.idx_cap = (p.SB * c.L + 3 + 3) & ~3;
  p.vec_store = 1; p.err_flag = c.err;
  const size_t smem = sls_persist_smem_bytes(p.T, p.SB, p.idx_cap, false);
  CK(cudaFuncSetAttribute(k_sls_gather_persist<8, false>,
                          cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  int grid = (c.B + p.SB - 1) / p.SB;
  if (grid > c.n_sm * ctas_per_sm) grid = c.n_sm * ctas_per_sm;
  CK(cudaMemset(p.out, 0, (size_t)c.B * c.T * c.D * 4));
  float ms = time_launches(reps, [&]() {
    k_sls_gather_persist<8, false><<<grid, kSlsThreads, smem>>>(p);
  });
  if (!is_ref) report(c, "prod", ctas_per_sm, kSlsUnroll, smem, ms, same_as_ref(c));
}

static void corun_prod(Ctx& c, int reps, int ctas_per_sm, float* sink, double ghz) {
  SlsLaunch p{};
  p.tables = c.descs; p.T = c.T; p.B = c.B; p.D = c.D;
  p.out = c.out; p.ld_out = (int64_t)c.T * c.D;
  const int groups = kSlsThreads / 8;
  p.SB = groups / c.T < 1 ? 1 : groups / c.T;
  p.idx_cap = (p.SB * c.L + 3 + 3) & ~3;
  p.vec_store = 1; p.err_flag = c.err;
  const size_t smem = sls_persist_smem_bytes(p.T, p.SB, p.idx_cap, false);
  int grid = (c.B + p.SB - 1) / p.SB;
  if (grid > c.n_sm * ctas_per_sm) grid = c.n_sm * ctas_per_sm;
  auto launch = [&](cudaStream_t st) {
    k_sls_gather_persist<8, false><<<grid, kSlsThreads, smem, st>>>(p);
  };
  report_corun(c, "prod", ctas_per_sm, kSlsUnroll, "160 regs, 160 KB",
               time_corun(reps, launch, sink, 0, 160 * 1024, ghz));
  report_corun(c, "prod", ctas_per_sm, kSlsUnroll, "80 regs, 100 KB",
               time_corun(reps, launch, sink, 1, 100 * 1024, ghz));
}

template <int S>
static void corun_stream(Ctx& c, int reps, int ctas_per_sm, float* sink, double ghz) {
  StreamParams p{};
  for (int t = 0; t < c.T; ++t) { p.rows[t] = c.rows[t]; p.idx[t] = c.idx[t]; }
  p.T = c.T; p.B = c.B; p.L = c.L; p.D = c.D; p.out = c.out; p.ld_out = (int64_t)c.T * c.D;
  const size_t smem = (size_t)S * 256 * 16;
  if (smem * ctas_per_sm > 220 * 1024) return;
  CK(cudaFuncSetAttribute(k_gather_stream<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)smem));
  const int grid = c.n_sm * ctas_per_sm;
  auto launch = [&](cudaStream_t st) { k_gather_stream<S><<<grid, 256, smem, st>>>(p); };
  report_corun(c, "stream", ctas_per_sm, S, "160 regs, 160 KB",
               time_corun(reps, launch, sink, 0, 160 * 1024, ghz));
  report_corun(c, "stream", ctas_per_sm, S, "80 regs, 100 KB",
               time_corun(reps, launch, sink, 1, 100 * 1024, ghz));
}

template <int S>
static void run_stream(Ctx& c, int reps, int ctas_per_sm) {
  StreamParams p{};
  for (int t = 0; t < c.T; ++t) { p.rows[t] = c.rows[t]; p.idx[t] = c.idx[t]; }
  p.T = c.T; p.B = c.B; p.L = c.L; p.D = c.D; p.out = c.out; p.ld_out = (int64_t)c.T * c.D;
  const size_t smem = (size_t)S * 256 * 16;
  if (smem * ctas_per_sm > 220 * 1024) return;
  CK(cudaFuncSetAttribute(k_gather_stream<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather_stream<S>, 256, smem));
  if (occ < ctas_per_sm) return;
  CK(cudaMemset(c.out, 0, (size_t)c.B * c.T * c.D * 4));
  const int grid = c.n_sm * ctas_per_sm;
  float ms = time_launches(reps, [&]() { k_gather_stream<S><<<grid, 256, smem>>>(p); });
  report(c, "stream", ctas_per_sm, S, smem, ms, same_as_ref(c));
}

int main(int argc, char** argv) {
  Ctx c{};
  c.B = argc > 1 ? atoi(argv[1]) : 16384;
  const int reps = argc > 2 ? atoi(argv[2]) : 10;
  c.T = argc > 3 ? atoi(argv[3]) : 8;
  c.N = argc > 4 ? atoll(argv[4]) : 4000000;
  c.L = argc > 5 ? atoi(argv[5]) : 80;
  c.D = 32;
  if (c.T > 16 || c.L % 8 != 0) { fprintf(stderr, "T <= 16 and L %% 8 == 0 required\n"); return 1; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  c.n_sm = prop.multiProcessorCount;
  c.bytes = (double)c.B * c.T * ((double)c.L * c.D * 4 + c.L * 4 + 4);
  std::vector<SlsTableDesc> hd(c.T);
  for (int t = 0; t < c.T; ++t) {
    float* w; int32_t *ix, *of;
    CK(cudaMalloc(&w, (size_t)c.N * c.D * 4 + 64));
    CK(cudaMalloc(&ix, ((size_t)c.B * c.L + 16) * 4));
    CK(cudaMemset(ix, 0, ((size_t)c.B * c.L + 16) * 4));
    CK(cudaMalloc(&of, ((size_t)c.B + 1) * 4));
    k_fill_table<<<c.n_sm * 8, 256>>>(w, c.N, c.D, t);
    k_fill_ids<<<c.n_sm * 8, 256>>>(ix, c.B, c.L, c.N, t);
    k_fill_off<<<64, 256>>>(of, c.B, c.L);
    c.rows.push_back(w); c.idx.push_back(ix); c.off.push_back(of);
    hd[t].rows = w; hd[t].idx = ix; hd[t].off = of; hd[t].n_rows = c.N; hd[t].dst_col = t * c.D;
  }
  CK(cudaMalloc(&c.descs, c.T * sizeof(SlsTableDesc)));
  CK(cudaMemcpy(c.descs, hd.data(), c.T * sizeof(SlsTableDesc), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c.out, (size_t)c.B * c.T * c.D * 4));
  CK(cudaMalloc(&c.ref, (size_t)c.B * c.T * c.D * 4));
  CK(cudaMalloc(&c.err, 16));
  CK(cudaMemset(c.err, 0, 16));
  CK(cudaDeviceSynchronize());
  printf("{\"probe\": \"k1\", \"T\": %d, \"N\": %lld, \"D\": %d, \"L\": %d, \"B\": %d, "
         "\"alg_bytes_per_launch\": %.0f, \"sms\": %d}\n",
         c.T, (long long)c.N, c.D, c.L, c.B, c.bytes, c.n_sm);
  run_prod(c, 2, 4, true);                                    // reference output
  for (int k : {2, 3, 4, 5, 6}) run_prod(c, reps, k, false);
  for (int k : {1, 2, 3, 4, 5, 6}) {
    run_stream<8>(c, reps, k);
    run_stream<12>(c, reps, k);
    run_stream<16>(c, reps, k);
    run_stream<24>(c, reps, k);
    run_stream<32>(c, reps, k);
    run_stream<48>(c, reps, k);
  }
  // the step as the engine runs it: gathers back to back on one stream, MLP stand-ins beside them
  float* sink;
  CK(cudaMalloc(&sink, 16));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6;
  for (int k : {3, 4}) corun_prod(c, reps, k, sink, ghz);
  for (int k : {2, 3, 4, 5, 6}) {
    corun_stream<8>(c, reps, k, sink, ghz);
    corun_stream<12>(c, reps, k, sink, ghz);
    corun_stream<16>(c, reps, k, sink, ghz);
  }
  int herr = 0;
  CK(cudaMemcpy(&herr, c.err, 4, cudaMemcpyDeviceToHost));
  printf("{\"err_flag\": %d}\n", herr);
  return 0;
}
