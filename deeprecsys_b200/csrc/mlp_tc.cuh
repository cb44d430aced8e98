// mlp_tc.cuh — K3 tensor-core path: FC + bias + Relu/Sigmoid on tcgen05 (sm_100a).
//
// Replaces the same FC -> Relu|Sigmoid pairs as mlp_ffma.cuh
// (/root/reference/models/dlrm_s_caffe2.py:258-272 and siblings) for the layers
// that really are dense contractions.  Y[M,N] = act(X[M,K] . W[N,K]^T + b).
//
//  * operands stay fp32 in memory; the tensor core consumes them as TF32
//    (kind::tf32, K = 8 per instruction), accumulating in fp32 in TMEM;
//  * TF32X3 mode (error-compensated, fp32-class accuracy): x = x_hi + x_lo,
//    w = w_hi + w_lo with *_hi = top 19 bits, and the kernel issues
//    x_lo.w_hi + x_hi.w_lo + x_hi.w_hi into the same accumulator (the dropped
//    x_lo.w_lo term is ~2^-22 relative).  W is split once at load time
//    ([2][N][K] in HBM); X tiles are split in shared memory by the transform
//    warps right after TMA lands them (elementwise, so swizzle-agnostic);
//  * A (X tile, 128 x 32 fp32) and B (W tile, BN x 32 fp32) are K-major,
//    128-byte-swizzled shared-memory tiles written by TMA
//    (cp.async.bulk.tensor.2d) and described to the MMA by UMMA smem
//    descriptors; one elected thread issues tcgen05.mma; tcgen05.commit
//    recycles stages and signals the epilogue; the epilogue reads the
//    accumulator with tcgen05.ld, adds bias, applies the activation and stores.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = X-split transform during the main loop, then epilogue
// (warp w owns TMEM lanes 32*(w%4)..+31).  grid = (ceil(N/BN), ceil(M/128), units).
#pragma once

#include <map>
#include <mutex>
#include <tuple>

#include <cuda.h>

#include "common.cuh"
#include "mlp_ffma.cuh"

namespace drs {

constexpr int kTcBM = 128;
#ifndef DRS_TC_BK
#define DRS_TC_BK 16
#endif
// fp32 per k-chunk.  32 -> 128-byte rows (SWIZZLE_128B); 16 -> 64-byte rows (SWIZZLE_64B), half the
// stage footprint, so twice the pipeline depth in the same shared memory.
constexpr int kTcBK = DRS_TC_BK;
static_assert(kTcBK == 16 || kTcBK == 32, "k-chunk must be one 64-B or 128-B swizzle row");
constexpr int kTcThreads = 192;
constexpr int kTcUmmaK = 8;        // tf32: 32 bytes per MMA K-step

struct TcFcLaunch {                // host-side description (engine.cu fills it)
  const float* X;  int64_t ldx;  int64_t x_rows;       // X view: x_rows rows in total (all units)
  int64_t x_rows_per_unit;                             // row offset per grid.z
  const float* Wsplit;                                 // [units][2][N][K] (hi, lo)
  const float* bias;  int64_t b_bstride;
  float* Y;  int64_t ldy;  int64_t y_bstride;
  int M, N, K;
  int act;
  bool split3;
};

struct TcFcParams {
  float* Y;  int64_t ldy;  int64_t y_bstride;
  const float* bias;  int64_t b_bstride;
  int M, N, K, act;
  int x_rows_per_unit;
  int w_rows_per_unit;             // 2*N
  int split3;
};

// ------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major swizzled operand tile: rows of kTcBK*4 bytes, 8-row groups SBO bytes apart.
// start>>4 | LBO=1 (unused for swizzled K-major) | SBO>>4 | version=1 | layout type
// (SWIZZLE_128B = 2 with 1024-B groups, SWIZZLE_64B = 4 with 512-B groups)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  constexpr uint64_t kSbo = 8 * kTcBK * 4;
  constexpr uint64_t kLayout = kTcBK == 32 ? 2 : 4;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (kSbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= kLayout << 61;
  return d;
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) |
         ((uint32_t)(kTcBM >> 4) << 24);
}

__host__ __device__ constexpr int tc_stage_bytes(int bn) {
  return 2 * kTcBM * kTcBK * 4 + 2 * bn * kTcBK * 4;   // A_hi, A_lo, B_hi, B_lo
}
__host__ __device__ constexpr int tc_stages(int bn) {
  return (192 * 1024) / tc_stage_bytes(bn) > 8 ? 8 : (192 * 1024) / tc_stage_bytes(bn);
}
__host__ __device__ constexpr int tc_smem_bytes(int bn) {
  return tc_stages(bn) * tc_stage_bytes(bn) + 1024 /*align*/ + 256 /*barriers*/;
}

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
k_fc_tc(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
        const TcFcParams p) {
  constexpr int STAGES = tc_stages(BN);
  constexpr int A_BYTES = kTcBM * kTcBK * 4;       // 16 KB
  constexpr int B_BYTES = BN * kTcBK * 4;
  constexpr int STAGE_BYTES = tc_stage_bytes(BN);
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  extern __shared__ unsigned char smem_dyn[];
  // 1024-B alignment required by SWIZZLE_128B
  // (pointer arithmetic on the shared array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic loads and stores)
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* tiles = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full = bars;                   // TMA -> transform
  uint64_t* xfrm = bars + STAGES;          // transform -> MMA
  uint64_t* empty = bars + 2 * STAGES;     // MMA -> TMA
  uint64_t* accum = bars + 3 * STAGES;     // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * kTcBM, z = blockIdx.z;
  const int nk = (p.K + kTcBK - 1) / kTcBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&xfrm[s], 4);              // one arrival per transform warp
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int xrow = z * p.x_rows_per_unit + m0;
      const int wrow = z * p.w_rows_per_unit + n0;
      const uint32_t tx = A_BYTES + (p.split3 ? 2 : 1) * B_BYTES;
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % STAGES, u = kc / STAGES;
        if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
        unsigned char* st = tiles + (size_t)s * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], tx);
        tma_load_2d(st, &mapX, kc * kTcBK, xrow, &full[s]);
        tma_load_2d(st + 2 * A_BYTES, &mapW, kc * kTcBK, wrow, &full[s]);
        if (p.split3) tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapW, kc * kTcBK, wrow + p.N, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BN);
      uint32_t acc = 0;
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % STAGES, u = kc / STAGES;
        mbar_wait(&xfrm[s], (uint32_t)(u & 1));
        tc_fence_after();
        const uint32_t a_hi = smem_u32(tiles + (size_t)s * STAGE_BYTES);
        const uint32_t a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_BYTES;
        const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int k = 0; k < kTcBK / kTcUmmaK; ++k) {
          const uint32_t koff = (uint32_t)(k * kTcUmmaK * 4);   // bytes inside the 128-B row
          if (p.split3) {
            umma_tf32(tmem_base, umma_desc_sw128(a_lo + koff), umma_desc_sw128(b_hi + koff), idesc, acc);
            acc = 1;
            umma_tf32(tmem_base, umma_desc_sw128(a_hi + koff), umma_desc_sw128(b_lo + koff), idesc, 1);
          }
          umma_tf32(tmem_base, umma_desc_sw128(a_hi + koff), umma_desc_sw128(b_hi + koff), idesc, acc);
          acc = 1;
        }
        umma_commit(&empty[s]);            // frees the stage when these MMAs retire
      }
      umma_commit(accum);                  // accumulator complete
    }
  } else {
    // ===== transform warps (2..5): split the X tile into hi / lo in place =====
    const int t = threadIdx.x - 2 * kWarp;           // 0..127
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % STAGES, u = kc / STAGES;
      mbar_wait(&full[s], (uint32_t)(u & 1));
      if (p.split3) {
        float4* hi = reinterpret_cast<float4*>(tiles + (size_t)s * STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(tiles + (size_t)s * STAGE_BYTES + A_BYTES);
#pragma unroll
        for (int j = 0; j < A_BYTES / 16 / 128; ++j) {   // 128 threads x 16 B per pass
          const int i = j * 128 + t;
          float4 v = hi[i];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          hi[i] = h;
          lo[i] = l;
        }
        fence_proxy_async();               // generic-proxy writes -> visible to the MMA (async proxy)
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&xfrm[s]);
    }

    // ===== epilogue: TMEM -> registers -> (+bias, act) -> smem transpose -> coalesced stores =====
    mbar_wait(accum, 0);                             // all MMAs retired: stage buffers are free
    tc_fence_after();
    const int q = warp % 4;                          // TMEM lane quarter this warp may access
    const int row0 = m0 + q * 32;                    // first row of this warp's 32-row slab
    float* __restrict__ ybase = p.Y + (int64_t)z * p.y_bstride;
    const float* __restrict__ bias = p.bias + (int64_t)z * p.b_bstride;
    constexpr int CH = BN >= 32 ? 32 : 16;
    // per-warp 32 x (CH+1) fp32 staging tile carved from the (now idle) pipeline stages
    float* stg = reinterpret_cast<float*>(tiles) + (size_t)(warp - 2) * 32 * 33;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CH) {
      uint32_t r[CH];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if (CH == 32) tmem_ld_32x32b_x32(taddr, r);
      else tmem_ld_32x32b_x16(taddr, r);
#pragma unroll
      for (int j = 0; j < CH; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);   // row = lane
      __syncwarp();
      const int n = n0 + c0 + lane;                  // lane <-> column: 128-B coalesced rows
      if (lane < CH && n < p.N) {
        const float bn_ = __ldg(bias + n);
        const int rmax = min(32, p.M - row0);
        for (int rr = 0; rr < rmax; ++rr)
          ybase[(int64_t)(row0 + rr) * p.ldy + n] = apply_act(stg[rr * 33 + lane] + bn_, p.act);
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ persistent variant
// Same arithmetic and tiles as k_fc_tc, restructured so that the tensor pipe does not idle
// between tiles (one CTA per tile paid TMEM allocation, the pipeline
// fill and a ~10 us register/smem epilogue per tile with nothing overlapping them — the pipe was
// 73 % busy on Wide&Deep's 1376 -> 1024 layer):
//   * one CTA per SM loops over its tiles (tile = blockIdx.x + i * gridDim.x, n fastest so that
//     neighbouring CTAs read the same X rows out of L2);
//   * TWO accumulators in TMEM (2 x BN columns): the MMA warp starts tile i+1 while the epilogue
//     warps drain tile i;
//   * the smem ring runs across tile boundaries (the first chunks of tile i+1 are loading while
//     tile i's last MMAs retire);
//   * dedicated epilogue warps (6..9) with their own staging tile, so the transform warps (2..5)
//     keep splitting X chunks of the next tile; in single-pass tf32 mode the MMA warp consumes
//     the TMA barrier directly and the transform warps stay out of the loop.
constexpr int kTcPThreads = 320;                 // TMA, MMA, 4 transform, 4 epilogue warps
constexpr int kTcPStageTile = 32 * 33 * 4;       // per epilogue warp: 32 rows x 33 floats

__host__ __device__ constexpr int tcp_stages(int bn) {
  return (200 * 1024) / tc_stage_bytes(bn) > 8 ? 8 : (200 * 1024) / tc_stage_bytes(bn);
}
__host__ __device__ constexpr int tcp_smem_bytes(int bn) {
  return tcp_stages(bn) * tc_stage_bytes(bn) + 4 * kTcPStageTile + 1024 /*align*/ + 512 /*barriers*/;
}

struct TcTileGrid { int tiles_n, tiles_m, units; };

template <int BN>
__global__ void __launch_bounds__(kTcPThreads, 1)
k_fc_tc_persist(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                const TcFcParams p, const TcTileGrid tg) {
  constexpr int STAGES = tcp_stages(BN);
  constexpr int A_BYTES = kTcBM * kTcBK * 4;
  constexpr int B_BYTES = BN * kTcBK * 4;
  constexpr int STAGE_BYTES = tc_stage_bytes(BN);
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;          // columns of one accumulator
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;              // power of two, <= 512

  extern __shared__ unsigned char smem_dyn[];
  // (pointer arithmetic on the shared array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic loads and stores)
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* tiles = base;
  float* stage_tiles = reinterpret_cast<float*>(base + (size_t)STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * STAGE_BYTES +
                                               4 * kTcPStageTile);
  uint64_t* full = bars;                    // TMA -> transform (tf32x3) / MMA (tf32)
  uint64_t* xfrm = bars + STAGES;           // transform -> MMA
  uint64_t* empty = bars + 2 * STAGES;      // MMA -> TMA
  uint64_t* acc_full = bars + 3 * STAGES;   // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;       // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int nk = (p.K + kTcBK - 1) / kTcBK;
  const int ntiles = tg.tiles_n * tg.tiles_m * tg.units;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&xfrm[s], 4);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);          // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: chunks of all my tiles, one continuous ring
    if (lane == 0) {
      const uint32_t tx = A_BYTES + (p.split3 ? 2 : 1) * B_BYTES;
      int c = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tn = tile % tg.tiles_n, tm = (tile / tg.tiles_n) % tg.tiles_m;
        const int z = tile / (tg.tiles_n * tg.tiles_m);
        const int xrow = z * p.x_rows_per_unit + tm * kTcBM;
        const int wrow = z * p.w_rows_per_unit + tn * BN;
        for (int kc = 0; kc < nk; ++kc, ++c) {
          const int s = c % STAGES, u = c / STAGES;
          if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
          unsigned char* st = tiles + (size_t)s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full[s], tx);
          tma_load_2d(st, &mapX, kc * kTcBK, xrow, &full[s]);
          tma_load_2d(st + 2 * A_BYTES, &mapW, kc * kTcBK, wrow, &full[s]);
          if (p.split3)
            tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapW, kc * kTcBK, wrow + p.N, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulators ping-pong between two TMEM regions
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BN);
      uint64_t* ready = p.split3 ? xfrm : full;
      int c = 0, t = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const int buf = t & 1;
        if (t >= 2) {                        // the epilogue of tile t-2 has drained this region
          mbar_wait(&acc_empty[buf], (uint32_t)(((t >> 1) - 1) & 1));
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + (uint32_t)buf * ACC_COLS;
        uint32_t acc = 0;
        for (int kc = 0; kc < nk; ++kc, ++c) {
          const int s = c % STAGES, u = c / STAGES;
          mbar_wait(&ready[s], (uint32_t)(u & 1));
          tc_fence_after();
          const uint32_t a_hi = smem_u32(tiles + (size_t)s * STAGE_BYTES);
          const uint32_t a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES;
          const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int k = 0; k < kTcBK / kTcUmmaK; ++k) {
            const uint32_t koff = (uint32_t)(k * kTcUmmaK * 4);
            if (p.split3) {
              umma_tf32(tacc, umma_desc_sw128(a_lo + koff), umma_desc_sw128(b_hi + koff), idesc, acc);
              acc = 1;
              umma_tf32(tacc, umma_desc_sw128(a_hi + koff), umma_desc_sw128(b_lo + koff), idesc, 1);
            }
            umma_tf32(tacc, umma_desc_sw128(a_hi + koff), umma_desc_sw128(b_hi + koff), idesc, acc);
            acc = 1;
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else if (warp < 6) {
    // ===== transform warps (2..5): tf32x3 only — split the landed X chunk into hi / lo in place
    if (p.split3) {
      const int tt = threadIdx.x - 2 * kWarp;          // 0..127
      int c = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nk; ++kc, ++c) {
          const int s = c % STAGES, u = c / STAGES;
          mbar_wait(&full[s], (uint32_t)(u & 1));
          float4* hi = reinterpret_cast<float4*>(tiles + (size_t)s * STAGE_BYTES);
          float4* lo = reinterpret_cast<float4*>(tiles + (size_t)s * STAGE_BYTES + A_BYTES);
#pragma unroll
          for (int j = 0; j < A_BYTES / 16 / 128; ++j) {
            const int i = j * 128 + tt;
            const float4 v = hi[i];
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            hi[i] = h;
            lo[i] = l;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&xfrm[s]);
        }
      }
    }
  } else {
    // ===== epilogue warps (6..9): TMEM -> registers -> staging tile -> coalesced stores
    const int q = warp % 4;                            // TMEM lane quarter this warp may access
    float* stg = stage_tiles + (size_t)(warp - 6) * 32 * 33;
    constexpr int CH = BN >= 32 ? 32 : 16;
    int t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      const int tn = tile % tg.tiles_n, tm = (tile / tg.tiles_n) % tg.tiles_m;
      const int z = tile / (tg.tiles_n * tg.tiles_m);
      const int n0 = tn * BN, row0 = tm * kTcBM + q * 32;
      float* __restrict__ ybase = p.Y + (int64_t)z * p.y_bstride;
      const float* __restrict__ bias = p.bias + (int64_t)z * p.b_bstride;
      mbar_wait(&acc_full[buf], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t r[CH];
        const uint32_t taddr = tmem_base + (uint32_t)buf * ACC_COLS + ((uint32_t)(q * 32) << 16) +
                               (uint32_t)c0;
        if (CH == 32) tmem_ld_32x32b_x32(taddr, r);
        else tmem_ld_32x32b_x16(taddr, r);
#pragma unroll
        for (int j = 0; j < CH; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int n = n0 + c0 + lane;
        if (lane < CH && n < p.N) {
          const float bn_ = __ldg(bias + n);
          const int rmax = min(32, p.M - row0);
          for (int rr = 0; rr < rmax; ++rr)
            ybase[(int64_t)(row0 + rr) * p.ldy + n] = apply_act(stg[rr * 33 + lane] + bn_, p.act);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);     // this warp's quarter is drained
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// W [n] -> Wsplit[0][n] = top 19 bits, Wsplit[1][n] = W - hi   (per unit: [2][N*K])
__global__ void k_split_tf32(const float* __restrict__ W, float* __restrict__ out, int64_t count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float w = W[i];
    const float h = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    out[i] = h;
    out[count + i] = w - h;
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D fp32 row-major view [rows, cols] with row pitch ld floats; box = [box_rows, 32 floats].
// A descriptor depends only on (address, shape, pitch, box), and the serving loop asks for the
// same few over and over (weights, per-lane activation buffers), so encoded descriptors are
// cached: cuTensorMapEncodeTiled costs ~1 us and a query needs 7 of them.
struct TcMapKey {
  const void* ptr; int64_t rows, cols, ld; int box_rows, bk;
  bool operator<(const TcMapKey& o) const {
    return std::tie(ptr, rows, cols, ld, box_rows, bk) <
           std::tie(o.ptr, o.rows, o.cols, o.ld, o.box_rows, o.bk);
  }
};

// bk: floats of K per box (32 -> SWIZZLE_128B, 16 -> SWIZZLE_64B)
inline bool tc_make_map_bk(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols,
                           int64_t ld, int box_rows, int bk) {
  static std::mutex mu;
  static std::map<TcMapKey, CUtensorMap> cache;
  const TcMapKey key{ptr, rows, cols, ld, box_rows, bk};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *map = it->second; return true; }
  }
  PFN_encodeTiled enc = tc_encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 4096) cache.clear();          // bounded: engines come and go in tests
  cache[key] = *map;
  return true;
}

inline bool tc_make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows) {
  return tc_make_map_bk(map, ptr, rows, cols, ld, box_rows, kTcBK);
}

inline int tc_pick_bn(int N) {
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  return 256;
}

inline bool tc_fc_supported(int N, int K, int64_t ldx, const float* X) {
  return N >= 1 && K >= 4 && (K % 4 == 0) && (ldx % 4 == 0) &&
         ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && tc_encode_fn() != nullptr;
}

inline int tc_sm_count() {
  static int n_sm[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return 148;
  if (n_sm[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;
    n_sm[dev] = v;
  }
  return n_sm[dev];
}

template <int BN>
inline cudaError_t tc_launch_persist_bn(const CUtensorMap& mx, const CUtensorMap& mw,
                                        const TcFcParams& p, dim3 grid, cudaStream_t stream,
                                        int max_ctas) {
  static std::atomic<uint64_t> attr{0};
  if (first_use_on_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(k_fc_tc_persist<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tcp_smem_bytes(BN));
    if (e != cudaSuccess) return e;
  }
  TcTileGrid tg{(int)grid.x, (int)grid.y, (int)grid.z};
  const int ntiles = tg.tiles_n * tg.tiles_m * tg.units;
  int ctas = ntiles < tc_sm_count() ? ntiles : tc_sm_count();
  if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
  k_fc_tc_persist<BN><<<ctas, kTcPThreads, tcp_smem_bytes(BN), stream>>>(mx, mw, p, tg);
  return cudaGetLastError();
}

template <int BN>
inline cudaError_t tc_launch_bn(const CUtensorMap& mx, const CUtensorMap& mw, const TcFcParams& p,
                                dim3 grid, cudaStream_t stream) {
  static std::atomic<uint64_t> attr{0};
  if (first_use_on_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(k_fc_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tc_smem_bytes(BN));
    if (e != cudaSuccess) return e;
  }
  k_fc_tc<BN><<<grid, kTcThreads, tc_smem_bytes(BN), stream>>>(mx, mw, p);
  return cudaGetLastError();
}

// returns a cudaError_t value (0 = ok); cudaErrorInvalidValue if the tensor map cannot be built
// persistent: 0 = one CTA per tile; 1 = persistent tiles when the grid exceeds one wave;
// n >= 2 = always persistent with at most n CTAs (tests: many tiles per CTA on small problems)
inline int tc_fc_launch(const TcFcLaunch& q, int n_units, cudaStream_t stream,
                        int persistent = 1) {
  const int bn = tc_pick_bn(q.N);
  CUtensorMap mx, mw;
  if (!tc_make_map(&mx, q.X, q.x_rows, q.K, q.ldx, kTcBM)) return (int)cudaErrorInvalidValue;
  if (!tc_make_map(&mw, q.Wsplit, (int64_t)n_units * 2 * q.N, q.K, q.K, bn))
    return (int)cudaErrorInvalidValue;
  TcFcParams p{};
  p.Y = q.Y; p.ldy = q.ldy; p.y_bstride = q.y_bstride;
  p.bias = q.bias; p.b_bstride = q.b_bstride;
  p.M = q.M; p.N = q.N; p.K = q.K; p.act = q.act;
  p.x_rows_per_unit = (int)q.x_rows_per_unit;
  p.w_rows_per_unit = 2 * q.N;
  p.split3 = q.split3 ? 1 : 0;
  dim3 grid(ceil_div(q.N, bn), ceil_div(q.M, kTcBM), n_units);
  // persistent tiles pay off once a CTA has more than one tile; a grid that fits one wave keeps
  // the one-CTA-per-tile kernel (less shared memory: it can share an SM with the gather)
  if (persistent >= 2 || (persistent == 1 && (int64_t)grid.x * grid.y * grid.z > tc_sm_count())) {
    const int cap = persistent >= 2 ? persistent : 0;
    switch (bn) {
      case 16:  return (int)tc_launch_persist_bn<16>(mx, mw, p, grid, stream, cap);
      case 32:  return (int)tc_launch_persist_bn<32>(mx, mw, p, grid, stream, cap);
      case 64:  return (int)tc_launch_persist_bn<64>(mx, mw, p, grid, stream, cap);
      case 128: return (int)tc_launch_persist_bn<128>(mx, mw, p, grid, stream, cap);
      default:  return (int)tc_launch_persist_bn<256>(mx, mw, p, grid, stream, cap);
    }
  }
  switch (bn) {
    case 16:  return (int)tc_launch_bn<16>(mx, mw, p, grid, stream);
    case 32:  return (int)tc_launch_bn<32>(mx, mw, p, grid, stream);
    case 64:  return (int)tc_launch_bn<64>(mx, mw, p, grid, stream);
    case 128: return (int)tc_launch_bn<128>(mx, mw, p, grid, stream);
    default:  return (int)tc_launch_bn<256>(mx, mw, p, grid, stream);
  }
}

inline int tc_split_weights(const float* W, float* Wsplit, int64_t count, cudaStream_t stream) {
  int64_t g = ceil_div64(count, 256);
  if (g > 148 * 8) g = 148 * 8;
  k_split_tf32<<<(int)g, 256, 0, stream>>>(W, Wsplit, count);
  return (int)cudaGetLastError();
}

}  // namespace drs
