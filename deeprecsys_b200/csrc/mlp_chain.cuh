// mlp_chain.cuh — K3 fused MLP chain: up to 4 FC(+bias+Relu/Sigmoid) layers in ONE
// tcgen05 kernel, activations never leaving the SM between layers.
//
// Replaces a whole create_mlp() stack (/root/reference/models/dlrm_s_caffe2.py:223-279:
// FC -> Relu ... FC -> Sigmoid) for one 128-row tile of samples per CTA.
//
//   layer 0   A chunks [128 x 16 fp32] come from HBM by TMA, are split hi/lo in shared memory;
//   layer l>0 A chunks are produced on-chip: the epilogue warps read layer l-1's accumulator
//             from TMEM (tcgen05.ld), add bias, apply the activation, split hi/lo and write the
//             chunk straight into the UMMA operand ring in the K-major SWIZZLE_64B layout;
//   W chunks  [N_l x 16] (hi, lo) of every layer stream through a second ring by TMA;
//   MMA       one elected thread issues tcgen05.mma kind::tf32 (lo.hi + hi.lo + hi.hi) into the
//             layer's TMEM accumulator (layers ping-pong between two 256-column regions);
//   output    last layer: TMEM -> registers -> bias/act -> smem transpose -> coalesced stores.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..5
// transform / inter-layer epilogue / output (warp w owns TMEM lanes 32*(w%4)..+31).
// Same numerics as k_fc_tc (mlp_tc.cuh): 3xTF32 error-compensated, fp32 accumulate.
#pragma once

#include "mlp_tc.cuh"

namespace drs {

constexpr int kChainMaxLayers = 4;
constexpr int kChainBK = 16;                 // fp32 per chunk = one 64-byte swizzle row
#ifndef DRS_CHAIN_A_SLOTS
#define DRS_CHAIN_A_SLOTS 4
#endif
#ifndef DRS_CHAIN_B_SLOTS
#define DRS_CHAIN_B_SLOTS 3
#endif
constexpr int kChainASlots = DRS_CHAIN_A_SLOTS;   // A ring: [128 x 16] hi + lo = 16 KB per slot
constexpr int kChainBSlots = DRS_CHAIN_B_SLOTS;   // B ring: [256 x 16] hi + lo = 32 KB per slot (3 slots: the
                                             // CTA must fit beside 2 persistent gather CTAs per SM)
constexpr int kChainABytes = kTcBM * kChainBK * 4;            // 8 KB (one of hi / lo)
constexpr int kChainBBytes = 256 * kChainBK * 4;              // 16 KB (one of hi / lo, BN = 256)
constexpr int kChainSmem = kChainASlots * 2 * kChainABytes + kChainBSlots * 2 * kChainBBytes +
                           1024 /*align*/ + 512 /*barriers*/ + kChainMaxLayers * 256 * 4 /*bias*/;

struct ChainParams {
  int n_layers;
  int M;                                     // valid rows
  int N[kChainMaxLayers], K[kChainMaxLayers], BN[kChainMaxLayers], act[kChainMaxLayers];
  const float* bias[kChainMaxLayers];
  float* tap[kChainMaxLayers];               // optional copy of hidden activations [M, N_l] (or null)
  int64_t ld_tap[kChainMaxLayers];
  float* Y;                                  // last layer output
  int64_t ldy;
  int split3;
};

// asynchronous 32-column TMEM load; pair with tmem_ld_wait() before touching r[]
__device__ __forceinline__ void tmem_ld_32x32b_x32_nowait(uint32_t taddr, uint32_t* r) {
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
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of (row, 16-byte chunk c) inside a [rows x 64 B] K-major SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64_offset(int row, int c) {
  return (uint32_t)(row * 64 + ((c ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;          // 8 rows x 64 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                   // SWIZZLE_64B
  return d;
}

// DRS_CHAIN_MIN_CTAS: min resident CTAs per SM the compiler must allow (register cap =
// 65536 / (192 * n)): 1 = unconstrained (~160 registers); 4 = 80 registers, which lets a chain
// CTA share an SM with three persistent gather CTAs (3 x 16,384 registers) at the price of
// spills in the epilogue (tools/build_variants.sh builds both for A/B runs).
#ifndef DRS_CHAIN_MIN_CTAS
#define DRS_CHAIN_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(kTcThreads, DRS_CHAIN_MIN_CTAS)
k_mlp_chain(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW0,
            const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
            const __grid_constant__ CUtensorMap mapW3, const ChainParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // (pointer arithmetic on the shared array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic loads and stores)
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* a_ring = base;                                              // slot: hi | lo
  unsigned char* b_ring = base + kChainASlots * 2 * kChainABytes;            // slot: hi | lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + kChainBSlots * 2 * kChainBBytes);
  uint64_t* a_full = bars;                                 // TMA landed X chunk (layer 0)
  uint64_t* a_ready = a_full + kChainASlots;               // hi/lo written (transform / epilogue)
  uint64_t* a_empty = a_ready + kChainASlots;              // MMA consumed
  uint64_t* b_full = a_empty + kChainASlots;
  uint64_t* b_empty = b_full + kChainBSlots;
  uint64_t* acc_full = b_empty + kChainBSlots;             // [kChainMaxLayers]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + kChainMaxLayers);
  float* sbias = reinterpret_cast<float*>(bars) + 128;      // [kChainMaxLayers][256], 512 B past bars

  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int m0 = blockIdx.x * kTcBM;
  for (int i = threadIdx.x; i < kChainMaxLayers * 256; i += kTcThreads) {
    const int l = i >> 8, n = i & 255;
    sbias[i] = (l < p.n_layers && n < p.N[l]) ? __ldg(p.bias[l] + n) : 0.f;
  }
  const CUtensorMap* mapW[kChainMaxLayers] = {&mapW0, &mapW1, &mapW2, &mapW3};

  if (threadIdx.x == 0) {
    for (int s = 0; s < kChainASlots; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_ready[s], 4);             // one arrival per transform/epilogue warp
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kChainBSlots; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int l = 0; l < kChainMaxLayers; ++l) mbar_init(&acc_full[l], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: X chunks (layer 0) and W chunks (all layers), in consumption order
    if (lane == 0) {
      int ac = 0, bc = 0;                    // global chunk counters of the two rings
      for (int l = 0; l < p.n_layers; ++l) {
        const int nk = (p.K[l] + kChainBK - 1) / kChainBK;
        const uint32_t btx = (uint32_t)p.BN[l] * 64u * (p.split3 ? 2u : 1u);
        for (int kc = 0; kc < nk; ++kc, ++bc) {
          if (l == 0) {
            const int s = ac % kChainASlots, u = ac / kChainASlots;
            if (u > 0) mbar_wait(&a_empty[s], (uint32_t)((u - 1) & 1));
            mbar_arrive_expect_tx(&a_full[s], (uint32_t)kChainABytes);
            tma_load_2d(a_ring + (size_t)s * 2 * kChainABytes, &mapX, kc * kChainBK, m0, &a_full[s]);
            ++ac;
          }
          const int s = bc % kChainBSlots, u = bc / kChainBSlots;
          if (u > 0) mbar_wait(&b_empty[s], (uint32_t)((u - 1) & 1));
          unsigned char* bs = b_ring + (size_t)s * 2 * kChainBBytes;
          mbar_arrive_expect_tx(&b_full[s], btx);
          tma_load_2d(bs, mapW[l], kc * kChainBK, 0, &b_full[s]);
          if (p.split3) tma_load_2d(bs + (size_t)p.BN[l] * 64, mapW[l], kc * kChainBK, p.N[l], &b_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if (lane == 0) {
      int ac = 0, bc = 0;
      for (int l = 0; l < p.n_layers; ++l) {
        const int nk = (p.K[l] + kChainBK - 1) / kChainBK;
        const uint32_t idesc = umma_idesc_tf32(p.BN[l]);
        const uint32_t tacc = tmem_base + (uint32_t)((l & 1) * 256);
        uint32_t acc = 0;
        for (int kc = 0; kc < nk; ++kc, ++ac, ++bc) {
          const int sa = ac % kChainASlots, sb = bc % kChainBSlots;
          mbar_wait(&a_ready[sa], (uint32_t)((ac / kChainASlots) & 1));
          mbar_wait(&b_full[sb], (uint32_t)((bc / kChainBSlots) & 1));
          tc_fence_after();
          const uint32_t a_hi = smem_u32(a_ring + (size_t)sa * 2 * kChainABytes);
          const uint32_t a_lo = a_hi + kChainABytes;
          const uint32_t b_hi = smem_u32(b_ring + (size_t)sb * 2 * kChainBBytes);
          const uint32_t b_lo = b_hi + (uint32_t)p.BN[l] * 64u;
#pragma unroll
          for (int k = 0; k < kChainBK / kTcUmmaK; ++k) {
            const uint32_t koff = (uint32_t)(k * kTcUmmaK * 4);
            if (p.split3) {
              umma_tf32(tacc, umma_desc_sw64(a_lo + koff), umma_desc_sw64(b_hi + koff), idesc, acc);
              acc = 1;
              umma_tf32(tacc, umma_desc_sw64(a_hi + koff), umma_desc_sw64(b_lo + koff), idesc, 1);
            }
            umma_tf32(tacc, umma_desc_sw64(a_hi + koff), umma_desc_sw64(b_hi + koff), idesc, acc);
            acc = 1;
          }
          umma_commit(&a_empty[sa]);
          umma_commit(&b_empty[sb]);
        }
        umma_commit(&acc_full[l]);
      }
    }
  } else {
    // ===== warps 2..5: layer-0 transform, inter-layer epilogue (produce next A chunks), output
    const int t = threadIdx.x - 2 * kWarp;           // 0..127
    const int q = warp % 4;                          // TMEM lane quarter
    const int row = q * 32 + lane;                   // row of the tile owned in TMEM
    int ac = 0;
    {  // layer 0: split the TMA-landed X chunk into hi / lo in place
      const int nk = (p.K[0] + kChainBK - 1) / kChainBK;
      for (int kc = 0; kc < nk; ++kc, ++ac) {
        const int s = ac % kChainASlots, u = ac / kChainASlots;
        mbar_wait(&a_full[s], (uint32_t)(u & 1));
        if (p.split3) {
          float4* hi = reinterpret_cast<float4*>(a_ring + (size_t)s * 2 * kChainABytes);
          float4* lo = reinterpret_cast<float4*>(a_ring + (size_t)s * 2 * kChainABytes + kChainABytes);
#pragma unroll
          for (int j = 0; j < kChainABytes / 16 / 128; ++j) {
            const int i = j * 128 + t;
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
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[s]);
      }
    }
    // layers 1..L-1: activations of layer l-1 -> A chunks of layer l.  32 accumulator columns
    // (two 16-wide A chunks) per step; the TMEM load of step i+1 is in flight while step i is
    // converted and written.
    for (int l = 1; l < p.n_layers; ++l) {
      const int pl = l - 1;
      mbar_wait(&acc_full[pl], 0);
      tc_fence_after();
      const uint32_t tsrc = tmem_base + (uint32_t)((pl & 1) * 256) + ((uint32_t)(q * 32) << 16);
      const float* __restrict__ bias = sbias + pl * 256;
      const int Np = p.N[pl], actp = p.act[pl];
      const int nk = (p.K[l] + kChainBK - 1) / kChainBK;     // K[l] == N[l-1]
      const int nsteps = (nk + 1) / 2;
      uint32_t rbuf[2][32];
      tmem_ld_32x32b_x32_nowait(tsrc, rbuf[0]);
#pragma unroll 1
      for (int st2 = 0; st2 < nsteps; st2 += 2) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int st = st2 + h2;
          if (st >= nsteps) break;
          uint32_t* r = rbuf[h2];
          tmem_ld_wait();
          if (st + 1 < nsteps) tmem_ld_32x32b_x32_nowait(tsrc + (uint32_t)((st + 1) * 32), rbuf[h2 ^ 1]);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = st * 32 + j;
            v[j] = n < Np ? apply_act(__uint_as_float(r[j]) + bias[n], actp) : 0.f;
          }
          if (p.tap[pl] != nullptr && m0 + row < p.M) {
            float* tp = p.tap[pl] + (int64_t)(m0 + row) * p.ld_tap[pl] + st * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (st * 32 + j < Np) tp[j] = v[j];
          }
          const int nch = (2 * st + 1 < nk) ? 2 : 1;         // A chunks produced by this step
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            if (hc >= nch) break;
            const int a = ac + hc;
            const int s = a % kChainASlots, u = a / kChainASlots;
            if (u > 0) mbar_wait(&a_empty[s], (uint32_t)((u - 1) & 1));
            unsigned char* hi = a_ring + (size_t)s * 2 * kChainABytes;
            unsigned char* lo = hi + kChainABytes;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float x0 = v[hc * 16 + 4 * c + 0], x1 = v[hc * 16 + 4 * c + 1];
              const float x2 = v[hc * 16 + 4 * c + 2], x3 = v[hc * 16 + 4 * c + 3];
              const uint32_t off = sw64_offset(row, c);
              if (p.split3) {
                float4 h, w;
                h.x = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
                h.y = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
                h.z = __uint_as_float(__float_as_uint(x2) & 0xFFFFE000u);
                h.w = __uint_as_float(__float_as_uint(x3) & 0xFFFFE000u);
                w.x = x0 - h.x; w.y = x1 - h.y; w.z = x2 - h.z; w.w = x3 - h.w;
                *reinterpret_cast<float4*>(hi + off) = h;
                *reinterpret_cast<float4*>(lo + off) = w;
              } else {
                *reinterpret_cast<float4*>(hi + off) = make_float4(x0, x1, x2, x3);
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&a_ready[ac % kChainASlots]);
            if (nch == 2) mbar_arrive(&a_ready[(ac + 1) % kChainASlots]);
          }
          ac += nch;
        }
      }
    }
    // output of the last layer
    {
      const int L = p.n_layers - 1;
      mbar_wait(&acc_full[L], 0);
      tc_fence_after();
      const uint32_t tsrc = tmem_base + (uint32_t)((L & 1) * 256) + ((uint32_t)(q * 32) << 16);
      const float* __restrict__ bias = sbias + L * 256;
      const int row0 = m0 + q * 32;
      // all MMAs have retired: ring memory is free -> per-warp 32 x 17 staging tile
      float* stg = reinterpret_cast<float*>(a_ring) + (size_t)(warp - 2) * 32 * 17;
      for (int c0 = 0; c0 < p.BN[L]; c0 += 16) {
        if (c0 >= p.N[L]) break;
        uint32_t r[16];
        tmem_ld_32x32b_x16(tsrc + (uint32_t)c0, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = __uint_as_float(r[j]);
        __syncwarp();
        // lane <-> (row parity, column): two rows of 16 columns per pass, 64-B coalesced
        const int cc = lane & 15, rsel = lane >> 4;
        const int n = c0 + cc;
        if (n < p.N[L]) {
          const float bn_ = bias[n];
          const int rmax = min(32, p.M - row0);
          for (int rr = rsel; rr < rmax; rr += 2)
            p.Y[(int64_t)(row0 + rr) * p.ldy + n] = apply_act(stg[rr * 17 + cc] + bn_, p.act[L]);
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
struct ChainLayerHost {
  const float* Wsplit;   // [2][N][K]
  const float* bias;
  int N, K, act;
  float* tap;            // optional hidden-activation copy (null for the last layer)
  int64_t ld_tap;
};

// tensor maps: tc_make_map_bk (mlp_tc.cuh), cached per (address, shape, pitch, box)

inline bool chain_supported(const ChainLayerHost* layers, int n, int64_t ldx, const float* X) {
  if (n < 1 || n > kChainMaxLayers || tc_encode_fn() == nullptr) return false;
  if (ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0) return false;
  for (int l = 0; l < n; ++l) {
    if (layers[l].N < 1 || layers[l].N > 256) return false;
    if (layers[l].K < 4 || layers[l].K % 4 != 0) return false;
    if (l > 0 && layers[l].K != layers[l - 1].N) return false;
  }
  return true;
}

// returns a cudaError_t value (0 = ok)
inline int chain_launch(const float* X, int64_t ldx, int64_t x_rows, const ChainLayerHost* layers,
                        int n, float* Y, int64_t ldy, int M, bool split3, cudaStream_t stream) {
  CUtensorMap mx, mw[kChainMaxLayers];
  if (!tc_make_map_bk(&mx, X, x_rows, layers[0].K, ldx, kTcBM, kChainBK))
    return (int)cudaErrorInvalidValue;
  ChainParams p{};
  p.n_layers = n; p.M = M; p.Y = Y; p.ldy = ldy; p.split3 = split3 ? 1 : 0;
  for (int l = 0; l < kChainMaxLayers; ++l) {
    const ChainLayerHost& h = layers[l < n ? l : n - 1];
    const int bn = (h.N + 15) & ~15;
    if (!tc_make_map_bk(&mw[l], h.Wsplit, 2 * (int64_t)h.N, h.K, h.K, bn, kChainBK))
      return (int)cudaErrorInvalidValue;
    if (l < n) {
      p.N[l] = h.N; p.K[l] = h.K; p.BN[l] = bn; p.act[l] = h.act;
      p.bias[l] = h.bias; p.tap[l] = (l < n - 1) ? h.tap : nullptr; p.ld_tap[l] = h.ld_tap;
    }
  }
  static std::atomic<uint64_t> attr{0};
  if (first_use_on_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_chain, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kChainSmem);
    if (e != cudaSuccess) return (int)e;
  }
  k_mlp_chain<<<ceil_div(M, kTcBM), kTcThreads, kChainSmem, stream>>>(mx, mw[0], mw[1], mw[2],
                                                                     mw[3], p);
  return (int)cudaGetLastError();
}

}  // namespace drs
