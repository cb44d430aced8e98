// rnn_tc.cuh — DIEN's tanh recurrence on tcgen05 (sm_100a), one BasicRNN cell per launch.
//
//   h_t = tanh( FC(h_{t-1}; gates_t_w, gates_t_b) + FC(x_t; i2h_w, i2h_b) ),  h_0 = 0
// (caffe2 rnn_cell.BasicRNN as called at /root/reference/models/dien.py:318-328, 370-378).
// Round 1 ran the input FC of all N steps as one launch and the recurrence as an FFMA kernel
// (csrc/rnn.cuh): 2 x 412 us + 2 x 110 us of a 1.2 ms DIEN step
// (profiles/r02g_launches_dien.csv).  Here both products of a step go to the tensor core and
// into ONE TMEM accumulator:
//     acc = x_t . Wi^T  +  h_{t-1} . Wg^T        (kind::tf32, 3xTF32 error-compensated)
//     h_t = tanh(acc + bi + bg)
// for a tile of 128 samples per CTA, the pre-activation never touching HBM.
//
//   * Wi and Wg (hi / lo parts, K-major SWIZZLE_64B chunks of 16 floats) are loaded by TMA once
//     and stay in shared memory for all N steps;
//   * x_t tiles [128 x Kin] stream through a small TMA ring and are split hi / lo in place;
//   * h_t is produced by the epilogue warps straight from TMEM (tcgen05.ld -> +bias -> tanhf ->
//     hi / lo split) into the swizzled A-operand tile of the next step — the recurrence never
//     leaves the SM; cell 0 also stores every h_t (cell 1's input), cell 1 only its last state.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 x-split +
// epilogue: thread <-> sample row (TMEM lane quarter = warp % 4), the two warps of a quarter
// take alternate 16-column chunks.  The epilogue is the step's critical path (64 tanh per
// row and step between two dependent MMAs), hence two warps per scheduler, biases in shared
// memory, tanh as 1 - 2 / (exp(2x) + 1) on the SFU (absolute error ~1e-7, below the 3xTF32
// products'), and h_t stored to HBM from its shared-memory tile in 64-byte row pieces instead of
// 16 bytes per lane.
// Shapes: H and Kin multiples of 16, 16 <= H <= 128.  Other shapes, and the fp32 parity mode,
// keep the FFMA kernels of rnn.cuh.
#pragma once

#include "mlp_chain.cuh"

namespace drs {

constexpr int kRnnTcBK = 16;                         // floats per K chunk (64-byte swizzle row)
constexpr int kRnnTcABytes = kTcBM * kRnnTcBK * 4;   // one 128-row chunk tile (hi or lo): 8 KB
constexpr int kRnnTcThreads = 320;
constexpr int kRnnTcEpiWarps = 8;

__device__ __forceinline__ float tanh_sfu(float x) {
  // tanh(x) = 1 - 2 / (e^{2x} + 1); clamped so that e^{2x} stays finite
  const float e = __expf(2.f * fminf(fmaxf(x, -15.f), 15.f));
  return 1.f - __fdividef(2.f, e + 1.f);
}

struct RnnTcParams {
  int N, B, H, Kin, XS;                 // steps, rows, hidden width, input width, x-ring slots
  const float* bias_i;                  // [H] i2h bias
  const float* bias_g;                  // [H] gates_t bias
  float* all_out;                       // [N*B, H] every h_t (or nullptr)
  float* last_out;                      // [B, ld_last] h_N (or nullptr)
  int64_t ld_last;
  int split3;
};

__host__ __device__ inline size_t rnn_tc_smem_bytes(int H, int Kin, int XS) {
  const size_t wi = (size_t)(Kin / kRnnTcBK) * 2 * H * 64;
  const size_t wg = (size_t)(H / kRnnTcBK) * 2 * H * 64;
  const size_t ht = (size_t)(H / kRnnTcBK) * 2 * kRnnTcABytes;
  const size_t xr = (size_t)XS * (Kin / kRnnTcBK) * 2 * kRnnTcABytes;
  return wi + wg + ht + xr + 1024 /*align*/ + 256 /*barriers*/ + 512 /*bias*/;
}

__global__ void __launch_bounds__(kRnnTcThreads, 1)
k_rnn_tc(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWi,
         const __grid_constant__ CUtensorMap mapWg, const RnnTcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // (pointer arithmetic on the shared array, not an integer round trip: the compiler keeps the
  // shared address space and emits LDS / STS instead of generic loads and stores)
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int H = p.H, nkx = p.Kin / kRnnTcBK, nkh = H / kRnnTcBK, XS = p.XS;
  const uint32_t wtile = (uint32_t)H * 64u;                 // one weight chunk tile (hi or lo)
  unsigned char* WI = base;                                 // [nkx][hi | lo]
  unsigned char* WG = WI + (size_t)nkx * 2 * wtile;         // [nkh][hi | lo]
  unsigned char* HT = WG + (size_t)nkh * 2 * wtile;         // [nkh][hi | lo] h_{t-1}, 128 rows
  unsigned char* XR = HT + (size_t)nkh * 2 * kRnnTcABytes;  // [XS][nkx][hi | lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(XR + (size_t)XS * nkx * 2 * kRnnTcABytes);
  uint64_t* w_full = bars;
  uint64_t* acc_full = bars + 1;
  uint64_t* h_ready = bars + 2;
  uint64_t* x_full = bars + 3;                              // [XS]
  uint64_t* x_ready = x_full + 4;                           // [XS]  (XS <= 4)
  uint64_t* x_empty = x_ready + 4;                          // [XS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + 4);
  float* sbias = reinterpret_cast<float*>(bars) + 64;       // [H] bi + bg, 256 B past the barriers

  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int b0 = blockIdx.x * kTcBM;
  const uint32_t tmem_cols = H <= 32 ? 32u : (H <= 64 ? 64u : 128u);

  for (int i = threadIdx.x; i < H; i += kRnnTcThreads)
    sbias[i] = __ldg(p.bias_g + i) + __ldg(p.bias_i + i);
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(h_ready, kRnnTcEpiWarps);
    for (int s = 0; s < XS; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_ready[s], kRnnTcEpiWarps);
      mbar_init(&x_empty[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: weights once, then the x tiles of every step
    if (lane == 0) {
      const uint32_t wbytes = (uint32_t)(nkx + nkh) * (p.split3 ? 2u : 1u) * wtile;
      mbar_arrive_expect_tx(w_full, wbytes);
      for (int kc = 0; kc < nkx; ++kc) {
        tma_load_2d(WI + (size_t)kc * 2 * wtile, &mapWi, kc * kRnnTcBK, 0, w_full);
        if (p.split3)
          tma_load_2d(WI + (size_t)kc * 2 * wtile + wtile, &mapWi, kc * kRnnTcBK, H, w_full);
      }
      for (int kc = 0; kc < nkh; ++kc) {
        tma_load_2d(WG + (size_t)kc * 2 * wtile, &mapWg, kc * kRnnTcBK, 0, w_full);
        if (p.split3)
          tma_load_2d(WG + (size_t)kc * 2 * wtile + wtile, &mapWg, kc * kRnnTcBK, H, w_full);
      }
      for (int t = 0; t < p.N; ++t) {
        const int s = t % XS, u = t / XS;
        if (u > 0) mbar_wait(&x_empty[s], (uint32_t)((u - 1) & 1));
        unsigned char* xs = XR + (size_t)s * nkx * 2 * kRnnTcABytes;
        mbar_arrive_expect_tx(&x_full[s], (uint32_t)nkx * kRnnTcABytes);
        for (int kc = 0; kc < nkx; ++kc)
          tma_load_2d(xs + (size_t)kc * 2 * kRnnTcABytes, &mapX, kc * kRnnTcBK, t * p.B + b0,
                      &x_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: acc = x_t . Wi^T (+ h_{t-1} . Wg^T for t > 0)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(H);
      mbar_wait(w_full, 0);
      for (int t = 0; t < p.N; ++t) {
        const int s = t % XS, u = t / XS;
        mbar_wait(&x_ready[s], (uint32_t)(u & 1));
        if (t > 0) mbar_wait(h_ready, (uint32_t)((t - 1) & 1));
        tc_fence_after();
        uint32_t acc = 0;
        const uint32_t xs = smem_u32(XR + (size_t)s * nkx * 2 * kRnnTcABytes);
        auto chunk = [&](uint32_t a_hi, uint32_t b_hi) {
          const uint32_t a_lo = a_hi + kRnnTcABytes, b_lo = b_hi + wtile;
#pragma unroll
          for (int k = 0; k < kRnnTcBK / kTcUmmaK; ++k) {
            const uint32_t koff = (uint32_t)(k * kTcUmmaK * 4);
            if (p.split3) {
              umma_tf32(tmem_base, umma_desc_sw64(a_lo + koff), umma_desc_sw64(b_hi + koff), idesc, acc);
              acc = 1;
              umma_tf32(tmem_base, umma_desc_sw64(a_hi + koff), umma_desc_sw64(b_lo + koff), idesc, 1);
            }
            umma_tf32(tmem_base, umma_desc_sw64(a_hi + koff), umma_desc_sw64(b_hi + koff), idesc, acc);
            acc = 1;
          }
        };
        for (int kc = 0; kc < nkx; ++kc)
          chunk(xs + (uint32_t)kc * 2u * kRnnTcABytes, smem_u32(WI) + (uint32_t)kc * 2u * wtile);
        if (t > 0)
          for (int kc = 0; kc < nkh; ++kc)
            chunk(smem_u32(HT) + (uint32_t)kc * 2u * kRnnTcABytes,
                  smem_u32(WG) + (uint32_t)kc * 2u * wtile);
        umma_commit(&x_empty[s]);          // the x slot may be refilled when these MMAs retire
        umma_commit(acc_full);
      }
    }
  } else {
    // ===== warps 2..9: split x tiles hi / lo, then per step the epilogue that produces h_t
    const int tt = threadIdx.x - 2 * kWarp;          // 0..255
    const int q = warp % 4;
    const int half = (warp - 2) / 4;                 // which 16-column chunks of the row: c16 % 2
    const int row = q * 32 + lane;                   // TMEM lane == sample row of the tile
    auto split_x = [&](int t) {
      const int s = t % XS, u = t / XS;
      mbar_wait(&x_full[s], (uint32_t)(u & 1));
      if (p.split3) {
        for (int kc = 0; kc < nkx; ++kc) {
          float4* hi = reinterpret_cast<float4*>(XR + ((size_t)s * nkx + kc) * 2 * kRnnTcABytes);
          float4* lo = hi + kRnnTcABytes / 16;
#pragma unroll
          for (int j = 0; j < kRnnTcABytes / 16 / 256; ++j) {
            const int i = j * 256 + tt;
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
        }
        fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_ready[s]);
    };
    split_x(0);
    for (int t = 0; t < p.N; ++t) {
      mbar_wait(acc_full, (uint32_t)(t & 1));
      tc_fence_after();
      const bool last = t == p.N - 1;
      for (int c16 = half; c16 < nkh; c16 += 2) {     // this warp's 16-column chunks
        const int c0 = c16 * 16;
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tanh_sfu(__uint_as_float(r[j]) + sbias[c0 + j]);
        if (p.last_out && last && b0 + row < p.B) {
          float* o = p.last_out + (int64_t)(b0 + row) * p.ld_last + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = v[j];
        }
        // h_t chunk -> swizzled A-operand tile of step t+1 (also the staging for the HBM store)
        unsigned char* hi = HT + (size_t)c16 * 2 * kRnnTcABytes;
        unsigned char* lo = hi + kRnnTcABytes;
        if (!last || p.all_out) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float x0 = v[4 * c + 0], x1 = v[4 * c + 1], x2 = v[4 * c + 2], x3 = v[4 * c + 3];
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
        if (p.all_out) {
          // every h_t to HBM (cell 1's input): read this warp's 32 rows x 64 bytes back from
          // the tile, 4 lanes per row piece -> full 32-byte sectors, 8 row pieces per request
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = q * 32 + it * 8 + (lane >> 2), c = lane & 3;
            const uint32_t off = sw64_offset(rr, c);
            float4 x = *reinterpret_cast<const float4*>(hi + off);
            if (p.split3) {
              const float4 y = *reinterpret_cast<const float4*>(lo + off);
              x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;       // hi + lo == h exactly
            }
            if (b0 + rr < p.B)
              *reinterpret_cast<float4*>(p.all_out + ((int64_t)t * p.B + b0 + rr) * H + c0 +
                                         4 * c) = x;
          }
        }
      }
      if (!last) {
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_ready);
        split_x(t + 1);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

inline bool rnn_tc_supported(int H, int Kin) {
  return H >= 16 && H <= 128 && H % 16 == 0 && Kin >= 16 && Kin % 16 == 0 && Kin <= 128 &&
         tc_encode_fn() != nullptr;
}

// X: [N*B, Kin] (row pitch ldx); Wi_split [2H, Kin], Wg_split [2H, H] (hi rows, then lo rows).
// returns a cudaError_t value (0 = ok)
inline int rnn_tc_launch(const float* X, int64_t ldx, const float* Wi_split, const float* Wg_split,
                         const float* bi, const float* bg, int N, int B, int H, int Kin,
                         float* all_out, float* last_out, int64_t ld_last, bool split3,
                         cudaStream_t stream) {
  CUtensorMap mx, mwi, mwg;
  if (!tc_make_map_bk(&mx, X, (int64_t)N * B, Kin, ldx, kTcBM, kRnnTcBK) ||
      !tc_make_map_bk(&mwi, Wi_split, 2 * (int64_t)H, Kin, Kin, H, kRnnTcBK) ||
      !tc_make_map_bk(&mwg, Wg_split, 2 * (int64_t)H, H, H, H, kRnnTcBK))
    return (int)cudaErrorInvalidValue;
  RnnTcParams p{};
  p.N = N; p.B = B; p.H = H; p.Kin = Kin;
  p.XS = 4;
  while (p.XS > 1 && rnn_tc_smem_bytes(H, Kin, p.XS) > 200 * 1024) --p.XS;
  if (rnn_tc_smem_bytes(H, Kin, p.XS) > 220 * 1024) return (int)cudaErrorInvalidValue;
  p.bias_i = bi; p.bias_g = bg;
  p.all_out = all_out; p.last_out = last_out; p.ld_last = ld_last;
  p.split3 = split3 ? 1 : 0;
  static std::atomic<uint64_t> attr{0};
  if (first_use_on_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(k_rnn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         220 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  k_rnn_tc<<<ceil_div(B, kTcBM), kRnnTcThreads, rnn_tc_smem_bytes(H, Kin, p.XS), stream>>>(
      mx, mwi, mwg, p);
  return (int)cudaGetLastError();
}

}  // namespace drs
