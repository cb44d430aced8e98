// interact.cuh — the small glue operators of the sibling graphs that are not
// absorbed by K1's output layout.  All are strided elementwise kernels.
//
//   k_copy_cols    Concat legs that are plain copies (WnD dense_in -> R,
//                  wide_and_deep.py:271-279; DIN profile/ad/context -> top input,
//                  din.py:319-325)
//   k_add_cols     Sum(e0, e1): NCF "MF" branch (ncf.py:303)
//   k_din_concat   per behaviour table u: [e_u, e_ad, e_u + e_ad]  (din.py:263-273)
//   k_sum_units    Sum over attention-unit outputs, in unit order (din.py:283)
//   k_stack_rows   coalesced serving: the dense rows of several requests (each "the first
//                  size_k rows of staged batch b_k", inferenceEngine.py:213) -> one stacked
//                  [sum size_k, m_den] matrix, so that ONE MLP launch serves all of them
#pragma once

#include "common.cuh"

namespace drs {

__global__ void k_copy_cols(const float* __restrict__ src, int64_t lds, float* __restrict__ dst,
                            int64_t ldd, int B, int ncols) {
  const int64_t n = (int64_t)B * ncols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ncols, c = i - b * ncols;
    dst[b * ldd + c] = src[b * lds + c];
  }
}

__global__ void k_add_cols(const float* __restrict__ a, const float* __restrict__ b2, int64_t lds,
                           float* __restrict__ dst, int64_t ldd, int B, int ncols) {
  const int64_t n = (int64_t)B * ncols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ncols, c = i - b * ncols;
    dst[b * ldd + c] = a[b * lds + c] + b2[b * lds + c];
  }
}

// pooled: [B, T*D]; cat: n_units blocks of [B, 3D], unit_stride floats apart; unit i uses
// behaviour table 1 + i, ad table = ad_t.
__global__ void k_din_concat(const float* __restrict__ pooled, int64_t ldp, float* __restrict__ cat,
                             int64_t unit_stride, int B, int D, int n_units, int ad_t) {
  const int64_t n = (int64_t)n_units * B * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = i / ((int64_t)B * D);
    const int64_t r = i - u * (int64_t)B * D;
    const int64_t b = r / D, c = r - b * D;
    const float eu = pooled[b * ldp + (1 + u) * D + c];
    const float ea = pooled[b * ldp + (int64_t)ad_t * D + c];
    float* o = cat + u * unit_stride + b * (3 * (int64_t)D);
    o[c] = eu;
    o[D + c] = ea;
    o[2 * D + c] = eu + ea;
  }
}

// units: [n_units] x [B, D] blocks unit_stride floats apart -> dst[b, c] = ((u0 + u1) + u2) + ...   (Caffe2 Sum is left-to-right)
__global__ void k_sum_units(const float* __restrict__ units, int64_t unit_stride,
                            float* __restrict__ dst, int64_t ldd, int B, int D, int n_units) {
  const int64_t n = (int64_t)B * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / D, c = i - b * D;
    float acc = units[i];
    for (int u = 1; u < n_units; ++u) acc += units[(int64_t)u * unit_stride + i];
    dst[b * ldd + c] = acc;
  }
}

// Unfused exchange baseline: recv = [world][S][tpr*D] blocks (block q = this rank's samples x
// the tables rank q owns, table q + j*world at columns j*D) -> out[row, col0 + t*D ...].
__global__ void k_unpack_exchange(const float* __restrict__ recv, int S, int tpr, int world, int T,
                                  int D, int rows, float* __restrict__ out, int64_t ld_out,
                                  int col0) {
  const int64_t n = (int64_t)rows * T * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / ((int64_t)T * D);
    const int r = (int)(i - row * (int64_t)T * D);
    const int t = r / D, c = r - t * D;
    const int q = t % world, j = t / world;
    out[row * ld_out + col0 + (int64_t)t * D + c] =
        recv[((int64_t)q * S + row) * ((int64_t)tpr * D) + (int64_t)j * D + c];
  }
}

// ---------------------------------------------------------------- fused DIN attention
// din.py:246-331 for one sample per warp, all behaviour tables in one pass over the pooled
// vectors: for every unit u (own weights each)
//     c   = Concat(e_u, e_ad, Sum(e_u, e_ad))                    [3D]     din.py:263-273
//     o_u = MLP_u(c), layers [3D] + arch_mlp_bot + [D], all ReLU           :253-258, :278
// atten = Sum_u o_u (left to right, :283), and the top-MLP input
//     R = Concat(e_profile, atten, e_ad, e_ctx)                  [4D]     :319-325
// Replaces k_din_concat + n_layers batched FC launches + k_sum_units + three k_copy_cols and the
// [n_units, B, 3D] / [n_units, B, h] intermediates they pass through HBM (DIN with 251 behaviour
// tables: ~1.3 GB per 4,096 samples) by one read of the pooled vectors.
// Lane j owns column j (D <= 32).  Layer 0 is a distributed dot product (each lane multiplies
// its three elements of c, warp-reduced: fp32 reassociation only); narrower layers keep their
// activations one per lane and broadcast them by shuffle.  Hidden widths <= 32.
constexpr int kDinMaxLayers = 4;
constexpr int kDinPrefetch = 8;             // units whose e_u load is in flight per sample
struct DinAttnParams {
  const float* P;  int64_t ldp;             // pooled [rows, T*D]
  float* R;  int64_t ldr;                   // top-MLP input [rows, 4*D]
  int rows, D, T, n_units, n_layers;
  int width[kDinMaxLayers + 1];             // 3D, hidden..., D
  const float* W[kDinMaxLayers];            // [n_units][out][in]
  const float* b[kDinMaxLayers];            // [n_units][out]
};

// U: units processed together (independent dependency chains interleaved); 1 for models with a
// handful of behaviour tables, kDinPrefetch for long behaviour sequences.
template <int U>
__global__ void __launch_bounds__(256) k_din_attention(const DinAttnParams p) {
  const int lane = threadIdx.x % kWarp;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int nwarps = (gridDim.x * blockDim.x) / kWarp;
  const int D = p.D, nl = p.n_layers;
  const bool col = lane < D;
  for (int s = warp; s < p.rows; s += nwarps) {
    const float* prow = p.P + (int64_t)s * p.ldp;
    // pooled vectors are read once: keep them out of L1, where the units' weights (re-read by
    // every sample) should stay
    const float ead = col ? ldg_stream_f1(prow + (int64_t)(p.T - 2) * D + lane) : 0.f;
    float atten = 0.f;
    float pre[U];
#pragma unroll
    for (int i = 0; i < U; ++i)
      pre[i] = (col && i < p.n_units) ? ldg_stream_f1(prow + (int64_t)(1 + i) * D + lane) : 0.f;
    for (int u0 = 0; u0 < p.n_units; u0 += U) {
      float eu[U], act[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        eu[i] = pre[i];
        const int un = u0 + i + U;
        pre[i] = (col && un < p.n_units) ? ldg_stream_f1(prow + (int64_t)(1 + un) * D + lane) : 0.f;
        act[i] = 0.f;
      }
      if (nl == 2) {
        // the shipped shape [3D, h, D] (din.json: h = 1): per-unit pointers once, 32-bit offsets
        const int h = p.width[1];
        const float* w0[U]; const float* w1[U]; const float* bb0[U];
        float acc[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int u = min(u0 + i, p.n_units - 1);
          w0[i] = p.W[0] + (int64_t)u * h * (3 * D) + lane;
          bb0[i] = p.b[0] + (int64_t)u * h;
          w1[i] = p.W[1] + ((int64_t)u * D + lane) * h;
          acc[i] = col ? __ldg(p.b[1] + (int64_t)u * D + lane) : 0.f;
        }
        for (int k = 0; k < h; ++k) {
          float part[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            part[i] = 0.f;
            if (col) {
              const float* w = w0[i] + k * 3 * D;
              part[i] = fmaf(__ldg(w), eu[i], part[i]);
              part[i] = fmaf(__ldg(w + D), ead, part[i]);
              part[i] = fmaf(__ldg(w + 2 * D), eu[i] + ead, part[i]);
            }
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
            for (int i = 0; i < U; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
#pragma unroll
          for (int i = 0; i < U; ++i) {          // every lane holds the hidden activation k
            const float a = fmaxf(part[i] + __ldg(bb0[i] + k), 0.f);
            if (col) acc[i] = fmaf(__ldg(w1[i] + k), a, acc[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < U; ++i)
          if (u0 + i < p.n_units) atten += fmaxf(acc[i], 0.f);     // Sum over units, in unit order
        continue;
      }
      // layer 0: [3D] -> width[1] for the U units at once; lane k ends up holding activation k
      {
        const int out = p.width[1], in = p.width[0];
        for (int k = 0; k < out; ++k) {
          float part[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const int u = min(u0 + i, p.n_units - 1);          // (clamped: result unused past the end)
            const float* w = p.W[0] + ((int64_t)u * out + k) * in;
            part[i] = 0.f;
            if (col) {
              part[i] = fmaf(__ldg(w + lane), eu[i], part[i]);
              part[i] = fmaf(__ldg(w + D + lane), ead, part[i]);
              part[i] = fmaf(__ldg(w + 2 * D + lane), eu[i] + ead, part[i]);
            }
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
            for (int i = 0; i < U; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], o);
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const int u = min(u0 + i, p.n_units - 1);
            const float a = fmaxf(part[i] + __ldg(p.b[0] + (int64_t)u * out + k), 0.f);
            if (lane == k) act[i] = a;
          }
        }
      }
      // layers 1 .. nl-1: lane j computes output j from the broadcast activations
      for (int l = 1; l < nl; ++l) {
        const int in = p.width[l], out = p.width[l + 1];
        float acc[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int u = min(u0 + i, p.n_units - 1);
          acc[i] = lane < out ? __ldg(p.b[l] + (int64_t)u * out + lane) : 0.f;
        }
        for (int k = 0; k < in; ++k) {
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const int u = min(u0 + i, p.n_units - 1);
            const float ak = __shfl_sync(0xffffffffu, act[i], k);
            if (lane < out)
              acc[i] = fmaf(__ldg(p.W[l] + ((int64_t)u * out + lane) * in + k), ak, acc[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < U; ++i) act[i] = lane < out ? fmaxf(acc[i], 0.f) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < U; ++i)
        if (u0 + i < p.n_units) atten += act[i];               // Sum over units, in unit order
    }
    if (col) {
      float* r = p.R + (int64_t)s * p.ldr;
      r[lane] = __ldg(prow + lane);                                   // user profile (table 0)
      r[D + lane] = atten;
      r[2 * D + lane] = ead;                                          // candidate ad (table T-2)
      r[3 * D + lane] = __ldg(prow + (int64_t)(p.T - 1) * D + lane);  // context (table T-1)
    }
  }
}

constexpr int kStackMaxSrc = 32;
struct StackRows {
  int n;                                  // requests
  int ncols;                              // floats per row
  int row_off[kStackMaxSrc + 1];          // first stacked row of request k; row_off[n] = total rows
  const float* src[kStackMaxSrc];         // [size_k, ncols] (row pitch = ncols)
};

__global__ void k_stack_rows(const StackRows p, float* __restrict__ dst) {
  const int64_t n = (int64_t)p.row_off[p.n] * p.ncols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / p.ncols);
    int k = 0;
    while (k + 1 < p.n && row >= p.row_off[k + 1]) ++k;
    dst[i] = __ldg(p.src[k] + (i - (int64_t)p.row_off[k] * p.ncols));
  }
}

}  // namespace drs
