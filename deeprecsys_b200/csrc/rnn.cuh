// rnn.cuh — DIEN's interest-evolution layers (models/dien.py:296-380) on the GPU.
//
// The reference builds, out of Caffe2's rnn_cell.BasicRNN (tanh cell, forward only):
//     rnn_inputs = Concat(behaviour embeddings)          [B, N*D]
//     rnn_shape  = Reshape(rnn_inputs, (N, -1, D))        a raw reinterpretation of that buffer,
//                                                         NOT a transpose (dien.py:303-305)
//     h0_all = BasicRNN(rnn_shape, D -> H);   h1_last = BasicRNN(h0_all, H -> H)[last]
// with  h_t = tanh(FC(h_{t-1}; gates_t_w, gates_t_b) + FC(x_t; i2h_w, i2h_b)),  h_0 = 0, and
// seq_lengths == N for every row (dien.py:110-113, 513-518), so no step is masked.  (The
// FC + Softmax between the two cells is dead code in the reference: brew.sum(rnn_0_out, output)
// overwrites its result with rnn_0_out, dien.py:336-355.)
//
// Here: k_dien_seq builds the [N*B, D] sequence input from the pooled vectors with the same
// reinterpretation; the i2h FCs of all N steps run as ONE FC launch each ([N*B, in] x [in, H],
// mlp_ffma / mlp_tc); k_rnn_tanh runs the recurrence for a tile of rows with the H x H recurrent
// weights resident in shared memory.
#pragma once

#include "common.cuh"

namespace drs {

// seq[f] = pooled[b, (1 + n) * D + d]  with  f = ((t * B + b') * D + d')  read as
// b = f / (N*D), n = (f % (N*D)) / D, d = f % D   (tables 1..N are the behaviour tables)
__global__ void k_dien_seq(const float* __restrict__ pooled, int64_t ldp, float* __restrict__ seq,
                           int B, int N, int D) {
  const int64_t total = (int64_t)N * B * D;
  const int64_t nd = (int64_t)N * D;
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < total;
       f += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = f / nd;
    const int rem = (int)(f - b * nd);
    seq[f] = pooled[b * ldp + D + rem];          // (1 + n) * D + d == D + rem
  }
}

constexpr int kRnnThreads = 256;
constexpr int kRnnRows = 32;         // rows of the batch per CTA
constexpr int kRnnMaxH = 128;

inline size_t rnn_smem_bytes(int H) {
  return ((size_t)H * H + 2 * (size_t)kRnnRows * (H + 1)) * sizeof(float);
}

// pre [N, B, H] = i2h(x_t) incl. its bias;  Wg [H, H] (out, in) row-major, bg [H].
// all_out [N, B, H] (nullable) receives every h_t; last_out [B, ld_last] (nullable) receives h_N.
__global__ void __launch_bounds__(kRnnThreads) k_rnn_tanh(const float* __restrict__ pre,
                                                          const float* __restrict__ Wg,
                                                          const float* __restrict__ bg, int N,
                                                          int B, int H,
                                                          float* __restrict__ all_out,
                                                          float* __restrict__ last_out,
                                                          int64_t ld_last) {
  extern __shared__ __align__(16) float rnn_smem[];
  float* wT = rnn_smem;                                  // [k][j] = Wg[j][k]
  float* hbuf = rnn_smem + (size_t)H * H;                // 2 x [kRnnRows][H + 1]
  const int Hp = H + 1;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * kRnnRows;
  const int nrows = min(kRnnRows, B - b0);
  for (int i = tid; i < H * H; i += kRnnThreads) {
    const int j = i / H, k = i - j * H;
    wT[k * H + j] = __ldg(Wg + i);
  }
  for (int i = tid; i < 2 * kRnnRows * Hp; i += kRnnThreads) hbuf[i] = 0.f;   // h_0 = 0
  __syncthreads();

  const int r = tid / 8, c0 = tid % 8;                   // 32 rows x 8 column phases
  constexpr int kMaxCols = kRnnMaxH / 8;
  for (int t = 0; t < N; ++t) {
    const float* hc = hbuf + (size_t)(t & 1) * kRnnRows * Hp + (size_t)r * Hp;
    float* hn = hbuf + (size_t)((t + 1) & 1) * kRnnRows * Hp + (size_t)r * Hp;
    float acc[kMaxCols];
#pragma unroll
    for (int i = 0; i < kMaxCols; ++i) acc[i] = 0.f;
    for (int k = 0; k < H; ++k) {                        // ascending k, fp32
      const float hv = hc[k];
      const float* wrow = wT + (size_t)k * H + c0;
#pragma unroll
      for (int i = 0; i < kMaxCols; ++i)
        if (c0 + 8 * i < H) acc[i] = fmaf(hv, wrow[8 * i], acc[i]);
    }
    if (r < nrows) {
      const int64_t row = (int64_t)t * B + b0 + r;
#pragma unroll
      for (int i = 0; i < kMaxCols; ++i) {
        const int j = c0 + 8 * i;
        if (j < H) {
          // FC(h_{t-1}) = h.Wg^T + bg, then Sum with the i2h term, then Tanh
          const float v = tanhf((acc[i] + __ldg(bg + j)) + __ldg(pre + row * H + j));
          hn[j] = v;
          if (all_out) all_out[row * H + j] = v;
          if (last_out && t == N - 1) last_out[(int64_t)(b0 + r) * ld_last + j] = v;
        }
      }
    }
    __syncthreads();
  }
}

// Register-tiled variant for H % 4 == 0 (the shipped configs: H = 64): 64 rows per CTA, each
// thread owns 4 rows x 4 adjacent columns per 64-column chunk, so one k-step costs four broadcast
// reads of h and one 16-byte read of the weights for 16 FMAs (the generic kernel above needs nine
// shared-memory reads for eight).
constexpr int kRnnRows4 = 64;

inline size_t rnn4_smem_bytes(int H) {
  return ((size_t)H * H + 2 * (size_t)kRnnRows4 * (H + 4)) * sizeof(float);
}

__global__ void __launch_bounds__(kRnnThreads) k_rnn_tanh4(const float* __restrict__ pre,
                                                           const float* __restrict__ Wg,
                                                           const float* __restrict__ bg, int N,
                                                           int B, int H,
                                                           float* __restrict__ all_out,
                                                           float* __restrict__ last_out,
                                                           int64_t ld_last) {
  extern __shared__ __align__(16) float rnn_smem[];
  float* wT = rnn_smem;                                  // [k][j] = Wg[j][k]
  float* hbuf = rnn_smem + (size_t)H * H;                // 2 x [kRnnRows4][H + 4]
  const int Hp = H + 4;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * kRnnRows4;
  const int nrows = min(kRnnRows4, B - b0);
  for (int i = tid; i < H * H; i += kRnnThreads) {
    const int j = i / H, k = i - j * H;
    wT[k * H + j] = __ldg(Wg + i);
  }
  for (int i = tid; i < 2 * kRnnRows4 * Hp; i += kRnnThreads) hbuf[i] = 0.f;   // h_0 = 0
  __syncthreads();

  const int r0 = (tid / 16) * 4, c0 = (tid % 16) * 4;    // 16 row groups x 16 column groups
  constexpr int kChunks = kRnnMaxH / 64;
  for (int t = 0; t < N; ++t) {
    const float* hc = hbuf + (size_t)(t & 1) * kRnnRows4 * Hp + (size_t)r0 * Hp;
    float* hn = hbuf + (size_t)((t + 1) & 1) * kRnnRows4 * Hp + (size_t)r0 * Hp;
    float acc[kChunks][4][4];
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][i][j] = 0.f;
    for (int k = 0; k < H; ++k) {                        // ascending k, fp32
      float hv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) hv[i] = hc[i * Hp + k];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        if (c0 + 64 * c < H) {
          const float4 w = *reinterpret_cast<const float4*>(wT + (size_t)k * H + c0 + 64 * c);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[c][i][0] = fmaf(hv[i], w.x, acc[c][i][0]);
            acc[c][i][1] = fmaf(hv[i], w.y, acc[c][i][1]);
            acc[c][i][2] = fmaf(hv[i], w.z, acc[c][i][2]);
            acc[c][i][3] = fmaf(hv[i], w.w, acc[c][i][3]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int j0 = c0 + 64 * c;
      if (j0 < H) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bg + j0));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (r0 + i < nrows) {
            const int64_t row = (int64_t)t * B + b0 + r0 + i;
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(pre + row * H + j0));
            float4 v;
            v.x = tanhf((acc[c][i][0] + bv.x) + p4.x);
            v.y = tanhf((acc[c][i][1] + bv.y) + p4.y);
            v.z = tanhf((acc[c][i][2] + bv.z) + p4.z);
            v.w = tanhf((acc[c][i][3] + bv.w) + p4.w);
            *reinterpret_cast<float4*>(hn + i * Hp + j0) = v;
            if (all_out) *reinterpret_cast<float4*>(all_out + row * H + j0) = v;
            if (last_out && t == N - 1) {
              float* o = last_out + (int64_t)(b0 + r0 + i) * ld_last + j0;
              o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace drs
