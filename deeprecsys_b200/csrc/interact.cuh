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
