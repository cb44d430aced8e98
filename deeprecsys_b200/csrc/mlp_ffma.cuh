// mlp_ffma.cuh — K3 (parity anchor): FC + bias + Relu/Sigmoid in plain fp32 FFMA.
//
// Replaces   FC([x, W, b]) -> Relu | Sigmoid
//   /root/reference/models/dlrm_s_caffe2.py:258-272 (and the create_mlp of
//   wide_and_deep.py / ncf.py / din.py): Y = X . W^T + b with W [out, in]
//   row-major, then max(0, y) or 1 / (1 + exp(-y)).
//
// fp32 accumulate over k in ascending order per output element (deterministic).
// Strided X/Y views let a layer read from / write into a column block of a
// wider activation buffer (that is how Concat disappears), and grid.z batches
// independent problems with their own weights (DIN attention units).
// The tensor-core path for the wide layers lives in mlp_tc.cuh.
#pragma once

#include "common.cuh"

namespace drs {

enum { kActRelu = 0, kActSigmoid = 1, kActNone = 2 };

struct FcLaunch {
  const float* X;  int64_t ldx;  int64_t x_bstride;
  const float* W;  int64_t w_bstride;            // [N, K], ldw == K
  const float* bias;  int64_t b_bstride;
  float* Y;  int64_t ldy;  int64_t y_bstride;
  int M, N, K;
  int act;
};

__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == kActRelu) return fmaxf(y, 0.f);
  if (act == kActSigmoid) return 1.f / (1.f + expf(-y));
  return y;
}

constexpr int kFcBM = 64, kFcBN = 64, kFcBK = 16, kFcThreads = 256, kFcPad = 4;

template <bool VEC>
__global__ void __launch_bounds__(kFcThreads) k_fc_ffma(const FcLaunch p) {
  __shared__ __align__(16) float As[kFcBK][kFcBM + kFcPad];
  __shared__ __align__(16) float Bs[kFcBK][kFcBN + kFcPad];

  const int z = blockIdx.z;
  const float* __restrict__ X = p.X + z * p.x_bstride;
  const float* __restrict__ W = p.W + z * p.w_bstride;
  const float* __restrict__ bias = p.bias + z * p.b_bstride;
  float* __restrict__ Y = p.Y + z * p.y_bstride;

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kFcBM, n0 = blockIdx.x * kFcBN;
  const int tx = tid % 16, ty = tid / 16;       // 16 x 16 threads, 4 x 4 outputs each
  const int lrow = tid / 4, lk = (tid % 4) * 4; // loader: 64 rows x 4 float4

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += kFcBK) {
    float a[4], b[4];
    {
      const int m = m0 + lrow, k = k0 + lk;
      if (VEC && m < p.M && k + 3 < p.K) {
        const float4 v = *reinterpret_cast<const float4*>(X + (int64_t)m * p.ldx + k);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          a[i] = (m < p.M && k + i < p.K) ? X[(int64_t)m * p.ldx + k + i] : 0.f;
      }
      const int n = n0 + lrow;
      if (VEC && n < p.N && k + 3 < p.K) {
        const float4 v = *reinterpret_cast<const float4*>(W + (int64_t)n * p.K + k);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          b[i] = (n < p.N && k + i < p.K) ? W[(int64_t)n * p.K + k + i] : 0.f;
      }
    }
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lrow] = a[i];
      Bs[lk + i][lrow] = b[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFcBK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.N) Y[(int64_t)m * p.ldy + n] = apply_act(acc[i][j] + bias[n], p.act);
    }
  }
}

}  // namespace drs
