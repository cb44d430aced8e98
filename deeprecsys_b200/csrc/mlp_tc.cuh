// mlp_tc.cuh — K3 tensor-core path (tcgen05 + TMEM) for the wide FC layers.
// Placeholder interface: the FFMA path in mlp_ffma.cuh is the only one wired
// until the tcgen05 kernel lands ("K3").
#pragma once

#include "common.cuh"

namespace drs {

struct TcFcLaunch {
  const float* X;  int64_t ldx;  int64_t x_bstride;
  const float* W;  const float* W_lo;  int64_t w_bstride;
  const float* bias;  int64_t b_bstride;
  float* Y;  int64_t ldy;  int64_t y_bstride;
  int M, N, K;
  int act;
  bool split3;
};

inline bool tc_fc_supported(int /*N*/, int /*K*/, int64_t /*ldx*/, int64_t /*ldy*/,
                            const float* /*X*/, const float* /*Y*/) {
  return false;
}

inline int tc_fc_launch(const TcFcLaunch&, int /*n_units*/, cudaStream_t) {
  return (int)cudaErrorNotSupported;
}

inline int tc_split_weights(const float*, float*, int64_t, cudaStream_t) { return 0; }

}  // namespace drs
