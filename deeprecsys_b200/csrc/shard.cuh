// shard.cuh — flag barrier for the table-sharded multi-GPU exchange.
//
// The exchange itself is K1's epilogue (sls_gather.cuh: pooled vectors are stored
// straight into the owning rank's buffer through peer-mapped memory).  These two
// tiny kernels order it: after K1, every rank publishes "my stores for epoch n
// have been issued" into each peer's flag array (system-scope release), then
// waits until all peers have published the same epoch into its own array
// (system-scope acquire).  The same pair, on a second flag array, tells the
// writers that a buffer has been consumed and may be overwritten.
#pragma once

#include "common.cuh"

namespace drs {

struct ShardFlags {
  uint32_t* peer_flags[8];   // flag array [world] inside each rank's exchange slab (peer-mapped)
  int world, rank;
};

// flags_q[rank] = epoch for every rank q (including self)
__global__ void k_shard_signal(const ShardFlags f, uint32_t epoch) {
  const int q = threadIdx.x;
  if (q < f.world) {
    __threadfence_system();   // earlier kernels' peer stores happen-before the flag
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.peer_flags[q] + f.rank), "r"(epoch)
                 : "memory");
  }
}

// spin until own_flags[q] >= epoch for every q
__global__ void k_shard_wait(const uint32_t* own_flags, int world, uint32_t epoch) {
  const int q = threadIdx.x;
  if (q < world) {
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(own_flags + q) : "memory");
    } while ((int32_t)(v - epoch) < 0);
  }
  __syncthreads();
  __threadfence_system();
}

}  // namespace drs
