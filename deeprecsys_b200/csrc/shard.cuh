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

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spin until own_flags[q] >= epoch for every q.  The spin is bounded (kShardWaitTimeoutNs): a
// peer that died or diverged must not leave this GPU spinning forever — the wait gives up, raises
// the lane's error flag (reported as DRS_ERR_COMM by the host) and lets the stream drain.
constexpr uint64_t kShardWaitTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__global__ void k_shard_wait(const uint32_t* own_flags, int world, uint32_t epoch,
                             int* comm_err_flag) {
  const int q = threadIdx.x;
  if (q < world) {
    const uint64_t t0 = global_timer_ns();
    uint32_t v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(own_flags + q) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (global_timer_ns() - t0 > kShardWaitTimeoutNs) {
        if (comm_err_flag) *comm_err_flag = 2;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  __threadfence_system();
}

// Before a query's peer stores on a lane: publish "the buffer I filled for the previous query on
// this lane has been consumed" (true by stream order: its MLPs are done) to every peer, THEN wait
// until every peer has published the same — signal before wait, so ranks never wait on each
// other's waits.  One warp: it can spin without holding an SM's worth of resources.
__global__ void k_shard_sync_consumed(const ShardFlags f, uint32_t epoch_done, int* comm_err_flag) {
  const int q = threadIdx.x;
  if (q < f.world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.peer_flags[q] + f.rank),
                 "r"(epoch_done)
                 : "memory");
    const uint32_t* own = f.peer_flags[f.rank] + q;
    const uint64_t t0 = global_timer_ns();
    uint32_t v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(own) : "memory");
      if ((int32_t)(v - epoch_done) >= 0) break;
      if (global_timer_ns() - t0 > kShardWaitTimeoutNs) {
        if (comm_err_flag) *comm_err_flag = 2;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  __threadfence_system();
}

}  // namespace drs