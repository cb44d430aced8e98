// rowshard.cuh — row-wise table sharding (SURVEY 8e "row-wise"): rank r holds rows
// [r*chunk, min(N, (r+1)*chunk)), chunk = ceil(N / world), of EVERY table.
//
//   1. k_row_bucket_count / k_row_bucket_fill: integer-exact bucketing of a query's ids by this
//      rank's row range.  For every (table, sample) segment the ids that fall into the range are
//      kept in their original order and rebased (id - row0); per-segment counts become the local
//      lengths (-> k_lengths_to_offsets -> local offsets).  One warp per segment, 32 ids per
//      step, ballot + popc compaction, coalesced reads and writes.
//   2. K1 (sls_gather.cuh) runs unchanged on the local ids / local table slices and stores the
//      PARTIAL pooled vectors into the owner rank's partial slab [world][rows][T*D] over NVLink.
//   3. k_row_reduce: the owner adds the world partials of each of its samples in rank order
//      (deterministic; reassociates the fp32 sum relative to the single-GPU order, which is why
//      row-wise pooled vectors are compared with rtol 1e-6 against the sequential oracle and
//      bit-exactly against an oracle that emulates this order).
#pragma once

#include "common.cuh"

namespace drs {

struct RowBucketDesc {            // one per table
  const int32_t* idx_raw;         // the query's ids of this table (global row ids)
  const int32_t* off_raw;         // [B+1] offsets of the raw segments
  int32_t* len_local;             // [B]   out (count pass)
  const int32_t* off_local;       // [B+1] in  (fill pass)
  int32_t* idx_local;             // out (fill pass): id - row0, original order
  int64_t n_rows;                 // full table rows (SparseLengthsSum enforce: 0 <= id < n_rows)
  int32_t row0, row1;             // this rank's row range [row0, row1)
};

constexpr int kBucketThreads = 256;

template <bool FILL>
__global__ void __launch_bounds__(kBucketThreads) k_row_bucket(const RowBucketDesc* tabs, int T,
                                                              int B, int* err_flag) {
  const int warp = (blockIdx.x * kBucketThreads + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= T * B) return;
  const int t = warp / B, b = warp - t * B;
  const RowBucketDesc td = tabs[t];
  const int beg = __ldg(td.off_raw + b), end = __ldg(td.off_raw + b + 1);
  int kept = 0;
  const int base = FILL ? __ldg(td.off_local + b) : 0;
  for (int i = beg; i < end; i += kWarp) {
    const bool live = i + lane < end;
    const int32_t id = live ? __ldg(td.idx_raw + i + lane) : -1;
    if (live && (uint32_t)id >= (uint32_t)td.n_rows) *err_flag = 1;
    const bool mine = live && id >= td.row0 && id < td.row1;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (FILL && mine) td.idx_local[base + kept + __popc(m & ((1u << lane) - 1u))] = id - td.row0;
    kept += __popc(m);
  }
  if (!FILL && lane == 0) td.len_local[b] = kept;
}

// out[row, col0 + j] = ((p_0 + p_1) + p_2) + ...   over the world partials, j in [0, TD)
__global__ void k_row_reduce(const float* __restrict__ partial, int world, int64_t cap_rows, int TD,
                             int rows, float* __restrict__ out, int64_t ld_out, int col0) {
  const int64_t n = (int64_t)rows * TD;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / TD;
    const int j = (int)(i - row * TD);
    float acc = partial[row * TD + j];
    for (int r = 1; r < world; ++r) acc += partial[((int64_t)r * cap_rows + row) * TD + j];
    out[row * ld_out + col0 + j] = acc;
  }
}

}  // namespace drs
