// Host-side id packer of the upload path (engine-internal wire layout; compiled by g++ from
// hostpack.cpp, not by nvcc: it uses GCC vector extensions and per-ISA function clones).
//
// The upload of a query is bound by the PCIe link (SURVEY hard part 2): an id of a 4 M-row table
// needs 22 bits, the reference hands the engine int64 (accelInferenceEngine -> run_queues,
// /root/reference/models/dlrm_s_caffe2.py:223-226).  With drs_set_option(e, "host_pack", n) the
// engine packs the caller's int32/int64 ids on n host threads into a pinned staging buffer while
// the previous query's copy is in flight, ships ceil(log2(rows)) bits per id and widens them on
// the GPU (k_unpack_ids_blk, sls_gather.cuh).
//
// Layout ("lane-interleaved blocks"): ids are cut into blocks of 512 = 32 groups of 16.  Group j
// of a block is one 16-lane vector; lane l carries ids l, 16+l, 32+l, ... of the block as one
// little-endian bit stream of `bits` words, and word w of lane l sits at block[w*16 + l].  The
// packer is then whole-vector shifts and ORs (one 64-byte store per 32/bits groups); the GPU
// reads id r of a block from words (r>>4)*bits>>5 (and the next) of lane r&15 — consecutive
// threads read consecutive words.  The last block is zero-padded.
#pragma once
#include <cstdint>

namespace drs {

constexpr int kPackBlockIds = 512;
constexpr int kPackLanes = 16;

inline int64_t pack_blocks_of(int64_t n_ids) { return (n_ids + kPackBlockIds - 1) / kPackBlockIds; }
inline int64_t pack_block_words(int64_t n_ids, int bits) {
  return pack_blocks_of(n_ids) * kPackLanes * bits;
}

struct PackJob {
  const void* ids = nullptr;   // int32 or int64, n of them
  int is_i64 = 0;
  int64_t n = 0;
  int bits = 0;                // 1..31
  uint32_t* out = nullptr;     // pack_block_words(n, bits) words
  int ok = 1;                  // result: 0 = some id is negative or does not fit `bits`
};

class HostPacker;
HostPacker* hostpack_create(int threads);     // threads >= 1 (the caller counts as one)
void hostpack_destroy(HostPacker* p);
int hostpack_threads(const HostPacker* p);
// Packs every job; jobs are worked on in order, and `done(j, ctx)` is called on the CALLING
// thread, in job order, as soon as job j is complete (so that its copy can be enqueued while
// later jobs are still being packed).
void hostpack_run(HostPacker* p, PackJob* jobs, int n_jobs, void (*done)(int, void*), void* ctx);

}  // namespace drs
