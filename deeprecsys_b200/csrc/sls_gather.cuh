// sls_gather.cuh — K1: multi-table SparseLengthsSum gather-sum for sm_100a,
// fused with the feature interaction's output layout (cat) or with the
// pairwise-dot interaction itself (dot).
//
// Replaces, for all T tables of a query in ONE launch:
//   SparseLengthsSum x T   /root/reference/models/dlrm_s_caffe2.py:319-325
//   Concat(axis=1)         dlrm_s_caffe2.py:358-360            (cat)
//   Concat(add_axis)+BatchMatMul+Flatten+BatchGather+Concat  :337-354  (dot)
//
// Numerics contract: Caffe2's CPU SparseLengthsSum accumulates
// every output column in fp32, sequentially in index order, starting from 0.
// Here one lane owns 4 adjacent columns of one (sample, table) segment and
// walks that segment's rows in index order, so the pooled vectors are
// bit-identical.  A segment is never split across lanes that would need a
// cross-lane reduction.
//
// Mapping: a CTA owns a tile of SB consecutive samples x all T tables.
// LPR = D/4 lanes cover one 4*LPR-float row with one 16-byte load each
// (D=32: 8 lanes x 16 B = one 128-B line; a warp therefore has 4 segments in
// flight).  Each lane group keeps UNROLL independent row loads outstanding and
// applies the adds in order.  The tile's offsets and index spans are staged in
// shared memory; index spans arrive by 1-D TMA bulk copies (cp.async.bulk ->
// UBLKCP) completing on an mbarrier, one copy per table.
#pragma once

#include "common.cuh"
#include "shard.cuh"

namespace drs {

#ifndef DRS_SLS_THREADS
#define DRS_SLS_THREADS 256
#endif
#ifndef DRS_SLS_UNROLL
#define DRS_SLS_UNROLL 8
#endif
// DRS_SLS_MIN_CTAS (tuning builds only): register cap of the persistent gather as "min resident
// CTAs per SM" (5 -> 48 registers).  Undefined in the product: a second __launch_bounds__
// argument — even 1 — changes ptxas' allocation (96 registers instead of 64, i.e. 2 CTAs per SM
// instead of 4; measured as 360 M instead of 515 M inf/s on 8 GPUs, profiles/r02v_*).
#ifdef DRS_SLS_MIN_CTAS
#define DRS_SLS_PERSIST_BOUNDS __launch_bounds__(kSlsThreads, DRS_SLS_MIN_CTAS)
#else
#define DRS_SLS_PERSIST_BOUNDS __launch_bounds__(kSlsThreads)
#endif
constexpr int kSlsThreads = DRS_SLS_THREADS;
constexpr int kSlsUnroll = DRS_SLS_UNROLL;
constexpr int kMaxShardRanks = 8;

struct SlsTableDesc {
  const float* rows;      // [n_rows, D] table in HBM
  const int32_t* idx;     // ids of this table for the staged batch / query
  const int32_t* off;     // [B+1] exclusive prefix of lengths
  int64_t n_rows;
  int32_t dst_col;        // first output column of this table's pooled vector
  int32_t pad_;
};

// Coalesced serving (drs_submit_many): several queued requests, each "the first size_k samples of
// staged batch b_k" (inferenceEngine.py:200-215), are pooled by ONE launch into consecutive row
// ranges of the interaction buffer.  A tile never straddles two requests.
constexpr int kMaxPackets = 32;
struct SlsPackets {
  int n;                               // 0: a single query (SlsLaunch::tables / B)
  int tile_off[kMaxPackets + 1];       // first tile of request k; tile_off[n] = number of tiles
  int row_off[kMaxPackets];            // first output row of request k
  int size[kMaxPackets];               // samples of request k
  const SlsTableDesc* descs[kMaxPackets];   // gather descriptors of request k's staged batch
};

struct SlsLaunch {
  const SlsTableDesc* tables;
  int T, B, D;
  float* out;             // [B, ld_out]
  int64_t ld_out;
  int SB;                 // samples per CTA tile
  int idx_cap;            // ints of smem per table index span (0 = read ids from global)
  int vec_store;          // out/ld_out/dst_col all 16-B aligned
  int* err_flag;          // SparseLengthsSum enforce violations
  // dot interaction fusion (FUSE_DOT kernels only)
  const float* dense;     // [B, ld_dense] bottom-MLP output (feature 0)
  int64_t ld_dense;
  int itself;             // arch_interaction_itself
  float* pooled_tap;      // optional [B, T*D] copy of pooled vectors (debug tap), or nullptr
  // table-sharded multi-GPU (shard_S > 0): sample b belongs to rank b / shard_S and its pooled
  // vector is stored straight into that rank's buffer (peer-mapped over NVLink) at local row
  // b % shard_S — the all-to-all exchange is this kernel's epilogue.
  int shard_S;
  float* peer_out[kMaxShardRanks];
  // CTA order.  0: sample-major (a CTA owns SB samples x all T tables; required by FUSE_DOT).
  // > 0: table-major — CTA b handles table b / tiles_per_table, samples tile b % tiles_per_table
  // (SB samples of that one table).  CTAs are scheduled in index order, so the whole chip sweeps
  // one table at a time; ids are sorted inside every sample (np.unique in the generator), so all
  // resident segments walk the table in ascending order roughly in lockstep and concurrent row
  // reads fall into a narrow band of the table (DRAM page locality, L2 hits on shared rows).
  int tiles_per_table;
  // short segments (avg length < 8: WnD / NCF / DIEN / DIN): the persistent kernel's SHORT
  // instantiation walks kSlsShort segments per lane group at once
  int short_seg;
  SlsPackets pk;          // MULTI instantiation of the persistent kernel only
  // Sharded exchange (shard_S > 0 and sync_counter != nullptr): the last CTA to finish publishes
  // "my stores of `epoch` are issued" to every peer — the arrival signal rides in the kernel's
  // tail instead of a launch of its own.  Nothing in this kernel WAITS for a peer: a persistent
  // kernel that spins on remote progress while it holds the SMs could starve the very kernel
  // (another lane's) whose signal a peer is waiting for.  The waits are one-warp kernels
  // (shard.cuh).
  uint32_t* sync_counter;               // device-local CTA counter, returns to 0 after every launch
  uint32_t epoch;
  int rank, world;
  uint32_t* arrived_flags[kMaxShardRanks];    // peer q's arrived[] array (peer-mapped)
};

// tail of a sharded gather: the last CTA publishes the arrival flags.
__device__ __forceinline__ void sls_shard_tail(const SlsLaunch& p) {
  if (p.shard_S <= 0 || p.sync_counter == nullptr) return;
  __threadfence_system();                       // this thread's peer stores before the flag
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(p.sync_counter, 1u);
    if (prev == gridDim.x - 1) {
      *p.sync_counter = 0;                      // next launch on this lane starts from zero
      __threadfence_system();
      for (int q = 0; q < p.world; ++q)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.arrived_flags[q] + p.rank),
                     "r"(p.epoch)
                     : "memory");
    }
  }
}

__host__ __device__ inline size_t sls_smem_bytes(int T, int SB, int idx_cap, int D, bool fuse_dot) {
  size_t b = 16;                                        // mbarrier (+pad)
  b += (size_t)T * idx_cap * sizeof(int32_t);           // index spans
  b += (size_t)T * (SB + 1) * sizeof(int32_t);          // offsets
  b = (b + 15) & ~size_t(15);
  if (fuse_dot) b += (size_t)SB * (T + 1) * (D + 1) * sizeof(float);
  return b;
}

// One segment: rows[ids[beg..end)] summed in order. Lane owns columns 4*lane..4*lane+3.
template <bool SMEM_IDX>
__device__ __forceinline__ float4 sls_segment_vec4(const float* __restrict__ rows_lane,
                                                   uint32_t n_rows, int D,
                                                   const int32_t* __restrict__ ids, int beg,
                                                   int end, int* err_flag) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int i = beg;
  for (; i + kSlsUnroll <= end; i += kSlsUnroll) {
    int32_t id[kSlsUnroll];
    float4 v[kSlsUnroll];
#pragma unroll
    for (int u = 0; u < kSlsUnroll; ++u) id[u] = SMEM_IDX ? ids[i + u] : __ldg(ids + i + u);
#pragma unroll
    for (int u = 0; u < kSlsUnroll; ++u) {
      if ((uint32_t)id[u] < n_rows) {
        v[u] = ldg_stream_f4(rows_lane + (int64_t)id[u] * D);
      } else {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        *err_flag = 1;
      }
    }
#pragma unroll
    for (int u = 0; u < kSlsUnroll; ++u) {
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
  }
  if (i < end) {  // tail: < UNROLL rows, still all loads first, adds in order
    int32_t id[kSlsUnroll];
    float4 v[kSlsUnroll];
#pragma unroll
    for (int u = 0; u < kSlsUnroll; ++u)
      id[u] = (i + u < end) ? (SMEM_IDX ? ids[i + u] : __ldg(ids + i + u)) : -1;
#pragma unroll
    for (int u = 0; u < kSlsUnroll; ++u) {
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i + u < end) {
        if ((uint32_t)id[u] < n_rows) v[u] = ldg_stream_f4(rows_lane + (int64_t)id[u] * D);
        else *err_flag = 1;
      }
    }
#pragma unroll
    for (int u = 0; u < kSlsUnroll; ++u)
      if (i + u < end) {
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
  }
  return acc;
}


// Short segments (WnD / NCF / DIEN / DIN have 1-3 lookups per table): one lane group walks
// kSlsShort segments AT ONCE — row j of each
// of them is loaded before any add — so that kSlsShort independent row reads are in flight per
// group instead of one.  Inside every segment the adds still happen in index order.
constexpr int kSlsShort = 4;

struct ShortSeg {
  const float* rows_lane;   // table base + this lane's column offset
  const int32_t* ids;       // table's id array: TMA-staged shared copy (rebased) or global
  uint32_t n_rows;
  int beg, end;
};

__device__ __forceinline__ void sls_short_segments(const ShortSeg (&sg)[kSlsShort], int D,
                                                   float4 (&acc)[kSlsShort], int* err_flag) {
  int maxlen = 0;
#pragma unroll
  for (int u = 0; u < kSlsShort; ++u) {
    acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    maxlen = max(maxlen, sg[u].end - sg[u].beg);
  }
  for (int j = 0; j < maxlen; ++j) {
    float4 v[kSlsShort];
    bool live[kSlsShort];
#pragma unroll
    for (int u = 0; u < kSlsShort; ++u) {
      live[u] = sg[u].beg + j < sg[u].end;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live[u]) {
        const int32_t id = sg[u].ids[sg[u].beg + j];     // shared (TMA-staged) or global
        if ((uint32_t)id < sg[u].n_rows) v[u] = ldg_stream_f4(sg[u].rows_lane + (int64_t)id * D);
        else *err_flag = 1;
      }
    }
#pragma unroll
    for (int u = 0; u < kSlsShort; ++u)
      if (live[u]) {
        acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
      }
  }
}

// LPR: lanes per row (D == 4*LPR).  LPR == 0 selects the generic any-D path
// (one warp per segment, lane <-> column, scalar loads).
template <int LPR, bool FUSE_DOT>
__global__ void __launch_bounds__(kSlsThreads) k_sls_gather(const SlsLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Tc = p.tiles_per_table > 0 ? 1 : p.T;   // tables per CTA tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  int32_t* sidx = reinterpret_cast<int32_t*>(smem_raw + 16);
  int32_t* soff = sidx + (size_t)Tc * p.idx_cap;
  float* spool = nullptr;  // [SB][T+1][D+1] (dot only)
  if (FUSE_DOT) {
    size_t o = 16 + ((size_t)Tc * p.idx_cap + (size_t)Tc * (p.SB + 1)) * sizeof(int32_t);
    o = (o + 15) & ~size_t(15);
    spool = reinterpret_cast<float*>(smem_raw + o);
  }

  const int tid = threadIdx.x;
  const bool tm = p.tiles_per_table > 0;
  const int t_first = tm ? blockIdx.x / p.tiles_per_table : 0;
  const int T = tm ? 1 : p.T;
  const SlsTableDesc* __restrict__ tabs = p.tables + t_first;
  const int b0 = (tm ? blockIdx.x - t_first * p.tiles_per_table : blockIdx.x) * p.SB;
  const int sbe = min(p.SB, p.B - b0);
  const int SB1 = p.SB + 1;
  const bool use_smem_idx = p.idx_cap > 0;

  // 1. offsets of the tile -> smem
  for (int i = tid; i < T * SB1; i += kSlsThreads) {
    const int t = i / SB1, j = i - t * SB1;
    soff[i] = (j <= sbe) ? __ldg(tabs[t].off + b0 + j) : 0;
  }
  if (use_smem_idx && tid == 0) {
    mbar_init(bar, (uint32_t)T);
    mbar_fence_init();
  }
  __syncthreads();

  // 2. index spans of the tile -> smem by TMA bulk copies (one per table)
  if (use_smem_idx) {
    for (int t = tid; t < T; t += kSlsThreads) {
      const int beg = soff[t * SB1], end = soff[t * SB1 + sbe];
      const int a0 = beg & ~3;
      const int n = (end - a0 + 3) & ~3;
      if (n > 0 && n <= p.idx_cap) {
        mbar_arrive_expect_tx(bar, (uint32_t)n * 4u);
        tma_load_1d(sidx + (size_t)t * p.idx_cap, tabs[t].idx + a0, (uint32_t)n * 4u, bar);
      } else {
        mbar_arrive(bar);
      }
    }
  }
  if (FUSE_DOT) {  // feature 0 = bottom-MLP output
    const int Dp = p.D + 1, F = p.T + 1;
    for (int i = tid; i < sbe * p.D; i += kSlsThreads) {
      const int lb = i / p.D, c = i - lb * p.D;
      spool[(size_t)(lb * F) * Dp + c] = __ldg(p.dense + (int64_t)(b0 + lb) * p.ld_dense + c);
    }
  }
  if (use_smem_idx) mbar_wait(bar, 0);

  // 3. gather: lane groups walk segments
  const int nseg = sbe * T;
  if (LPR > 0) {
    constexpr int L = LPR > 0 ? LPR : 1;
    const int group = tid / L, lane = tid - group * L;
    constexpr int ngroups = kSlsThreads / L;
    for (int s = group; s < nseg; s += ngroups) {
      const int t = s / sbe, lb = s - t * sbe;
      const SlsTableDesc td = tabs[t];
      const int tile_beg = soff[t * SB1], tile_end = soff[t * SB1 + sbe];
      const int beg = soff[t * SB1 + lb], end = soff[t * SB1 + lb + 1];
      const int a0 = tile_beg & ~3;
      const bool in_smem = use_smem_idx && (((tile_end - a0 + 3) & ~3) <= p.idx_cap);
      const float* rows_lane = td.rows + lane * 4;
      float4 acc;
      if (in_smem)
        acc = sls_segment_vec4<true>(rows_lane, (uint32_t)td.n_rows, p.D,
                                     sidx + (size_t)t * p.idx_cap - a0, beg, end, p.err_flag);
      else
        acc = sls_segment_vec4<false>(rows_lane, (uint32_t)td.n_rows, p.D, td.idx, beg, end,
                                      p.err_flag);
      const int64_t b = b0 + lb;
      if (FUSE_DOT) {
        float* d = spool + (size_t)(lb * (p.T + 1) + 1 + t) * (p.D + 1) + lane * 4;
        d[0] = acc.x; d[1] = acc.y; d[2] = acc.z; d[3] = acc.w;
        if (p.pooled_tap)
          *reinterpret_cast<float4*>(p.pooled_tap + b * ((int64_t)p.T * p.D) + (int64_t)t * p.D +
                                     lane * 4) = acc;
      } else {
        float* o = (p.shard_S > 0
                        ? p.peer_out[b / p.shard_S] + (b % p.shard_S) * p.ld_out
                        : p.out + b * p.ld_out) + td.dst_col + lane * 4;
        if (p.vec_store) {
          *reinterpret_cast<float4*>(o) = acc;
        } else {
          o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
        }
      }
    }
  } else {
    // generic D: one warp per segment, lane <-> column (cols lane, lane+32, ...)
    const int warp = tid / kWarp, lane = tid % kWarp;
    constexpr int nwarps = kSlsThreads / kWarp;
    for (int s = warp; s < nseg; s += nwarps) {
      const int t = s / sbe, lb = s - t * sbe;
      const SlsTableDesc td = tabs[t];
      const int beg = soff[t * SB1 + lb], end = soff[t * SB1 + lb + 1];
      const int64_t b = b0 + lb;
      for (int c = lane; c < p.D; c += kWarp) {
        float acc = 0.f;
        for (int i = beg; i < end; ++i) {
          const int32_t id = __ldg(td.idx + i);
          if ((uint32_t)id < (uint32_t)td.n_rows) acc += ldg_stream_f1(td.rows + (int64_t)id * p.D + c);
          else *p.err_flag = 1;
        }
        if (FUSE_DOT) {
          spool[(size_t)(lb * (p.T + 1) + 1 + t) * (p.D + 1) + c] = acc;
          if (p.pooled_tap) p.pooled_tap[b * ((int64_t)p.T * p.D) + (int64_t)t * p.D + c] = acc;
        } else if (p.shard_S > 0) {
          p.peer_out[b / p.shard_S][(b % p.shard_S) * p.ld_out + td.dst_col + c] = acc;
        } else {
          p.out[b * p.ld_out + td.dst_col + c] = acc;
        }
      }
    }
  }

  if (!FUSE_DOT) sls_shard_tail(p);

  if (FUSE_DOT) {
    // 4. pairwise dot: R[b] = [x, (T_i . T_j) for i in 0..F-1, j < i + itself]
    __syncthreads();
    const int F = p.T + 1, Dp = p.D + 1;
    const int npair = p.itself ? F * (F + 1) / 2 : F * (F - 1) / 2;
    for (int w = tid; w < sbe * npair; w += kSlsThreads) {
      const int lb = w / npair, q = w - lb * npair;
      // invert the row-major lower-triangle enumeration
      int i, j;
      if (p.itself) {
        i = (int)((sqrtf(8.f * q + 1.f) - 1.f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= q) ++i;
        while (i * (i + 1) / 2 > q) --i;
        j = q - i * (i + 1) / 2;
      } else {
        i = (int)((sqrtf(8.f * q + 1.f) + 1.f) * 0.5f);
        while (i * (i + 1) / 2 <= q) ++i;
        while (i * (i - 1) / 2 > q) --i;
        j = q - i * (i - 1) / 2;
      }
      const float* a = spool + (size_t)(lb * F + i) * Dp;
      const float* c = spool + (size_t)(lb * F + j) * Dp;
      float z = 0.f;
      for (int k = 0; k < p.D; ++k) z = fmaf(a[k], c[k], z);
      p.out[(int64_t)(b0 + lb) * p.ld_out + p.D + q] = z;
    }
    for (int w = tid; w < sbe * p.D; w += kSlsThreads) {
      const int lb = w / p.D, c = w - lb * p.D;
      p.out[(int64_t)(b0 + lb) * p.ld_out + c] = spool[(size_t)(lb * F) * Dp + c];
    }
  }
}

// Persistent variant of K1 (cat layout / plain pooling, vectorised rows only): a fixed number of
// CTAs per SM loop over the sample tiles (tile = blockIdx.x + k * gridDim.x).  While tile k is
// being gathered, the offsets of tile k+2 are in registers and the index spans of tile k+1 are
// landing in the other half of the double-buffered shared memory (TMA bulk copies on a second
// mbarrier).  Because the CTAs never retire, their SM footprint is static, which leaves a fixed
// share of registers / shared memory on every SM for the tensor-core MLP kernel of the
// neighbouring lane to run *concurrently* with the gather (engine.cu lanes).
constexpr int kPersistOffRegs = 2;     // offsets per thread kept in registers: T*(SB+1) <= 512

__host__ __device__ inline size_t sls_persist_smem_bytes(int T, int SB, int idx_cap,
                                                         bool short_seg = false) {
  size_t b = 32;                                              // two mbarriers (+pad)
  b += 2 * (size_t)T * idx_cap * sizeof(int32_t);             // index spans, double-buffered
  b += 2 * (size_t)T * (SB + 1) * sizeof(int32_t);            // offsets, double-buffered
  b = (b + 15) & ~size_t(15);
  if (short_seg) b += (size_t)T * sizeof(SlsTableDesc);       // table descriptors (SHORT path)
  return (b + 15) & ~size_t(15);
}

// SHORT: compile the multi-segment path for short segments (ids from global memory) instead of
// the long-segment path — a separate instantiation so that its extra registers do not lower the
// occupancy of the HBM-bound long-segment kernel.
// MULTI: the tiles enumerate the requests of SlsLaunch::pk instead of one query.
template <int LPR, bool SHORT, bool MULTI = false>
__global__ void DRS_SLS_PERSIST_BOUNDS k_sls_gather_persist(const SlsLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);            // bar[0], bar[1]
  int32_t* sidx0 = reinterpret_cast<int32_t*>(smem_raw + 32);
  const size_t idx_words = (size_t)p.T * p.idx_cap;
  const int SB1 = p.SB + 1;
  const int noff = p.T * SB1;
  int32_t* soff0 = sidx0 + 2 * idx_words;

  const int tid = threadIdx.x;
  const int T = p.T;
  const int ntiles = MULTI ? p.pk.tile_off[p.pk.n] : (p.B + p.SB - 1) / p.SB;
  const bool use_smem_idx = p.idx_cap > 0;
  constexpr int L = LPR;
  const int group = tid / L, lane = tid - group * L;
  constexpr int ngroups = kSlsThreads / L;

  // where a tile lives: descriptors of its request, first sample inside the request's staged
  // batch, samples in the tile, first output row
  struct TileLoc { const SlsTableDesc* tabs; int b0, sbe, row0; };
  auto locate = [&](int tile) -> TileLoc {
    if (!MULTI) {
      const int b0 = tile * p.SB;
      return TileLoc{p.tables, b0, min(p.SB, p.B - b0), b0};
    }
    int k = 0;
    while (k + 1 < p.pk.n && tile >= p.pk.tile_off[k + 1]) ++k;
    const int b0 = (tile - p.pk.tile_off[k]) * p.SB;
    return TileLoc{p.pk.descs[k], b0, min(p.SB, p.pk.size[k] - b0), p.pk.row_off[k] + b0};
  };

  // offsets of tile `tile` held by this thread (entries tid, tid + kSlsThreads)
  auto load_offs = [&](int tile, int32_t* regs) {
    TileLoc loc{p.tables, 0, 0, 0};
    if (tile < ntiles) loc = locate(tile);
    const int b0 = loc.b0, sbe = loc.sbe;
#pragma unroll
    for (int k = 0; k < kPersistOffRegs; ++k) {
      const int i = tid + k * kSlsThreads;
      regs[k] = 0;
      if (i < noff && tile < ntiles) {
        const int t = i / SB1, j = i - t * SB1;
        if (j <= sbe) regs[k] = __ldg(loc.tabs[t].off + b0 + j);
      }
    }
  };
  auto store_offs = [&](int buf, const int32_t* regs) {
#pragma unroll
    for (int k = 0; k < kPersistOffRegs; ++k) {
      const int i = tid + k * kSlsThreads;
      if (i < noff) soff0[(size_t)buf * noff + i] = regs[k];
    }
  };
  auto issue_spans = [&](int tile, int buf) {      // index spans of `tile` -> sidx[buf] by TMA
    const TileLoc loc = locate(tile);
    const int sbe = loc.sbe;
    const int32_t* so = soff0 + (size_t)buf * noff;
    for (int t = tid; t < T; t += kSlsThreads) {
      const int beg = so[t * SB1], end = so[t * SB1 + sbe];
      const int a0 = beg & ~3;
      const int n = (end - a0 + 3) & ~3;
      if (n > 0 && n <= p.idx_cap) {
        mbar_arrive_expect_tx(&bar[buf], (uint32_t)n * 4u);
        tma_load_1d(sidx0 + (size_t)buf * idx_words + (size_t)t * p.idx_cap, loc.tabs[t].idx + a0,
                    (uint32_t)n * 4u, &bar[buf]);
      } else {
        mbar_arrive(&bar[buf]);
      }
    }
  };

  int tile = blockIdx.x;
  if (tile >= ntiles) {                 // (the grid never exceeds the tiles; kept for safety)
    sls_shard_tail(p);
    return;
  }
  // SHORT: a segment is a handful of rows, so even the 40-byte descriptor fetch from global
  // memory would sit on its critical path -> keep the descriptors in shared memory (single
  // query only: coalesced requests bring one descriptor set each)
  const SlsTableDesc* __restrict__ tabs = p.tables;
  const SlsTableDesc* sdesc = tabs;
  if (SHORT && !MULTI) {
    size_t o = 32 + (2 * idx_words + 2 * (size_t)noff) * sizeof(int32_t);
    o = (o + 15) & ~size_t(15);
    SlsTableDesc* sd = reinterpret_cast<SlsTableDesc*>(smem_raw + o);
    for (int i = tid; i < T * (int)(sizeof(SlsTableDesc) / 8); i += kSlsThreads)
      reinterpret_cast<uint64_t*>(sd)[i] = reinterpret_cast<const uint64_t*>(tabs)[i];
    sdesc = sd;                        // visible after the first __syncthreads below
  }
  int32_t regs[kPersistOffRegs];
  load_offs(tile, regs);
  if (use_smem_idx && tid == 0) {
    mbar_init(&bar[0], (uint32_t)T);
    mbar_init(&bar[1], (uint32_t)T);
    mbar_fence_init();
  }
  store_offs(0, regs);
  load_offs(tile + gridDim.x, regs);               // offsets of the next tile stay in registers
  __syncthreads();
  if (use_smem_idx) issue_spans(tile, 0);

  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int cur = it & 1, nxt = cur ^ 1;
    const int next_tile = tile + gridDim.x;
    if (it > 0) __syncthreads();                   // everyone is done with buffers [nxt] (tile it-1)
    if (next_tile < ntiles) {
      store_offs(nxt, regs);
      load_offs(next_tile + gridDim.x, regs);
    }
    __syncthreads();
    if (use_smem_idx && next_tile < ntiles) issue_spans(next_tile, nxt);
    if (use_smem_idx) mbar_wait(&bar[cur], (uint32_t)((it >> 1) & 1));

    const TileLoc loc = locate(tile);
    const int b0 = loc.row0;                       // first OUTPUT row of the tile
    const int sbe = loc.sbe;
    if (MULTI) { tabs = loc.tabs; sdesc = loc.tabs; }
    const int32_t* so = soff0 + (size_t)cur * noff;
    const int32_t* si = sidx0 + (size_t)cur * idx_words;
    const int nseg = sbe * T;
    if (SHORT) {                         // short segments: kSlsShort of them per group at once
      for (int s0 = group; s0 < nseg; s0 += ngroups * kSlsShort) {
        ShortSeg sg[kSlsShort];
        float* dst[kSlsShort];
#pragma unroll
        for (int u = 0; u < kSlsShort; ++u) {
          const int s = s0 + u * ngroups;
          dst[u] = nullptr;
          sg[u].beg = sg[u].end = 0;
          sg[u].rows_lane = nullptr; sg[u].ids = nullptr; sg[u].n_rows = 0;
          if (s < nseg) {
            const int t = s / sbe, lb = s - t * sbe;
            const SlsTableDesc& td = sdesc[t];
            const int tile_beg = so[t * SB1], tile_end = so[t * SB1 + sbe];
            const int a0 = tile_beg & ~3;
            const bool in_smem = use_smem_idx && (((tile_end - a0 + 3) & ~3) <= p.idx_cap);
            sg[u].rows_lane = td.rows + lane * 4;
            sg[u].ids = in_smem ? si + (size_t)t * p.idx_cap - a0 : td.idx;
            sg[u].n_rows = (uint32_t)td.n_rows;
            sg[u].beg = so[t * SB1 + lb];
            sg[u].end = so[t * SB1 + lb + 1];
            const int64_t b = b0 + lb;
            dst[u] = (p.shard_S > 0 ? p.peer_out[b / p.shard_S] + (b % p.shard_S) * p.ld_out
                                    : p.out + b * p.ld_out) + td.dst_col + lane * 4;
          }
        }
        float4 acc[kSlsShort];
        sls_short_segments(sg, p.D, acc, p.err_flag);
#pragma unroll
        for (int u = 0; u < kSlsShort; ++u) {
          if (!dst[u]) continue;
          if (p.vec_store) {
            *reinterpret_cast<float4*>(dst[u]) = acc[u];
          } else {
            dst[u][0] = acc[u].x; dst[u][1] = acc[u].y; dst[u][2] = acc[u].z; dst[u][3] = acc[u].w;
          }
        }
      }
      continue;
    }
    for (int s = group; s < nseg; s += ngroups) {
      const int t = s / sbe, lb = s - t * sbe;
      const SlsTableDesc td = tabs[t];
      const int tile_beg = so[t * SB1], tile_end = so[t * SB1 + sbe];
      const int beg = so[t * SB1 + lb], end = so[t * SB1 + lb + 1];
      const int a0 = tile_beg & ~3;
      const bool in_smem = use_smem_idx && (((tile_end - a0 + 3) & ~3) <= p.idx_cap);
      const float* rows_lane = td.rows + lane * 4;
      float4 acc;
      if (in_smem)
        acc = sls_segment_vec4<true>(rows_lane, (uint32_t)td.n_rows, p.D,
                                     si + (size_t)t * p.idx_cap - a0, beg, end, p.err_flag);
      else
        acc = sls_segment_vec4<false>(rows_lane, (uint32_t)td.n_rows, p.D, td.idx, beg, end,
                                      p.err_flag);
      const int64_t b = b0 + lb;
      float* o = (p.shard_S > 0 ? p.peer_out[b / p.shard_S] + (b % p.shard_S) * p.ld_out
                                : p.out + b * p.ld_out) + td.dst_col + lane * 4;
      if (p.vec_store) {
        *reinterpret_cast<float4*>(o) = acc;
      } else {
        o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
      }
    }
  }
  sls_shard_tail(p);
}

// ---------------------------------------------------------------------------------------------
// Fixed-length short segments (Wide&Deep, NCF, DIN, DIEN, MT-WnD: every (sample, table) segment
// has the same 1..4 lookups — the reference's request slicing REQUIRES fixed L,
// inferenceEngine.py:204-206).  With L known the offsets are b*L: no offset arrays to stage, and
// the ids of 32 consecutive samples of one table are one contiguous span.  One warp takes work
// items (table t, 32 samples), t fastest, so that at any instant the chip reads many tables at
// once while every warp's id loads and pooled-vector stores are coalesced; the tile machinery of
// the persistent kernel (per-tile offsets from T arrays, T tiny TMA copies — all of a tile's cost
// when a tile is one 254-table sample) is not needed.  In-segment adds stay in index order.
constexpr int kFixedMaxL = 4;

// L: lookups per segment (compile time: the in-flight rows live in registers);
// a lane group keeps kFixedSeg segments = kFixedSeg * L row loads in flight
template <int LPR, int L>
__global__ void __launch_bounds__(kSlsThreads) k_sls_fixed(const SlsLaunch p) {
  constexpr int kFixedSeg = (L == 1 && LPR <= 8) ? 8 : 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlsTableDesc* sdesc = reinterpret_cast<SlsTableDesc*>(smem_raw);
  for (int i = threadIdx.x; i < p.T * (int)(sizeof(SlsTableDesc) / 8); i += kSlsThreads)
    reinterpret_cast<uint64_t*>(sdesc)[i] = reinterpret_cast<const uint64_t*>(p.tables)[i];
  __syncthreads();
  constexpr int G = kWarp / LPR;                        // lane groups (segments) per warp pass
  constexpr int SPW = 32;                               // samples per work item
  const int lane = threadIdx.x % kWarp;
  const int gi = lane / LPR, l = lane % LPR;
  const int warp = (blockIdx.x * kSlsThreads + threadIdx.x) / kWarp;
  const int nwarps = (gridDim.x * kSlsThreads) / kWarp;
  const int nblk = (p.B + SPW - 1) / SPW;
  const int64_t items = (int64_t)nblk * p.T;
  for (int64_t item = warp; item < items; item += nwarps) {
    const int t = (int)(item % p.T);
    const int b0 = (int)(item / p.T) * SPW;
    const SlsTableDesc& td = sdesc[t];
    const int nsamp = min(SPW, p.B - b0);
    // the item's ids: nsamp * L consecutive ints, loaded coalesced (<= kFixedMaxL per lane)
    int32_t myid[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int pos = j * kWarp + lane;
      myid[j] = pos < nsamp * L ? __ldg(td.idx + (int64_t)b0 * L + pos) : 0;
    }
    const float* rows_lane = td.rows + l * 4;
    for (int s0 = gi * (SPW / G); s0 < (gi + 1) * (SPW / G); s0 += kFixedSeg) {
      float4 v[kFixedSeg][L];
#pragma unroll
      for (int u = 0; u < kFixedSeg; ++u) {
#pragma unroll
        for (int r = 0; r < L; ++r) {                    // row r of segment s0 + u
          v[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int pos = (s0 + u) * L + r;              // position in the item's id span
          int32_t id = 0;
#pragma unroll
          for (int j = 0; j < L; ++j) {                  // the shuffle is executed by every lane
            const int32_t c = __shfl_sync(0xffffffffu, myid[j], pos & (kWarp - 1));
            if ((pos >> 5) == j) id = c;
          }
          if (s0 + u < nsamp) {
            if ((uint32_t)id < (uint32_t)td.n_rows) v[u][r] = ldg_stream_f4(rows_lane + (int64_t)id * p.D);
            else *p.err_flag = 1;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kFixedSeg; ++u) {
        if (s0 + u >= nsamp) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < L; ++r) {
          acc.x += v[u][r].x; acc.y += v[u][r].y; acc.z += v[u][r].z; acc.w += v[u][r].w;
        }
        const int64_t b = b0 + s0 + u;
        float* o = (p.shard_S > 0 ? p.peer_out[b / p.shard_S] + (b % p.shard_S) * p.ld_out
                                  : p.out + b * p.ld_out) + td.dst_col + l * 4;
        if (p.vec_store) {
          *reinterpret_cast<float4*>(o) = acc;
        } else {
          o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
        }
      }
    }
  }
  sls_shard_tail(p);
}

// lengths -> offsets (exclusive prefix sum), one CTA per table; also checks
// Caffe2's enforce  sum(lengths) == len(indices)  when n_idx >= 0.
// Integer-exact by construction (int32 adds in any grouping are associative).
struct ScanTableDesc {
  const int32_t* lengths;   // [B]
  int32_t* off;             // [B+1]
  int64_t n_idx;            // expected total, or -1
};

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 16;                       // lengths per thread and pass
constexpr int kScanChunk = kScanThreads * kScanItems;

// One CTA per table, 16,384 lengths per pass: every thread scans 16 consecutive lengths in
// registers, the 1,024 thread totals are scanned with two levels of warp shuffles, and each
// thread adds its prefix while storing.  B = 16,384 is ONE pass (a few microseconds) instead of
// 64 dependent 256-wide block scans.
__global__ void __launch_bounds__(kScanThreads) k_lengths_to_offsets(const ScanTableDesc* tabs,
                                                                     int B, int* err_flag) {
  __shared__ int32_t warp_sums[kScanThreads / kWarp];
  __shared__ int32_t carry_s;
  const ScanTableDesc td = tabs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid % kWarp, warp = tid / kWarp;
  if (tid == 0) { carry_s = 0; td.off[0] = 0; }
  __syncthreads();
  for (int base = 0; base < B; base += kScanChunk) {
    const int i0 = base + tid * kScanItems;
    int32_t v[kScanItems];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      int32_t x = (i0 + k < B) ? __ldg(td.lengths + i0 + k) : 0;
      if (x < 0) { bad = true; x = 0; }
      v[k] = x;
    }
    if (bad) *err_flag = 1;
#pragma unroll
    for (int k = 1; k < kScanItems; ++k) v[k] += v[k - 1];      // inclusive, thread-local
    const int32_t total = v[kScanItems - 1];
    int32_t x = total;                                          // inclusive scan of the thread totals
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == kWarp - 1) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {                                            // 32 warp totals -> exclusive prefixes
      const int32_t w = warp_sums[lane];
      int32_t z = w;
#pragma unroll
      for (int d = 1; d < kWarp; d <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, z, d);
        if (lane >= d) z += y;
      }
      warp_sums[lane] = z - w;
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t prefix = carry + warp_sums[warp] + (x - total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (i0 + k < B) td.off[i0 + k + 1] = prefix + v[k];
    __syncthreads();
    if (tid == kScanThreads - 1) carry_s = prefix + total;
    __syncthreads();
  }
  if (tid == 0 && td.n_idx >= 0 && (int64_t)carry_s != td.n_idx) *err_flag = 1;
}

// int64 ids -> int32 (the reference's Cast op, dlrm_s_caffe2.py:308)
__global__ void k_cast_i64_i32(const int64_t* __restrict__ src, int32_t* __restrict__ dst,
                               int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (int32_t)src[i];
}

// Packed id wire format (drs_pack_ids): id i occupies bits [i*bits, (i+1)*bits) of a
// little-endian uint32 stream; one pad word follows so that a 64-bit window is always readable.
__host__ __device__ inline int64_t packed_words(int64_t n_ids, int bits) {
  return (n_ids * bits + 31) / 32 + 1;
}

__global__ void k_unpack_ids(const uint32_t* __restrict__ packed, int bits, int64_t n,
                             int32_t* __restrict__ dst) {
  const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bit = i * bits;
    const int64_t w = bit >> 5;
    const uint64_t win = (uint64_t)__ldg(packed + w) | ((uint64_t)__ldg(packed + w + 1) << 32);
    dst[i] = (int32_t)((uint32_t)(win >> (bit & 31)) & mask);
  }
}

// T equally spaced packed streams (stride_words apart) -> T equally spaced int32 arrays
__global__ void k_unpack_ids_multi(const uint32_t* __restrict__ packed, int64_t stride_words,
                                   int bits, int64_t n, int T, int32_t* __restrict__ dst,
                                   int64_t dst_stride) {
  const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
  const int64_t total = n * T;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < total;
       j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = j / n, i = j - t * n;
    const uint32_t* p = packed + t * stride_words;
    const int64_t bit = i * bits;
    const int64_t w = bit >> 5;
    const uint64_t win = (uint64_t)__ldg(p + w) | ((uint64_t)__ldg(p + w + 1) << 32);
    dst[t * dst_stride + i] = (int32_t)((uint32_t)(win >> (bit & 31)) & mask);
  }
}

// Lane-interleaved packed blocks written by the host packer (hostpack.h): 512 ids per block,
// word w of lane l at block[w*16 + l], id r of a block = bits [(r>>4)*bits, +bits) of lane r&15.
__global__ void k_unpack_ids_blk(const uint32_t* __restrict__ packed, int bits, int64_t n,
                                 int32_t* __restrict__ dst) {
  const uint32_t mask = (1u << bits) - 1u;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i & 511), l = r & 15, bit = (r >> 4) * bits;
    const int w = bit >> 5, sh = bit & 31;
    const uint32_t* blk = packed + (i >> 9) * (int64_t)(16 * bits);
    uint32_t v = __ldg(blk + w * 16 + l) >> sh;
    if (sh + bits > 32) v |= __ldg(blk + (w + 1) * 16 + l) << (32 - sh);
    dst[i] = (int32_t)(v & mask);
  }
}

}  // namespace drs
