/*
 * drs_b200.h — C-ABI of the B200 accelerator backend for DeepRecSys.
 *
 * This is the drop-in boundary for the per-query accelerator inference path.
 * The reference has no native code; the interface it would bind is the python
 * hand-off between the accel engine process and the model net:
 *
 *   accelInferenceEngine.py:46-83   per-packet serve loop (today: sleep)
 *   inferenceEngine.py:200-215      request -> (dense, lengths, ids) slicing
 *   models/dlrm_s_caffe2.py:162-174 DLRM_Wrapper.run_queues(ids, lengths, fc, B)
 *   models/dlrm_s_caffe2.py:549-569 DLRM_Net.run(X, S_lengths, S_indices)
 *   models/dlrm_s_caffe2.py:281-329 create_emb   -> SparseLengthsSum per table
 *   models/dlrm_s_caffe2.py:223-279 create_mlp   -> FC + Relu/Sigmoid per layer
 *   models/dlrm_s_caffe2.py:331-365 create_interactions (cat | dot)
 *   models/{wide_and_deep,ncf,din}.py  sibling graphs
 *
 * Rules of the boundary: extern "C", plain pointers and sizes, caller-owned
 * host buffers, engine-owned device memory, int status (0 = ok) plus a
 * thread-local message from drs_last_error().  One submitting thread per
 * engine handle.  There is no CPU fallback behind any entry point.
 */
#ifndef DRS_B200_H_
#define DRS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRS_ABI_VERSION 7

/* status codes */
enum {
  DRS_OK = 0,
  DRS_ERR_INVALID = 1,     /* bad argument / spec (mirrors the reference's sys.exit checks) */
  DRS_ERR_CUDA = 2,        /* CUDA runtime failure, message has the cudaError string */
  DRS_ERR_STATE = 3,       /* weights/tables/batch not loaded yet */
  DRS_ERR_INDEX = 4,       /* SparseLengthsSum enforce: idx out of range or sum(len) != len(idx) */
  DRS_ERR_COMM = 5,        /* peer-memory / multi-GPU exchange setup failure */
  DRS_ERR_HOST = 6         /* host-side failure inside the library (out of memory, any C++
                              exception): caught at the boundary, never thrown through it */
};

/* model graphs (reference: models/dlrm_s_caffe2.py, wide_and_deep.py, ncf.py, din.py,
 * multi_task_wnd.py) */
enum { DRS_MODEL_DLRM = 0, DRS_MODEL_WND = 1, DRS_MODEL_NCF = 2, DRS_MODEL_DIN = 3,
       DRS_MODEL_MTWND = 4, DRS_MODEL_DIEN = 5 /* models/dien.py */ };

/* arch_interaction_op (dlrm_s_caffe2.py:334-365) */
enum { DRS_INTERACT_CAT = 0, DRS_INTERACT_DOT = 1 };

/* MLP arithmetic: FP32 = FFMA parity path; TF32X3 = tcgen05 error-compensated
 * (fp32-class accuracy); TF32 = tcgen05 single pass (throughput mode). */
enum { DRS_PREC_FP32 = 0, DRS_PREC_TF32X3 = 1, DRS_PREC_TF32 = 2 };

/* multi-GPU sharding of the embedding tables (one process per GPU).
 * TABLE: table t lives on rank t % world.  ROW: every table is split by rows, rank r holds rows
 * [r*chunk, min(N, (r+1)*chunk)), chunk = ceil(N / world). */
enum { DRS_SHARD_TABLE = 0, DRS_SHARD_ROW = 1 };

/* mlp ids for drs_load_fc */
enum {
  DRS_MLP_BOT = 0,         /* DLRM bottom MLP (dlrm_s_caffe2.py:376)                        */
  DRS_MLP_TOP = 1,         /* DLRM/WnD/DIN top MLP; NCF: the MLP tower ln_top[:-1] (ncf.py:332) */
  DRS_MLP_FINAL = 2,       /* NCF final FC over [Zmf, tower] (ncf.py:339-343)               */
  DRS_MLP_ATTN_BASE = 16,  /* DIN attention unit i -> DRS_MLP_ATTN_BASE + i (din.py:260-278) */
  DRS_MLP_TASK_BASE = 4096,/* MT-WnD task head i -> DRS_MLP_TASK_BASE + i (multi_task_wnd.py:304-316) */
  DRS_MLP_RNN_BASE = 8192  /* DIEN recurrent cell l (0, 1) -> DRS_MLP_RNN_BASE + l; its "layer" 0 is the
                              input FC i2h [H, in], "layer" 1 the recurrent FC gates_t [H, H]
                              (rnn_cell.BasicRNN as called at dien.py:318-328, 370-378) */
};

/* stage taps for drs_fetch_stage (stage-level parity) */
enum {
  DRS_STAGE_POOLED = 0,    /* [B, T*D] pooled vectors, table-major columns   */
  DRS_STAGE_BOT = 1,       /* layer k output of the bottom MLP [B, ln_bot[k+1]] (DIN: summed attention output,
                              DIEN: last hidden state of the second recurrent cell [B, H]; layer ignored) */
  DRS_STAGE_INTERACT = 2,  /* top-MLP input [B, ln_top[0]]                   */
  DRS_STAGE_TOP = 3,       /* layer k output of the top MLP                  */
  DRS_STAGE_OUT = 4,       /* final output [B, out_dim]                      */
  DRS_STAGE_OFFSETS = 5,   /* int32 offsets [T, B+1] reinterpret as float bits (bit-exact bucketing check) */
  DRS_STAGE_TASK = 6,      /* MT-WnD: layer (task * 256 + k) = output of layer k of task head `task` */
  DRS_STAGE_IDS = 7        /* int32 ids the gather reads for table `layer` ([1, n] as float bits); row-sharded
                              mode: this rank's bucketed, rebased ids (bit-exact bucketing check) */
};

typedef struct drs_engine drs_engine;

/* Integers of the model spec (utils/utils.py:22-37 flags after the per-model
 * adjustments of dlrm_s_caffe2.py:404-440, wide_and_deep.py:332-350,
 * ncf.py:373-388, din.py:346-359). */
typedef struct {
  int32_t model_type;            /* DRS_MODEL_*                                              */
  int32_t n_tables;              /* T                                                        */
  const int64_t* table_rows;     /* [T] rows per table                                       */
  int32_t emb_dim;               /* D = arch_sparse_feature_size                             */
  int32_t n_bot;                 /* number of widths in ln_bot (layers + 1); 0 if none       */
  const int32_t* ln_bot;         /* DLRM: bottom widths; WnD: [m_den]; DIN: attention [3D,..,D] */
  int32_t n_top;                 /* number of widths in ln_top                               */
  const int32_t* ln_top;         /* adjusted: ln_top[0] = interaction width                  */
  int32_t interaction_op;        /* DRS_INTERACT_*  (DLRM only)                              */
  int32_t interaction_itself;    /* arch_interaction_itself                                  */
  int32_t max_batch;             /* largest query (samples)                                  */
  int32_t max_lookups_per_table; /* largest sum(lengths) of one table in one query           */
  int32_t precision;             /* DRS_PREC_*                                               */
  int32_t n_lanes;               /* concurrent query lanes (stream + activation set); 0 = default 2 */
  /* MT-WnD only (multi_task_wnd.py:304-316, :358-362); zero for the other graphs */
  int32_t n_tasks;               /* num_multi_tasks: task heads above the shared top MLP     */
  int32_t n_task;                /* number of widths in ln_task                              */
  const int32_t* ln_task;        /* arch_mlp_tasks; ln_task[0] == ln_top[n_top-1]            */
  int32_t task_sigmoid_layer;    /* 1-based task layer that gets Sigmoid (the reference passes
                                    sigmoid_top = len(ln_top) - 1 to every head), 0 = none    */
  int32_t hidden_size;           /* DIEN: width H of the two recurrent cells (--hidden_size)   */
  /* multi-GPU (zero = single GPU or table-wise sharding decided at drs_shard_connect) */
  int32_t shard_mode;            /* DRS_SHARD_*; ROW must be declared here (sizes the partial slab) */
  int32_t shard_world;           /* ROW: number of ranks (2..8)                              */
} drs_spec;

/* per-call timing, milliseconds */
typedef struct {
  float h2d_ms;      /* host->device input copies (0 for staged batches) */
  float gpu_ms;      /* kernels, CUDA-event time on the engine stream    */
  float d2h_ms;      /* result read-back                                  */
  float total_ms;    /* wall clock of the call                            */
} drs_timing;

const char* drs_last_error(void);
int drs_abi_version(void);
int drs_device_count(int* n_out);
/* NUMA node of the host bridge GPU `device` hangs off (from its PCI bus id via sysfs), or -1 when
 * the platform does not say.  Serving processes bind themselves (CPU affinity + first-touch of
 * their pinned staging buffers) to that node so that H2D copies do not cross sockets. */
int drs_device_numa_node(int device, int* node_out);

/* lifecycle */
int drs_create(const drs_spec* spec, int device, drs_engine** out);
int drs_destroy(drs_engine* e);
int drs_out_dim(const drs_engine* e, int* out_dim);

/* parameters: host -> HBM.  W is [out_dim, in_dim] row-major exactly as the
 * Caffe2 FC blob (dlrm_s_caffe2.py:247), b is [out_dim]. */
int drs_load_table(drs_engine* e, int table, const float* rows, int64_t n_rows, int dim);
int drs_load_fc(drs_engine* e, int mlp_id, int layer, const float* W, const float* b,
                int out_dim, int in_dim);

/* Pre-stage one generated batch in HBM (the reference pre-generates all
 * batches at engine start: inferenceEngine.py:80-88).  idx[t] has
 * sum(lengths[t][0..n_samples)) entries, int64 (as np.array(list) yields) or
 * int32 (after the reference's Cast, dlrm_s_caffe2.py:308).  dense may be NULL
 * for models without dense input (NCF, DIN). */
int drs_stage_batch(drs_engine* e, int batch_id, const float* dense,
                    const void* const* idx, int idx_is_i64,
                    const int32_t* const* lengths, int n_samples);

/* Serve "first batch_size samples of staged batch batch_id"
 * (inferenceEngine.py:200-215).  out may be NULL (result stays in HBM). */
int drs_infer_staged(drs_engine* e, int batch_id, int batch_size, float* out, drs_timing* t);

/* Serve a query from HOST buffers: H2D of dense/ids/lengths, lengths->offsets
 * on the GPU, gather, interaction, MLPs, D2H of [batch_size, out_dim]. */
int drs_infer(drs_engine* e, const float* dense, const void* const* idx, int idx_is_i64,
              const int32_t* const* lengths, int batch_size, float* out, drs_timing* t);

/* Asynchronous halves used by the serving loop and the bench harness:
 * enqueue does not synchronise; drs_sync waits for the stream and reports the
 * SparseLengthsSum enforce flag. */
int drs_enqueue_staged(drs_engine* e, int batch_id, int batch_size);
int drs_enqueue_host(drs_engine* e, const float* dense, const void* const* idx, int idx_is_i64,
                     const int32_t* const* lengths, int batch_size, float* out_pinned);
int drs_sync(drs_engine* e);

/* Packed ids (opt-in wire format for the upload path).  The ids of an n-row table need
 * ceil(log2(n)) bits, not 32: RMC1's 4 M-row tables -> 22 bits, 31 % fewer bytes over PCIe, which
 * is what bounds the upload path (SURVEY hard part 2).  drs_pack_ids writes id i into bits
 * [i*bits, (i+1)*bits) of a little-endian uint32 stream of drs_packed_words(n, bits) words (the
 * client's job, like producing int32 instead of int64 today); drs_enqueue_host_packed takes one
 * such stream per table (bits[t] >= ceil(log2(rows[t])), <= 31) and widens them on the GPU to
 * the int32 ids the gather reads — same Cast-to-int32 semantics as dlrm_s_caffe2.py:308. */
int64_t drs_packed_words(int64_t n_ids, int bits);
int drs_pack_ids(const int32_t* ids, int64_t n_ids, int bits, uint32_t* out);
int drs_enqueue_host_packed(drs_engine* e, const float* dense, const uint32_t* const* packed,
                            const int32_t* bits, const int32_t* const* lengths, int batch_size,
                            float* out_pinned);

/* Engine-side packing: drs_set_option(e, "host_pack", n_threads) makes every host-path call
 * (drs_infer / drs_enqueue_host / drs_submit_host, int32 or int64 ids — the dtype the reference
 * hands run_queues, dlrm_s_caffe2.py:223-226,308) pack the ids of tables with <= 2^24 rows on
 * n host threads into pinned staging, table by table, each table's copy leaving while the next is
 * packed; queries with fewer than "host_pack_min_ids" lookups (default 1,048,576) go up unpacked.
 * Results and the DRS_ERR_INDEX behaviour are those of the unpacked path (a table holding an id
 * that is negative or past 2^bits is uploaded as it came, and the gather raises the enforce).
 * The wire layout is engine-internal (lane-interleaved 512-id blocks, csrc/hostpack.h);
 * drs_hostpack_ids exposes the packer for tests and timing: it packs n int32/int64 ids into
 * drs_hostpack_words(n, bits) words on `threads` threads; *ok = 0 when some id does not fit. */
int64_t drs_hostpack_words(int64_t n_ids, int bits);
int drs_hostpack_ids(const void* ids, int idx_is_i64, int64_t n_ids, int bits, int threads,
                     uint32_t* out, int* ok);

/* Submit / wait pair for pipelined serving: up to drs_lane_count() queries may be in flight
 * (the copy and MLP kernels of one overlap the gather of the next).  out_pinned (nullable) must
 * be pinned host memory and stay valid until drs_wait(ticket) returns; drs_wait reports the
 * SparseLengthsSum enforce flag of that query's lane. */
int drs_submit_staged(drs_engine* e, int batch_id, int batch_size, float* out_pinned,
                      int64_t* ticket);
int drs_submit_host(drs_engine* e, const float* dense, const void* const* idx, int idx_is_i64,
                    const int32_t* const* lengths, int batch_size, float* out_pinned,
                    int64_t* ticket);
/* Coalesced serving ("multi-query batching on the accel queue"): n <= 32
 * queued requests, request k = "the first batch_sizes[k] samples of staged batch batch_ids[k]"
 * (inferenceEngine.py:200-215), are served as ONE stacked query: one K1 launch pools all of
 * them into consecutive row ranges and the MLPs run once over sum(batch_sizes) <= max_batch
 * rows.  Row r of request k lands at out_pinned[(sum_{j<k} batch_sizes[j] + r) * out_dim].
 * Results are the same values as n separate drs_submit_staged calls (pooled vectors
 * bit-identical; every sample's MLP arithmetic is independent of its row position). */
int drs_submit_many(drs_engine* e, int n, const int32_t* batch_ids, const int32_t* batch_sizes,
                    float* out_pinned, int64_t* ticket);
int drs_wait(drs_engine* e, int64_t ticket);
int drs_lane_count(const drs_engine* e, int* n_lanes);

/* CUDA-event stopwatch on the engine's stream (the stream the kernels run on). */
int drs_timer_start(drs_engine* e);
int drs_timer_stop(drs_engine* e, float* elapsed_ms);

/* Operator entry: SparseLengthsSum(table, idx, lengths) -> out[n_segments, D]
 * on loaded table `table` with host idx/lengths (Caffe2 op semantics used at
 * dlrm_s_caffe2.py:319-325). */
int drs_sparse_lengths_sum(drs_engine* e, int table, const void* idx, int idx_is_i64,
                           int64_t n_idx, const int32_t* lengths, int n_segments, float* out);

/* Debug taps of the last served query (device -> host). */
int drs_fetch_stage(drs_engine* e, int stage, int layer, float* dst, int64_t capacity_floats,
                    int* rows, int* cols);

/* Bookkeeping the bench reports. */
int drs_launch_count(const drs_engine* e, int64_t* n_kernel_launches);
int drs_hbm_bytes(const drs_engine* e, int64_t* bytes_allocated);

/* Runtime options.  "mlp_chain" (default 1): fuse whole FC stacks into one tcgen05 kernel when
 * the shapes allow (tensor-core precisions only).  "tap_hidden" (default 0): in chain mode also
 * store the hidden activations so that drs_fetch_stage can return them.  "gather_ctas_per_sm"
 * (default 4): K1 runs as a persistent kernel with that many CTAs per SM (0 = one CTA per tile).
 * "cuda_graphs" (default 1): unsharded engines replay a captured CUDA graph for a staged step they
 * have served before (same lane, batch id and size) — one cudaGraphLaunch instead of 3-4 kernel
 * launches, which is what bounds the rate of small queries.  "exchange" (default 0): 1 selects
 * the unfused exchange baseline of drs_exchange_begin / drs_exchange_end.  "tc_persistent"
 * (default 1): wide FC layers whose tile grid exceeds one wave run the persistent tcgen05 kernel
 * (two TMEM accumulators, epilogue of tile i under the MMAs of tile i+1); 0 = one CTA per tile,
 * n >= 2 = always persistent with at most n CTAs.  "short_fixed" (default 1): queries whose
 * segments all have the same 1..4 lookups (the reference's request slicing requires fixed L) are
 * gathered by the flat fixed-length kernel instead of the tiled one.  "din_fused" (default 1): DIN's
 * attention units, their Sum and the Concat into the top-MLP input run as one kernel when the
 * attention MLP's widths are <= 32.  "rnn_tc" (default 1): DIEN's recurrent cells run on tcgen05
 * in the tensor-core precisions (H and D multiples of 16).  Setting any of the last three to 0
 * selects the per-operator graph they replace (same values within the stated tolerances). */
int drs_set_option(drs_engine* e, const char* key, int64_t value);

/* Per-segment CUDA-event profiling of the served graph, on the engine stream:
 * between begin and end every query records 4 events (start | before K1 |
 * after K1 | end); end returns the summed times of the pre-gather kernels
 * (bottom MLP / copies), the K1 gather(+interaction) kernel, and the
 * post-gather kernels (top MLP ...) over n_steps queries. */
int drs_profile_begin(drs_engine* e, int max_steps);
int drs_profile_end(drs_engine* e, int* n_steps, float* pre_ms, float* gather_ms, float* post_ms);

/* Pinned host memory for the upload path (so H2D/D2H are true async DMA). */
int drs_host_alloc(void** ptr, int64_t bytes);
int drs_host_free(void* ptr);
/* Page-lock memory the caller already owns — e.g. a shared-memory ring another process writes
 * query tensors into (deeprecsys_b200/utils/input_ring.py) — so that drs_submit_host /
 * drs_enqueue_host copy from it by asynchronous DMA.  Optional: unregistered (pageable) buffers
 * are accepted by every entry point, the copy is then staged by the driver. */
int drs_host_register(void* ptr, int64_t bytes);
int drs_host_unregister(void* ptr);

/* Multi-GPU, one process per GPU.
 * Row-wise sharding (spec.shard_mode = DRS_SHARD_ROW): connect first, then drs_load_table with
 * the FULL table (the rank keeps only its row slice).  Per query every rank buckets the ids by
 * its row range on the GPU (integer-exact, order-preserving), gathers PARTIAL pooled vectors for
 * all samples and stores them into the owner rank's partial slab; the owner adds the `world`
 * partials in rank order.  The fp32 sum is therefore reassociated relative to one GPU
 * (rtol 1e-6), while table-wise sharding below stays bit-exact.
 *
 * Table-wise sharding (default):
 * Each rank owns tables {t : t % world == rank}; after the gather every rank
 * holds all T pooled vectors for its slice of the query's samples.  The
 * exchange is done by the gather kernel itself storing pooled rows straight
 * into the destination rank's slab through peer-mapped memory (CUDA IPC over
 * NVLink), followed by a flag barrier — no pack/unpack kernels.
 *   1. every rank: drs_shard_export(e, handle_out)   (64-byte IPC handle + metadata)
 *   2. ranks all-gather the blobs (torch.distributed / any transport)
 *   3. every rank: drs_shard_connect(e, rank, world, all_blobs)
 */
#define DRS_SHARD_BLOB_BYTES 192
int drs_shard_export(drs_engine* e, void* blob_out);
int drs_shard_connect(drs_engine* e, int rank, int world, const void* all_blobs);

/* Unfused exchange BASELINE for table-wise sharding (SURVEY 2a/8e name a single NCCL all-to-all;
 * the product fuses the exchange into K1's stores instead, see above).  After
 * drs_set_option(e, "exchange", 1) a staged DLRM query is served in two halves around a
 * collective the caller runs on the returned stream:
 *   drs_exchange_begin: bottom MLP + K1 -> pooled vectors into a local send slab
 *                       [world][ceil(B/world)][ceil(T/world)*D] (block q = rank q's samples)
 *   caller:             all-to-all of floats_per_peer floats per peer, send -> recv, on `stream`
 *                       (bench.py: torch.distributed.all_to_all_single over NCCL)
 *   drs_exchange_end:   unpack the received blocks into the interaction layout + top MLP
 * It exists so that the cost of not fusing can be measured next to the fused path. */
int drs_exchange_begin(drs_engine* e, int batch_id, int batch_size, void** send, void** recv,
                       int64_t* floats_per_peer, void** stream);
int drs_exchange_end(drs_engine* e);

/* Queue wire codec of the serving loop (host only; csrc/wire.cpp).  The reference's serving
 * protocol is pickled packets over multiprocessing Queues: a frame on the pipe is
 * struct.pack("!i", len) + pickle.dumps(obj) (DeepRecSys.py:44-66, loadGenerator.py:162-177,
 * accelInferenceEngine.py:46-83, utils/packets.py:6-59).  One record per packet, the layout of
 * deeprecsys_b200/utils/response_ring.py REC: `kinds` holds two bits per field (epoch, batch_id,
 * batch_size, sub_id, total_sub_batches, exp_packet, arrival_time): 0 int / float, 1 bool,
 * 2 None.
 *   drs_wire_parse_requests: decodes the leading frames of buf[0, len) that are exactly a pickled
 *     `ServiceRequest` with int / bool / None fields (arrival_time: float / None) into recs;
 *     `modules` = the module paths the class may come from, NUL-separated and ended by an empty
 *     name ("utils.packets\0deeprecsys_b200.utils.packets\0"; NULL = any) — a frame naming
 *     another module is left to pickle, which knows whether that module exists;
 *     *consumed = bytes of the decoded frames; *stop says why it stopped: NEED_MORE (the rest is
 *     an incomplete frame), FULL (max_recs reached), PYTHON (the next frame is something else —
 *     the final None, a request carrying tensors, an unexpected type: the caller unpickles it).
 *   drs_wire_emit_responses: recs -> frames holding pickled `ServiceResponse` objects
 *     (protocol 4).  class_ops = the pickle opcodes that push the class (SHORT_BINUNICODE module,
 *     MEMOIZE, SHORT_BINUNICODE qualname, MEMOIZE, STACK_GLOBAL, MEMOIZE); field_order[12] = the
 *     attribute order of the class' __dict__ as indices into (consumer_id, epoch, batch_id,
 *     batch_size, arrival_time, queue_start_time, queue_end_time, inference_end_time,
 *     out_batch_size, total_sub_batches, exp_packet, sub_id); queue_end_time and
 *     inference_end_time are both rec.end_time, queue_start_time is rec.start_time
 *     (accelInferenceEngine.py:68-80). */
typedef struct drs_wire_rec {
  int64_t epoch, batch_id, batch_size, sub_id, total_sub_batches, out_batch_size, exp_packet,
      kinds;
  double arrival_time, start_time, end_time;
} drs_wire_rec;
enum { DRS_WIRE_NEED_MORE = 0, DRS_WIRE_PYTHON = 1, DRS_WIRE_FULL = 2 };
int drs_wire_parse_requests(const uint8_t* buf, int64_t len, const char* modules,
                            drs_wire_rec* recs, int max_recs, int32_t* n_recs, int64_t* consumed,
                            int32_t* stop);
int drs_wire_emit_responses(const drs_wire_rec* recs, int n, const uint8_t* class_ops,
                            int class_ops_len, const int32_t* field_order, int64_t consumer_id,
                            uint8_t* out, int64_t cap, int64_t* out_len);

#ifdef __cplusplus
}
#endif
#endif  /* DRS_B200_H_ */
