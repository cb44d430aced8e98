// engine.cu — C-ABI implementation (include/drs_b200.h) of the B200 accel path.
//
// Graph construction per model type follows the reference builders:
//   DLRM  /root/reference/models/dlrm_s_caffe2.py:367-389
//   WnD   models/wide_and_deep.py:281-305
//   NCF   models/ncf.py:198-346
//   DIN   models/din.py:246-331
// but executes as: [lengths->offsets] -> bottom MLP -> K1 gather(+cat|dot) ->
// top MLP on one CUDA stream, all activations resident in HBM buffers owned by
// the engine.  No CPU fallback: every compute entry point launches kernels.
#include "../../include/drs_b200.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "interact.cuh"
#include "mlp_ffma.cuh"
#include "mlp_tc.cuh"
#include "sls_gather.cuh"

using namespace drs;

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t err__ = (call);                                                               \
    if (err__ != cudaSuccess)                                                                 \
      return fail(DRS_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                  cudaGetErrorString(err__));                                                 \
  } while (0)

#define CHECK_ARG(cond, ...) \
  do {                       \
    if (!(cond)) return fail(DRS_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != DRS_OK) return rc__; \
  } while (0)

// ------------------------------------------------------------------ engine
namespace {

struct Layer {
  float* W = nullptr;   // [n_units, out, in]
  float* b = nullptr;   // [n_units, out]
  float* Wsplit = nullptr;  // tensor-core path: [n_units][2][out][in] = (top-19-bit part, remainder)
  int out = 0, in = 0;
  int n_units = 1;
  int act = kActRelu;
  std::vector<char> loaded;   // per unit
  float* y = nullptr;   // own output buffer [n_units, max_batch, out] unless redirected
};

struct Mlp {
  std::vector<Layer> layers;
  bool all_loaded() const {
    for (const Layer& l : layers)
      for (char c : l.loaded)
        if (!c) return false;
    return true;
  }
};

struct InputSet {                 // one query's inputs in HBM
  float* dense = nullptr;         // [n, m_den]
  int32_t* idx = nullptr;         // slab; table t at idx + idx_pos[t]
  int32_t* lengths = nullptr;     // [T, n]
  int32_t* off = nullptr;         // [T, n+1]
  int64_t* idx64 = nullptr;       // scratch for int64 uploads (host path only)
  SlsTableDesc* descs = nullptr;  // [T] device
  ScanTableDesc* scan = nullptr;  // [T] device
  std::vector<int64_t> idx_pos;   // per table start (ints, multiple of 4)
  std::vector<int64_t> idx_cap;   // per table capacity (ints)
  int n = 0;                      // samples capacity / staged count
  bool valid = false;
};

}  // namespace

struct drs_engine {
  drs_spec spec{};
  std::vector<int64_t> table_rows;
  std::vector<int32_t> ln_bot, ln_top;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_timer[2] = {nullptr, nullptr};

  int T = 0, D = 0, m_den = 0, out_dim = 0, ld_R = 0, n_units = 0;
  std::vector<float*> tables;
  std::vector<char> table_loaded;
  Mlp bot, top, fin;          // DIN: bot = attention MLP batched over n_units

  // activations
  float* R = nullptr;         // top-MLP input [maxB, ld_R]
  float* P = nullptr;         // pooled [maxB, T*D] (models that need it / dot tap)
  float* dense_out = nullptr; // DLRM-dot: bottom output [maxB, D]
  float* F = nullptr;         // NCF: [maxB, D + ln_top[-2]]
  float* C = nullptr;         // DIN: [n_units, maxB, 3D]
  float* out = nullptr;       // [maxB, out_dim] (alias of last layer's y)

  std::map<int, InputSet> staged;
  InputSet slot;              // host-path inputs
  const InputSet* last_inputs = nullptr;
  int last_B = 0;

  int* d_err = nullptr;
  int* h_err = nullptr;       // pinned
  int64_t launches = 0;
  int64_t hbm_bytes = 0;

  // gather launch geometry
  int lpr = 0, SB = 1, idx_cap = 0;
  size_t sls_smem = 0;

  // per-segment CUDA-event profiling (drs_profile_begin/end)
  std::vector<cudaEvent_t> prof_ev;   // 4 per step: start, after pre-gather, after gather, end
  int prof_cap = 0, prof_n = 0;
  bool prof_on = false;
};

namespace {

int dev_alloc(drs_engine* e, void** p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  CU(cudaMalloc(p, bytes));
  e->hbm_bytes += (int64_t)bytes;
  return DRS_OK;
}

template <typename T>
int dev_alloc_t(drs_engine* e, T** p, size_t count) {
  return dev_alloc(e, reinterpret_cast<void**>(p), count * sizeof(T));
}

int alloc_mlp_buffers(drs_engine* e, Mlp& m, const std::vector<int32_t>& ln, int n_units,
                      int last_act) {
  const int nl = (int)ln.size() - 1;
  m.layers.resize(nl > 0 ? nl : 0);
  for (int i = 0; i < nl; ++i) {
    Layer& l = m.layers[i];
    l.in = ln[i];
    l.out = ln[i + 1];
    l.n_units = n_units;
    l.act = (i == nl - 1) ? last_act : kActRelu;
    l.loaded.assign(n_units, 0);
    TRY(dev_alloc_t(e, &l.W, (size_t)n_units * l.out * l.in));
    TRY(dev_alloc_t(e, &l.b, (size_t)n_units * l.out));
    if (e->spec.precision != DRS_PREC_FP32)
      TRY(dev_alloc_t(e, &l.Wsplit, (size_t)n_units * 2 * l.out * l.in));
    TRY(dev_alloc_t(e, &l.y, (size_t)n_units * e->spec.max_batch * l.out));
  }
  return DRS_OK;
}

int alloc_input_set(drs_engine* e, InputSet& s, int n, const std::vector<int64_t>& idx_counts,
                    bool with_i64_scratch) {
  const int T = e->T;
  s.n = n;
  s.idx_pos.assign(T, 0);
  s.idx_cap.assign(T, 0);
  int64_t pos = 0;
  for (int t = 0; t < T; ++t) {
    s.idx_pos[t] = pos;
    s.idx_cap[t] = idx_counts[t];
    pos += ((idx_counts[t] + 3) & ~int64_t(3)) + 4;   // 16-B aligned starts, TMA over-read pad
  }
  if (e->m_den > 0) TRY(dev_alloc_t(e, &s.dense, (size_t)n * e->m_den));
  TRY(dev_alloc_t(e, &s.idx, (size_t)pos + 4));
  CU(cudaMemsetAsync(s.idx, 0, ((size_t)pos + 4) * sizeof(int32_t), e->stream));
  if (with_i64_scratch) TRY(dev_alloc_t(e, &s.idx64, (size_t)pos + 4));
  TRY(dev_alloc_t(e, &s.lengths, (size_t)T * n));
  TRY(dev_alloc_t(e, &s.off, (size_t)T * (n + 1)));
  TRY(dev_alloc_t(e, &s.descs, (size_t)T));
  TRY(dev_alloc_t(e, &s.scan, (size_t)T));
  return DRS_OK;
}

void free_input_set(InputSet& s) {
  cudaFree(s.dense); cudaFree(s.idx); cudaFree(s.idx64); cudaFree(s.lengths);
  cudaFree(s.off); cudaFree(s.descs); cudaFree(s.scan);
  s = InputSet();
}

// destination column of table t's pooled vector inside the gather's output buffer
int gather_dst_col(const drs_engine* e, int t) {
  switch (e->spec.model_type) {
    case DRS_MODEL_DLRM: return e->D + t * e->D;           // R = [bot_out, e_0 .. e_{T-1}]
    case DRS_MODEL_WND:  return e->m_den + t * e->D;       // R = [dense_in, e_0 ..]
    default:             return t * e->D;                  // P = [e_0 .. e_{T-1}]
  }
}

int write_descs(drs_engine* e, InputSet& s, const std::vector<int64_t>& n_idx) {
  std::vector<SlsTableDesc> d(e->T);
  std::vector<ScanTableDesc> sc(e->T);
  for (int t = 0; t < e->T; ++t) {
    d[t].rows = e->tables[t];
    d[t].idx = s.idx + s.idx_pos[t];
    d[t].off = s.off + (size_t)t * (s.n + 1);
    d[t].n_rows = e->table_rows[t];
    d[t].dst_col = gather_dst_col(e, t);
    d[t].pad_ = 0;
    sc[t].lengths = s.lengths + (size_t)t * s.n;
    sc[t].off = s.off + (size_t)t * (s.n + 1);
    sc[t].n_idx = n_idx.empty() ? -1 : n_idx[t];
  }
  // pageable -> device copies of small structs are staged by the runtime before returning
  CU(cudaMemcpyAsync(s.descs, d.data(), d.size() * sizeof(SlsTableDesc), cudaMemcpyHostToDevice,
                     e->stream));
  CU(cudaMemcpyAsync(s.scan, sc.data(), sc.size() * sizeof(ScanTableDesc), cudaMemcpyHostToDevice,
                     e->stream));
  return DRS_OK;
}

// ------------------------------------------------------------ launchers
int launch_fc(drs_engine* e, const Layer& l, const float* X, int64_t ldx, int64_t x_bstride,
              int64_t x_rows_total, float* Y, int64_t ldy, int64_t y_bstride, int M) {
  if (M <= 0) return DRS_OK;
  if (e->spec.precision != DRS_PREC_FP32 && l.Wsplit && tc_fc_supported(l.out, l.in, ldx, X) &&
      (l.n_units == 1 || x_bstride % ldx == 0)) {
    TcFcLaunch q{};
    q.X = X; q.ldx = ldx; q.x_rows = x_rows_total;
    q.x_rows_per_unit = l.n_units > 1 ? x_bstride / ldx : 0;
    q.Wsplit = l.Wsplit;
    q.bias = l.b; q.b_bstride = l.out;
    q.Y = Y; q.ldy = ldy; q.y_bstride = y_bstride;
    q.M = M; q.N = l.out; q.K = l.in; q.act = l.act;
    q.split3 = (e->spec.precision == DRS_PREC_TF32X3);
    int rc = tc_fc_launch(q, l.n_units, e->stream);
    if (rc != 0) return fail(DRS_ERR_CUDA, "tensor-core FC launch failed: %s",
                             cudaGetErrorString((cudaError_t)rc));
    e->launches += 1;
    return DRS_OK;
  }
  FcLaunch p{};
  p.X = X; p.ldx = ldx; p.x_bstride = x_bstride;
  p.W = l.W; p.w_bstride = (int64_t)l.out * l.in;
  p.bias = l.b; p.b_bstride = l.out;
  p.Y = Y; p.ldy = ldy; p.y_bstride = y_bstride;
  p.M = M; p.N = l.out; p.K = l.in; p.act = l.act;
  dim3 grid(ceil_div(p.N, kFcBN), ceil_div(p.M, kFcBM), l.n_units);
  const bool vec = (l.in % 4 == 0) && (ldx % 4 == 0) && (x_bstride % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (vec) k_fc_ffma<true><<<grid, kFcThreads, 0, e->stream>>>(p);
  else     k_fc_ffma<false><<<grid, kFcThreads, 0, e->stream>>>(p);
  CU(cudaGetLastError());
  e->launches += 1;
  return DRS_OK;
}

// chain an MLP; the last layer may be redirected into a column block of a wider buffer
int run_mlp(drs_engine* e, Mlp& m, const float* X, int64_t ldx, int64_t x_rows_total, int M,
            float* last_y = nullptr, int64_t last_ldy = 0) {
  const int nl = (int)m.layers.size();
  const float* in = X;
  int64_t ld_in = ldx, bs_in = 0, rows_in = x_rows_total;
  for (int i = 0; i < nl; ++i) {
    Layer& l = m.layers[i];
    float* y = l.y;
    int64_t ldy = l.out, bsy = (int64_t)e->spec.max_batch * l.out;
    if (i == nl - 1 && last_y) { y = last_y; ldy = last_ldy; bsy = 0; }
    TRY(launch_fc(e, l, in, ld_in, bs_in, rows_in, y, ldy, bsy, M));
    in = y; ld_in = ldy; bs_in = bsy; rows_in = e->spec.max_batch;
  }
  return DRS_OK;
}

template <int LPR, bool DOT>
int launch_gather_t(drs_engine* e, const SlsLaunch& p, int grid, size_t smem) {
  static bool attr_set = false;
  if (!attr_set) {
    CU(cudaFuncSetAttribute(k_sls_gather<LPR, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            160 * 1024));
    attr_set = true;
  }
  k_sls_gather<LPR, DOT><<<grid, kSlsThreads, smem, e->stream>>>(p);
  CU(cudaGetLastError());
  e->launches += 1;
  return DRS_OK;
}

template <bool DOT>
int launch_gather_lpr(drs_engine* e, const SlsLaunch& p, int lpr, int grid, size_t smem) {
  switch (lpr) {
    case 1:  return launch_gather_t<1, DOT>(e, p, grid, smem);
    case 2:  return launch_gather_t<2, DOT>(e, p, grid, smem);
    case 4:  return launch_gather_t<4, DOT>(e, p, grid, smem);
    case 8:  return launch_gather_t<8, DOT>(e, p, grid, smem);
    case 16: return launch_gather_t<16, DOT>(e, p, grid, smem);
    case 32: return launch_gather_t<32, DOT>(e, p, grid, smem);
    default: return launch_gather_t<0, DOT>(e, p, grid, smem);
  }
}

void plan_gather(int T, int D, int64_t avg_len, bool dot, int* lpr_out, int* SB_out,
                 int* idx_cap_out, size_t* smem_out) {
  int lpr = 0;
  if (D % 4 == 0) {
    const int q = D / 4;
    if (q == 1 || q == 2 || q == 4 || q == 8 || q == 16 || q == 32) lpr = q;
  }
  const int groups = lpr > 0 ? kSlsThreads / lpr : kSlsThreads / kWarp;
  int SB = groups / (T > 0 ? T : 1);
  if (SB < 1) SB = 1;
  if (SB > 16) SB = 16;
  int idx_cap = 0;
  if (avg_len >= 8) {
    int64_t want = ((int64_t)SB * avg_len + 3 + 3) & ~int64_t(3);
    if (sls_smem_bytes(T, SB, (int)want, D, dot) <= 96 * 1024) idx_cap = (int)want;
  }
  *lpr_out = lpr; *SB_out = SB; *idx_cap_out = idx_cap;
  *smem_out = sls_smem_bytes(T, SB, idx_cap, D, dot);
}

int launch_gather(drs_engine* e, const InputSet& s, int B, float* out, int64_t ld_out, bool dot,
                  float* pooled_tap) {
  SlsLaunch p{};
  p.tables = s.descs; p.T = e->T; p.B = B; p.D = e->D;
  p.out = out; p.ld_out = ld_out;
  p.SB = e->SB; p.idx_cap = e->idx_cap;
  p.err_flag = e->d_err;
  p.dense = e->dense_out; p.ld_dense = e->D;
  p.itself = e->spec.interaction_itself;
  p.pooled_tap = pooled_tap;
  bool vec = (ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int t = 0; t < e->T && vec; ++t) vec = (gather_dst_col(e, t) % 4 == 0);
  p.vec_store = vec ? 1 : 0;
  const int grid = ceil_div(B, e->SB);
  if (dot) return launch_gather_lpr<true>(e, p, e->lpr, grid, e->sls_smem);
  return launch_gather_lpr<false>(e, p, e->lpr, grid, e->sls_smem);
}

int launch_scan(drs_engine* e, const InputSet& s, int B) {
  k_lengths_to_offsets<<<e->T, kScanThreads, 0, e->stream>>>(s.scan, B, e->d_err);
  CU(cudaGetLastError());
  e->launches += 1;
  return DRS_OK;
}

int ew_grid(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------ model graphs
int check_ready(drs_engine* e) {
  for (int t = 0; t < e->T; ++t)
    if (!e->table_loaded[t]) return fail(DRS_ERR_STATE, "embedding table %d not loaded", t);
  if (!e->bot.all_loaded()) return fail(DRS_ERR_STATE, "bottom/attention MLP weights not loaded");
  if (!e->top.all_loaded()) return fail(DRS_ERR_STATE, "top MLP weights not loaded");
  if (!e->fin.all_loaded()) return fail(DRS_ERR_STATE, "final FC weights not loaded");
  return DRS_OK;
}

// profiling marks: k = 0 start, 1 before gather, 2 after gather, 3 end of graph
#define PROF_MARK(k)                                                              \
  do {                                                                            \
    if (e->prof_on && e->prof_n < e->prof_cap)                                    \
      CU(cudaEventRecord(e->prof_ev[(size_t)e->prof_n * 4 + (k)], e->stream));    \
  } while (0)

int enqueue_graph(drs_engine* e, const InputSet& s, int B) {
  const int D = e->D, T = e->T;
  PROF_MARK(0);
  switch (e->spec.model_type) {
    case DRS_MODEL_DLRM: {
      const bool dot = e->spec.interaction_op == DRS_INTERACT_DOT;
      if (dot) {
        TRY(run_mlp(e, e->bot, s.dense, e->m_den, s.n, B, e->dense_out, D));
        PROF_MARK(1);
        TRY(launch_gather(e, s, B, e->R, e->ld_R, true, e->P));
        PROF_MARK(2);
      } else {
        TRY(run_mlp(e, e->bot, s.dense, e->m_den, s.n, B, e->R, e->ld_R));
        PROF_MARK(1);
        TRY(launch_gather(e, s, B, e->R, e->ld_R, false, nullptr));
        PROF_MARK(2);
      }
      TRY(run_mlp(e, e->top, e->R, e->ld_R, e->spec.max_batch, B));
      break;
    }
    case DRS_MODEL_WND: {
      k_copy_cols<<<ew_grid((int64_t)B * e->m_den), 256, 0, e->stream>>>(s.dense, e->m_den, e->R,
                                                                         e->ld_R, B, e->m_den);
      CU(cudaGetLastError());
      e->launches += 1;
      PROF_MARK(1);
      TRY(launch_gather(e, s, B, e->R, e->ld_R, false, nullptr));
      PROF_MARK(2);
      TRY(run_mlp(e, e->top, e->R, e->ld_R, e->spec.max_batch, B));
      break;
    }
    case DRS_MODEL_NCF: {
      const int64_t ldp = (int64_t)T * D;
      const int ldf = D + e->top.layers.back().out;
      PROF_MARK(1);
      TRY(launch_gather(e, s, B, e->P, ldp, false, nullptr));
      PROF_MARK(2);
      k_add_cols<<<ew_grid((int64_t)B * D), 256, 0, e->stream>>>(e->P, e->P + D, ldp, e->F, ldf, B,
                                                                 D);
      CU(cudaGetLastError());
      e->launches += 1;
      TRY(run_mlp(e, e->top, e->P + 2 * D, ldp, e->spec.max_batch, B, e->F + D, ldf));
      TRY(run_mlp(e, e->fin, e->F, ldf, e->spec.max_batch, B));
      break;
    }
    case DRS_MODEL_DIN: {
      const int64_t ldp = (int64_t)T * D;
      const int nu = e->n_units;
      PROF_MARK(1);
      TRY(launch_gather(e, s, B, e->P, ldp, false, nullptr));
      PROF_MARK(2);
      k_din_concat<<<ew_grid((int64_t)nu * B * D), 256, 0, e->stream>>>(
          e->P, ldp, e->C, (int64_t)e->spec.max_batch * 3 * D, B, D, nu, T - 2);
      CU(cudaGetLastError());
      e->launches += 1;
      // attention MLPs, batched over units (each unit has its own weights)
      {
        const int nl = (int)e->bot.layers.size();
        const float* in = e->C;
        int64_t ld_in = 3 * D, bs_in = (int64_t)e->spec.max_batch * 3 * D;
        for (int i = 0; i < nl; ++i) {
          Layer& l = e->bot.layers[i];
          const int64_t bsy = (int64_t)e->spec.max_batch * l.out;
          TRY(launch_fc(e, l, in, ld_in, bs_in, (int64_t)nu * e->spec.max_batch, l.y, l.out, bsy,
                        B));
          in = l.y; ld_in = l.out; bs_in = bsy;
        }
      }
      {
        const Layer& last = e->bot.layers.back();
        k_sum_units<<<ew_grid((int64_t)B * D), 256, 0, e->stream>>>(
            last.y, (int64_t)e->spec.max_batch * D, e->R + D, e->ld_R, B, D, nu);
        CU(cudaGetLastError());
        e->launches += 1;
      }
      const int cols[3] = {0, T - 2, T - 1};       // profile, candidate ad, context
      const int dsts[3] = {0, 2 * D, 3 * D};
      for (int k = 0; k < 3; ++k) {
        k_copy_cols<<<ew_grid((int64_t)B * D), 256, 0, e->stream>>>(
            e->P + (int64_t)cols[k] * D, ldp, e->R + dsts[k], e->ld_R, B, D);
        CU(cudaGetLastError());
        e->launches += 1;
      }
      TRY(run_mlp(e, e->top, e->R, e->ld_R, e->spec.max_batch, B));
      break;
    }
    default:
      return fail(DRS_ERR_INVALID, "unknown model_type %d", e->spec.model_type);
  }
  PROF_MARK(3);
  if (e->prof_on && e->prof_n < e->prof_cap) e->prof_n += 1;
  CU(cudaMemcpyAsync(e->h_err, e->d_err, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  e->last_inputs = &s;
  e->last_B = B;
  return DRS_OK;
}

int finish_sync(drs_engine* e) {
  CU(cudaStreamSynchronize(e->stream));
  if (*e->h_err) {
    *e->h_err = 0;
    cudaMemsetAsync(e->d_err, 0, sizeof(int), e->stream);
    cudaStreamSynchronize(e->stream);
    return fail(DRS_ERR_INDEX,
                "SparseLengthsSum enforce failed: index out of range, negative length, or "
                "sum(lengths) != len(indices)");
  }
  return DRS_OK;
}

// upload one query into an InputSet (host -> device on the engine stream)
int upload_inputs(drs_engine* e, InputSet& s, const float* dense, const void* const* idx,
                  int idx_is_i64, const int32_t* const* lengths, int B,
                  std::vector<int64_t>* n_idx_out) {
  const int T = e->T;
  CHECK_ARG(B >= 0 && B <= s.n, "batch_size %d exceeds capacity %d", B, s.n);
  CHECK_ARG(idx && lengths, "idx/lengths pointer arrays must not be NULL");
  if (e->m_den > 0) {
    CHECK_ARG(dense != nullptr, "model has %d dense features but dense == NULL", e->m_den);
    CU(cudaMemcpyAsync(s.dense, dense, (size_t)B * e->m_den * sizeof(float),
                       cudaMemcpyHostToDevice, e->stream));
  }
  std::vector<int64_t>& n_idx = *n_idx_out;
  n_idx.assign(T, 0);
  for (int t = 0; t < T; ++t) {
    const int32_t* len = lengths[t];
    int64_t tot = 0;
    for (int b = 0; b < B; ++b) {
      if (len[b] < 0) return fail(DRS_ERR_INDEX, "negative length at table %d sample %d", t, b);
      tot += len[b];
    }
    CHECK_ARG(tot <= s.idx_cap[t], "table %d: %lld lookups exceed capacity %lld", t,
              (long long)tot, (long long)s.idx_cap[t]);
    n_idx[t] = tot;
  }
  // lengths: one strided copy when the caller's rows are equally spaced
  bool len_strided = T > 1;
  const ptrdiff_t lstride = T > 1 ? (const char*)lengths[1] - (const char*)lengths[0] : 0;
  for (int t = 1; t < T && len_strided; ++t)
    len_strided = ((const char*)lengths[t] - (const char*)lengths[t - 1]) == lstride;
  if (len_strided && lstride >= (ptrdiff_t)(B * sizeof(int32_t)) && B > 0) {
    CU(cudaMemcpy2DAsync(s.lengths, (size_t)s.n * sizeof(int32_t), lengths[0], (size_t)lstride,
                         (size_t)B * sizeof(int32_t), T, cudaMemcpyHostToDevice, e->stream));
  } else {
    for (int t = 0; t < T; ++t)
      CU(cudaMemcpyAsync(s.lengths + (size_t)t * s.n, lengths[t], (size_t)B * sizeof(int32_t),
                         cudaMemcpyHostToDevice, e->stream));
  }
  const size_t isz = idx_is_i64 ? sizeof(int64_t) : sizeof(int32_t);
  // ids: equal counts + equal host stride + equal device stride -> one 2-D copy
  bool idx_strided = T > 1 && !idx_is_i64;
  const ptrdiff_t istride = T > 1 ? (const char*)idx[1] - (const char*)idx[0] : 0;
  const int64_t dstride = T > 1 ? s.idx_pos[1] - s.idx_pos[0] : 0;
  for (int t = 1; t < T && idx_strided; ++t)
    idx_strided = ((const char*)idx[t] - (const char*)idx[t - 1]) == istride &&
                  (s.idx_pos[t] - s.idx_pos[t - 1]) == dstride && n_idx[t] == n_idx[0];
  if (idx_strided && n_idx[0] > 0 && istride >= (ptrdiff_t)(n_idx[0] * isz)) {
    CU(cudaMemcpy2DAsync(s.idx + s.idx_pos[0], (size_t)dstride * sizeof(int32_t), idx[0],
                         (size_t)istride, (size_t)n_idx[0] * sizeof(int32_t), T,
                         cudaMemcpyHostToDevice, e->stream));
  } else {
    for (int t = 0; t < T; ++t) {
      if (n_idx[t] == 0) continue;
      if (idx_is_i64) {
        CU(cudaMemcpyAsync(s.idx64 + s.idx_pos[t], idx[t], (size_t)n_idx[t] * sizeof(int64_t),
                           cudaMemcpyHostToDevice, e->stream));
        k_cast_i64_i32<<<ew_grid(n_idx[t]), 256, 0, e->stream>>>(s.idx64 + s.idx_pos[t],
                                                                 s.idx + s.idx_pos[t], n_idx[t]);
        CU(cudaGetLastError());
        e->launches += 1;
      } else {
        CU(cudaMemcpyAsync(s.idx + s.idx_pos[t], idx[t], (size_t)n_idx[t] * sizeof(int32_t),
                           cudaMemcpyHostToDevice, e->stream));
      }
    }
  }
  return DRS_OK;
}

}  // namespace

// ------------------------------------------------------------------ C-ABI
extern "C" {

const char* drs_last_error(void) { return g_last_error.c_str(); }
int drs_abi_version(void) { return DRS_ABI_VERSION; }

int drs_device_count(int* n_out) {
  CHECK_ARG(n_out, "n_out is NULL");
  CU(cudaGetDeviceCount(n_out));
  return DRS_OK;
}

int drs_host_alloc(void** ptr, int64_t bytes) {
  CHECK_ARG(ptr && bytes >= 0, "bad arguments");
  CU(cudaHostAlloc(ptr, (size_t)(bytes > 0 ? bytes : 16), cudaHostAllocDefault));
  return DRS_OK;
}

int drs_host_free(void* ptr) {
  if (ptr) CU(cudaFreeHost(ptr));
  return DRS_OK;
}

int drs_create(const drs_spec* spec, int device, drs_engine** out) {
  CHECK_ARG(spec && out, "spec/out is NULL");
  CHECK_ARG(spec->n_tables >= 1 && spec->table_rows, "need at least one embedding table");
  CHECK_ARG(spec->emb_dim >= 1, "emb_dim must be >= 1");
  CHECK_ARG(spec->max_batch >= 1, "max_batch must be >= 1");
  CHECK_ARG(spec->n_top >= 2 && spec->ln_top, "top MLP needs at least one layer");
  CHECK_ARG(spec->precision >= DRS_PREC_FP32 && spec->precision <= DRS_PREC_TF32,
            "unknown precision %d", spec->precision);
  for (int t = 0; t < spec->n_tables; ++t)
    CHECK_ARG(spec->table_rows[t] >= 1 && spec->table_rows[t] < (int64_t(1) << 31),
              "table %d: rows must be in [1, 2^31) (ids are int32 after the Cast)", t);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  CHECK_ARG(device >= 0 && device < ndev, "device %d not present (%d visible)", device, ndev);
  CU(cudaSetDevice(device));

  drs_engine* e = new drs_engine();
  e->spec = *spec;
  e->device = device;
  e->T = spec->n_tables;
  e->D = spec->emb_dim;
  e->table_rows.assign(spec->table_rows, spec->table_rows + spec->n_tables);
  if (spec->n_bot > 0 && spec->ln_bot) e->ln_bot.assign(spec->ln_bot, spec->ln_bot + spec->n_bot);
  e->ln_top.assign(spec->ln_top, spec->ln_top + spec->n_top);
  e->spec.table_rows = e->table_rows.data();
  e->spec.ln_bot = e->ln_bot.data();
  e->spec.ln_top = e->ln_top.data();
  const int T = e->T, D = e->D, maxB = spec->max_batch;

  auto bail = [&](int rc) { drs_destroy(e); return rc; };
#define TRYB(expr) do { int rc_ = (expr); if (rc_ != DRS_OK) return bail(rc_); } while (0)
#define CHECKB(cond, ...) do { if (!(cond)) return bail(fail(DRS_ERR_INVALID, __VA_ARGS__)); } while (0)
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(DRS_ERR_CUDA, "%s -> %s", #call, cudaGetErrorString(e_))); } while (0)

  CUB(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (auto& v : e->ev) CUB(cudaEventCreate(&v));
  for (auto& v : e->ev_timer) CUB(cudaEventCreate(&v));
  TRYB(dev_alloc_t(e, &e->d_err, 4));
  CUB(cudaMemset(e->d_err, 0, 16));
  CUB(cudaHostAlloc((void**)&e->h_err, 16, cudaHostAllocDefault));
  *e->h_err = 0;

  // --- dims and checks per model (mirroring the reference's sys.exit checks)
  bool dot = false;
  switch (spec->model_type) {
    case DRS_MODEL_DLRM: {
      CHECKB(e->ln_bot.size() >= 2, "DLRM needs a bottom MLP");
      e->m_den = e->ln_bot.front();
      CHECKB(e->ln_bot.back() == D,
             "arch-sparse-feature-size %d does not match last dim of bottom mlp %d", D,
             e->ln_bot.back());                                  // dlrm_s_caffe2.py:435
      const int F = T + 1;
      dot = spec->interaction_op == DRS_INTERACT_DOT;
      CHECKB(dot || spec->interaction_op == DRS_INTERACT_CAT, "interaction op not supported");
      const int num_int = dot ? (spec->interaction_itself ? F * (F + 1) / 2 : F * (F - 1) / 2) + D
                              : F * D;                           // :419-424
      CHECKB(e->ln_top[0] == num_int,
             "# of feature interactions %d does not match first dim of top mlp %d", num_int,
             e->ln_top[0]);                                      // :438
      e->ld_R = num_int;
      TRYB(alloc_mlp_buffers(e, e->bot, e->ln_bot, 1, kActRelu));       // sigmoid_bot = -1
      TRYB(alloc_mlp_buffers(e, e->top, e->ln_top, 1, kActSigmoid));    // sigmoid_top = last
      if (dot) {
        TRYB(dev_alloc_t(e, &e->dense_out, (size_t)maxB * D));
        TRYB(dev_alloc_t(e, &e->P, (size_t)maxB * T * D));
      }
      break;
    }
    case DRS_MODEL_WND: {
      CHECKB(e->ln_bot.size() == 1, "Wide&Deep has no bottom MLP layers (arch_mlp_bot = m_den)");
      e->m_den = e->ln_bot[0];
      CHECKB(e->ln_top[0] == T * D + e->m_den, "top mlp input %d != T*D + m_den = %d",
             e->ln_top[0], T * D + e->m_den);                    // wide_and_deep.py:345
      e->ld_R = e->ln_top[0];
      TRYB(alloc_mlp_buffers(e, e->top, e->ln_top, 1, kActSigmoid));
      break;
    }
    case DRS_MODEL_NCF: {
      CHECKB(T == 4, "NCF only has 4 embedding tables");                 // ncf.py:352
      CHECKB(e->ln_top.size() >= 3 && e->ln_top[0] == 2 * D, "NCF top mlp input must be 2*D");
      e->m_den = 0;
      std::vector<int32_t> tower(e->ln_top.begin(), e->ln_top.end() - 1);   // ncf.py:332
      std::vector<int32_t> fin = {D + tower.back(), e->ln_top.back()};      // :339-343
      TRYB(alloc_mlp_buffers(e, e->top, tower, 1, kActRelu));
      TRYB(alloc_mlp_buffers(e, e->fin, fin, 1, kActRelu));
      TRYB(dev_alloc_t(e, &e->P, (size_t)maxB * T * D));
      TRYB(dev_alloc_t(e, &e->F, (size_t)maxB * (D + tower.back())));
      e->ld_R = 0;
      break;
    }
    case DRS_MODEL_DIN: {
      CHECKB(T >= 4, "DIN needs at least 4 embedding tables");           // din.py:356
      CHECKB(e->ln_top[0] == 4 * D, "DIN top mlp input must be 4*D");
      CHECKB(e->ln_bot.size() >= 2 && e->ln_bot.front() == 3 * D && e->ln_bot.back() == D,
             "DIN attention MLP must be [3D, ..., D]");                  // din.py:253-258
      e->m_den = 0;
      e->n_units = T - 3;
      e->ld_R = 4 * D;
      TRYB(alloc_mlp_buffers(e, e->bot, e->ln_bot, e->n_units, kActRelu));
      TRYB(alloc_mlp_buffers(e, e->top, e->ln_top, 1, kActRelu));
      TRYB(dev_alloc_t(e, &e->P, (size_t)maxB * T * D));
      TRYB(dev_alloc_t(e, &e->C, (size_t)e->n_units * maxB * 3 * D));
      break;
    }
    default:
      return bail(fail(DRS_ERR_INVALID, "unknown model_type %d", spec->model_type));
  }
  if (e->ld_R > 0) TRYB(dev_alloc_t(e, &e->R, (size_t)maxB * e->ld_R));
  Mlp& lastm = (spec->model_type == DRS_MODEL_NCF) ? e->fin : e->top;
  e->out = lastm.layers.back().y;
  e->out_dim = lastm.layers.back().out;

  e->tables.assign(T, nullptr);
  e->table_loaded.assign(T, 0);

  const int64_t max_lk = spec->max_lookups_per_table > 0 ? spec->max_lookups_per_table : maxB;
  const int64_t avg_len = ceil_div64(max_lk, maxB);
  plan_gather(T, D, avg_len, dot, &e->lpr, &e->SB, &e->idx_cap, &e->sls_smem);

  std::vector<int64_t> caps(T, max_lk);
  TRYB(alloc_input_set(e, e->slot, maxB, caps, true));
  TRYB(write_descs(e, e->slot, {}));
  CUB(cudaStreamSynchronize(e->stream));
  *out = e;
  return DRS_OK;
#undef TRYB
#undef CHECKB
#undef CUB
}

int drs_destroy(drs_engine* e) {
  if (!e) return DRS_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (float* t : e->tables) cudaFree(t);
  for (Mlp* m : {&e->bot, &e->top, &e->fin})
    for (Layer& l : m->layers) { cudaFree(l.W); cudaFree(l.b); cudaFree(l.Wsplit); cudaFree(l.y); }
  cudaFree(e->R); cudaFree(e->P); cudaFree(e->dense_out); cudaFree(e->F); cudaFree(e->C);
  for (auto& kv : e->staged) free_input_set(kv.second);
  free_input_set(e->slot);
  cudaFree(e->d_err);
  if (e->h_err) cudaFreeHost(e->h_err);
  for (auto& v : e->ev) if (v) cudaEventDestroy(v);
  for (auto& v : e->ev_timer) if (v) cudaEventDestroy(v);
  for (auto& v : e->prof_ev) cudaEventDestroy(v);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return DRS_OK;
}

int drs_out_dim(const drs_engine* e, int* out_dim) {
  CHECK_ARG(e && out_dim, "NULL argument");
  *out_dim = e->out_dim;
  return DRS_OK;
}

int drs_load_table(drs_engine* e, int table, const float* rows, int64_t n_rows, int dim) {
  CHECK_ARG(e && rows, "NULL argument");
  CHECK_ARG(table >= 0 && table < e->T, "table %d out of range [0,%d)", table, e->T);
  CHECK_ARG(n_rows == e->table_rows[table] && dim == e->D,
            "table %d: got [%lld,%d], spec says [%lld,%d]", table, (long long)n_rows, dim,
            (long long)e->table_rows[table], e->D);
  CU(cudaSetDevice(e->device));
  if (!e->tables[table]) {
    // +16 floats of slack keeps 16-B vector loads of the last row in-bounds for any D
    TRY(dev_alloc_t(e, &e->tables[table], (size_t)n_rows * dim + 16));
  }
  CU(cudaMemcpyAsync(e->tables[table], rows, (size_t)n_rows * dim * sizeof(float),
                     cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->table_loaded[table] = 1;
  // descriptors hold the table pointers
  TRY(write_descs(e, e->slot, {}));
  CU(cudaStreamSynchronize(e->stream));
  return DRS_OK;
}

int drs_load_fc(drs_engine* e, int mlp_id, int layer, const float* W, const float* b, int out_dim,
                int in_dim) {
  CHECK_ARG(e && W && b, "NULL argument");
  CU(cudaSetDevice(e->device));
  Mlp* m = nullptr;
  int unit = 0;
  if (mlp_id == DRS_MLP_BOT && e->spec.model_type == DRS_MODEL_DLRM) m = &e->bot;
  else if (mlp_id == DRS_MLP_TOP) m = &e->top;
  else if (mlp_id == DRS_MLP_FINAL && e->spec.model_type == DRS_MODEL_NCF) m = &e->fin;
  else if (mlp_id >= DRS_MLP_ATTN_BASE && e->spec.model_type == DRS_MODEL_DIN &&
           mlp_id - DRS_MLP_ATTN_BASE < e->n_units) {
    m = &e->bot;
    unit = mlp_id - DRS_MLP_ATTN_BASE;
  }
  CHECK_ARG(m != nullptr, "mlp_id %d does not exist in this model", mlp_id);
  CHECK_ARG(layer >= 0 && layer < (int)m->layers.size(), "layer %d out of range [0,%d)", layer,
            (int)m->layers.size());
  Layer& l = m->layers[layer];
  CHECK_ARG(l.out == out_dim && l.in == in_dim, "mlp %d layer %d: got [%d,%d], spec says [%d,%d]",
            mlp_id, layer, out_dim, in_dim, l.out, l.in);
  const size_t wcount = (size_t)l.out * l.in;
  CU(cudaMemcpyAsync(l.W + unit * wcount, W, wcount * sizeof(float), cudaMemcpyHostToDevice,
                     e->stream));
  CU(cudaMemcpyAsync(l.b + (size_t)unit * l.out, b, (size_t)l.out * sizeof(float),
                     cudaMemcpyHostToDevice, e->stream));
  if (l.Wsplit) {
    int rc = tc_split_weights(l.W + unit * wcount, l.Wsplit + unit * 2 * wcount, (int64_t)wcount,
                              e->stream);
    if (rc != 0) return fail(DRS_ERR_CUDA, "weight split failed: %s",
                             cudaGetErrorString((cudaError_t)rc));
    e->launches += 1;
  }
  CU(cudaStreamSynchronize(e->stream));
  l.loaded[unit] = 1;
  return DRS_OK;
}

int drs_stage_batch(drs_engine* e, int batch_id, const float* dense, const void* const* idx,
                    int idx_is_i64, const int32_t* const* lengths, int n_samples) {
  CHECK_ARG(e, "engine is NULL");
  CHECK_ARG(n_samples >= 1, "n_samples must be >= 1");
  CHECK_ARG(idx && lengths, "idx/lengths pointer arrays must not be NULL");
  CU(cudaSetDevice(e->device));
  for (int t = 0; t < e->T; ++t)
    if (!e->table_loaded[t]) return fail(DRS_ERR_STATE, "stage after loading tables (table %d)", t);
  const int T = e->T;
  std::vector<int64_t> counts(T, 0);
  for (int t = 0; t < T; ++t)
    for (int b = 0; b < n_samples; ++b) {
      if (lengths[t][b] < 0) return fail(DRS_ERR_INDEX, "negative length");
      counts[t] += lengths[t][b];
    }
  auto it = e->staged.find(batch_id);
  if (it != e->staged.end()) { free_input_set(it->second); e->staged.erase(it); }
  InputSet s;
  int rc = alloc_input_set(e, s, n_samples, counts, idx_is_i64 != 0);
  if (rc != DRS_OK) { free_input_set(s); return rc; }
  std::vector<int64_t> n_idx;
  rc = upload_inputs(e, s, dense, idx, idx_is_i64, lengths, n_samples, &n_idx);
  if (rc == DRS_OK) rc = write_descs(e, s, n_idx);
  if (rc == DRS_OK) rc = launch_scan(e, s, n_samples);
  if (rc == DRS_OK) {
    cudaError_t ce = cudaMemcpyAsync(e->h_err, e->d_err, sizeof(int), cudaMemcpyDeviceToHost,
                                     e->stream);
    if (ce != cudaSuccess) rc = fail(DRS_ERR_CUDA, "%s", cudaGetErrorString(ce));
  }
  if (rc == DRS_OK) rc = finish_sync(e);
  if (rc != DRS_OK) { free_input_set(s); return rc; }
  if (s.idx64) { cudaFree(s.idx64); s.idx64 = nullptr; }
  s.valid = true;
  e->staged[batch_id] = s;
  return DRS_OK;
}

int drs_enqueue_staged(drs_engine* e, int batch_id, int batch_size) {
  CHECK_ARG(e, "engine is NULL");
  auto it = e->staged.find(batch_id);
  if (it == e->staged.end()) return fail(DRS_ERR_STATE, "batch %d was not staged", batch_id);
  CHECK_ARG(batch_size >= 1 && batch_size <= it->second.n && batch_size <= e->spec.max_batch,
            "batch_size %d out of range [1, min(%d staged, %d max_batch)]", batch_size,
            it->second.n, e->spec.max_batch);
  CU(cudaSetDevice(e->device));
  TRY(check_ready(e));
  return enqueue_graph(e, it->second, batch_size);
}

int drs_sync(drs_engine* e) {
  CHECK_ARG(e, "engine is NULL");
  CU(cudaSetDevice(e->device));
  return finish_sync(e);
}

int drs_infer_staged(drs_engine* e, int batch_id, int batch_size, float* out, drs_timing* tm) {
  CHECK_ARG(e, "engine is NULL");
  const auto t0 = std::chrono::steady_clock::now();
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev[1], e->stream));
  TRY(drs_enqueue_staged(e, batch_id, batch_size));
  CU(cudaEventRecord(e->ev[2], e->stream));
  if (out)
    CU(cudaMemcpyAsync(out, e->out, (size_t)batch_size * e->out_dim * sizeof(float),
                       cudaMemcpyDeviceToHost, e->stream));
  CU(cudaEventRecord(e->ev[3], e->stream));
  TRY(finish_sync(e));
  if (tm) {
    tm->h2d_ms = 0.f;
    CU(cudaEventElapsedTime(&tm->gpu_ms, e->ev[1], e->ev[2]));
    CU(cudaEventElapsedTime(&tm->d2h_ms, e->ev[2], e->ev[3]));
    tm->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0)
                       .count();
  }
  return DRS_OK;
}

int drs_enqueue_host(drs_engine* e, const float* dense, const void* const* idx, int idx_is_i64,
                     const int32_t* const* lengths, int batch_size, float* out_pinned) {
  CHECK_ARG(e, "engine is NULL");
  CHECK_ARG(batch_size >= 1 && batch_size <= e->spec.max_batch,
            "batch_size %d out of range [1,%d]", batch_size, e->spec.max_batch);
  CU(cudaSetDevice(e->device));
  TRY(check_ready(e));
  std::vector<int64_t> n_idx;
  CU(cudaEventRecord(e->ev[0], e->stream));
  TRY(upload_inputs(e, e->slot, dense, idx, idx_is_i64, lengths, batch_size, &n_idx));
  CU(cudaEventRecord(e->ev[1], e->stream));
  TRY(launch_scan(e, e->slot, batch_size));
  TRY(enqueue_graph(e, e->slot, batch_size));
  CU(cudaEventRecord(e->ev[2], e->stream));
  if (out_pinned)
    CU(cudaMemcpyAsync(out_pinned, e->out, (size_t)batch_size * e->out_dim * sizeof(float),
                       cudaMemcpyDeviceToHost, e->stream));
  CU(cudaEventRecord(e->ev[3], e->stream));
  return DRS_OK;
}

int drs_infer(drs_engine* e, const float* dense, const void* const* idx, int idx_is_i64,
              const int32_t* const* lengths, int batch_size, float* out, drs_timing* tm) {
  const auto t0 = std::chrono::steady_clock::now();
  TRY(drs_enqueue_host(e, dense, idx, idx_is_i64, lengths, batch_size, out));
  TRY(finish_sync(e));
  if (tm) {
    CU(cudaEventElapsedTime(&tm->h2d_ms, e->ev[0], e->ev[1]));
    CU(cudaEventElapsedTime(&tm->gpu_ms, e->ev[1], e->ev[2]));
    CU(cudaEventElapsedTime(&tm->d2h_ms, e->ev[2], e->ev[3]));
    tm->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0)
                       .count();
  }
  return DRS_OK;
}

int drs_timer_start(drs_engine* e) {
  CHECK_ARG(e, "engine is NULL");
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev_timer[0], e->stream));
  return DRS_OK;
}

int drs_timer_stop(drs_engine* e, float* elapsed_ms) {
  CHECK_ARG(e && elapsed_ms, "NULL argument");
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev_timer[1], e->stream));
  CU(cudaEventSynchronize(e->ev_timer[1]));
  CU(cudaEventElapsedTime(elapsed_ms, e->ev_timer[0], e->ev_timer[1]));
  return DRS_OK;
}

int drs_sparse_lengths_sum(drs_engine* e, int table, const void* idx, int idx_is_i64,
                           int64_t n_idx, const int32_t* lengths, int n_segments, float* out) {
  CHECK_ARG(e && lengths && out && (idx || n_idx == 0), "NULL argument");
  CHECK_ARG(table >= 0 && table < e->T, "table %d out of range", table);
  CHECK_ARG(n_segments >= 0 && n_idx >= 0, "negative sizes");
  if (!e->table_loaded[table]) return fail(DRS_ERR_STATE, "table %d not loaded", table);
  if (n_segments == 0) return DRS_OK;
  CU(cudaSetDevice(e->device));
  const int D = e->D;
  int32_t *d_idx = nullptr, *d_len = nullptr, *d_off = nullptr;
  int64_t* d_idx64 = nullptr;
  float* d_out = nullptr;
  SlsTableDesc* d_desc = nullptr;
  ScanTableDesc* d_scan = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_idx); cudaFree(d_len); cudaFree(d_off); cudaFree(d_idx64); cudaFree(d_out);
    cudaFree(d_desc); cudaFree(d_scan);
  };
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(DRS_ERR_CUDA, "%s -> %s", #call, cudaGetErrorString(e_)); } } while (0)
  CUC(cudaMalloc(&d_idx, ((size_t)n_idx + 8) * sizeof(int32_t)));
  CUC(cudaMemsetAsync(d_idx, 0, ((size_t)n_idx + 8) * sizeof(int32_t), e->stream));
  CUC(cudaMalloc(&d_len, (size_t)n_segments * sizeof(int32_t)));
  CUC(cudaMalloc(&d_off, ((size_t)n_segments + 1) * sizeof(int32_t)));
  CUC(cudaMalloc(&d_out, (size_t)n_segments * D * sizeof(float)));
  CUC(cudaMalloc(&d_desc, sizeof(SlsTableDesc)));
  CUC(cudaMalloc(&d_scan, sizeof(ScanTableDesc)));
  if (n_idx > 0) {
    if (idx_is_i64) {
      CUC(cudaMalloc(&d_idx64, (size_t)n_idx * sizeof(int64_t)));
      CUC(cudaMemcpyAsync(d_idx64, idx, (size_t)n_idx * sizeof(int64_t), cudaMemcpyHostToDevice,
                          e->stream));
      k_cast_i64_i32<<<ew_grid(n_idx), 256, 0, e->stream>>>(d_idx64, d_idx, n_idx);
      e->launches += 1;
    } else {
      CUC(cudaMemcpyAsync(d_idx, idx, (size_t)n_idx * sizeof(int32_t), cudaMemcpyHostToDevice,
                          e->stream));
    }
  }
  CUC(cudaMemcpyAsync(d_len, lengths, (size_t)n_segments * sizeof(int32_t),
                      cudaMemcpyHostToDevice, e->stream));
  SlsTableDesc hd{};
  hd.rows = e->tables[table]; hd.idx = d_idx; hd.off = d_off; hd.n_rows = e->table_rows[table];
  hd.dst_col = 0;
  ScanTableDesc hs{};
  hs.lengths = d_len; hs.off = d_off; hs.n_idx = n_idx;
  CUC(cudaMemcpyAsync(d_desc, &hd, sizeof(hd), cudaMemcpyHostToDevice, e->stream));
  CUC(cudaMemcpyAsync(d_scan, &hs, sizeof(hs), cudaMemcpyHostToDevice, e->stream));
  k_lengths_to_offsets<<<1, kScanThreads, 0, e->stream>>>(d_scan, n_segments, e->d_err);
  e->launches += 1;
  {
    int lpr, SB, cap; size_t smem;
    const int64_t avg = n_segments > 0 ? ceil_div64(n_idx, n_segments) : 0;
    plan_gather(1, D, avg, false, &lpr, &SB, &cap, &smem);
    SlsLaunch p{};
    p.tables = d_desc; p.T = 1; p.B = n_segments; p.D = D;
    p.out = d_out; p.ld_out = D; p.SB = SB; p.idx_cap = cap;
    p.vec_store = (D % 4 == 0) ? 1 : 0;
    p.err_flag = e->d_err;
    int rc = launch_gather_lpr<false>(e, p, lpr, ceil_div(n_segments, SB), smem);
    if (rc != DRS_OK) { cleanup(); return rc; }
  }
  CUC(cudaMemcpyAsync(out, d_out, (size_t)n_segments * D * sizeof(float), cudaMemcpyDeviceToHost,
                      e->stream));
  CUC(cudaMemcpyAsync(e->h_err, e->d_err, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  int rc = finish_sync(e);
  cleanup();
  return rc;
#undef CUC
}

int drs_fetch_stage(drs_engine* e, int stage, int layer, float* dst, int64_t cap, int* rows,
                    int* cols) {
  CHECK_ARG(e && dst && rows && cols, "NULL argument");
  CHECK_ARG(e->last_inputs && e->last_B > 0, "no query has been served yet");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  const int B = e->last_B, D = e->D, T = e->T;
  const float* src = nullptr;
  int64_t ld = 0;
  int r = B, c = 0;
  switch (stage) {
    case DRS_STAGE_POOLED:
      c = T * D;
      if (e->P) { src = e->P; ld = (int64_t)T * D; }
      else { src = e->R + gather_dst_col(e, 0); ld = e->ld_R; }
      break;
    case DRS_STAGE_BOT: {
      if (e->spec.model_type == DRS_MODEL_DIN) { src = e->R + D; ld = e->ld_R; c = D; break; }
      CHECK_ARG(e->spec.model_type == DRS_MODEL_DLRM, "model has no bottom MLP");
      CHECK_ARG(layer >= 0 && layer < (int)e->bot.layers.size(), "bottom layer out of range");
      const Layer& l = e->bot.layers[layer];
      c = l.out;
      if (layer == (int)e->bot.layers.size() - 1) {
        if (e->dense_out) { src = e->dense_out; ld = D; } else { src = e->R; ld = e->ld_R; }
      } else { src = l.y; ld = l.out; }
      break;
    }
    case DRS_STAGE_INTERACT:
      if (e->spec.model_type == DRS_MODEL_NCF) {
        src = e->F; c = D + e->top.layers.back().out; ld = c;
      } else { src = e->R; c = e->ld_R; ld = e->ld_R; }
      break;
    case DRS_STAGE_TOP: {
      CHECK_ARG(layer >= 0 && layer < (int)e->top.layers.size(), "top layer out of range");
      const Layer& l = e->top.layers[layer];
      c = l.out;
      if (e->spec.model_type == DRS_MODEL_NCF && layer == (int)e->top.layers.size() - 1) {
        src = e->F + D; ld = D + l.out;
      } else { src = l.y; ld = l.out; }
      break;
    }
    case DRS_STAGE_OUT:
      src = e->out; c = e->out_dim; ld = c;
      break;
    case DRS_STAGE_OFFSETS: {
      r = T; c = B + 1;
      CHECK_ARG((int64_t)r * c <= cap, "capacity too small: need %lld floats", (long long)r * c);
      CU(cudaMemcpy2D(dst, (size_t)c * 4, e->last_inputs->off, (size_t)(e->last_inputs->n + 1) * 4,
                      (size_t)c * 4, r, cudaMemcpyDeviceToHost));
      *rows = r; *cols = c;
      return DRS_OK;
    }
    default:
      return fail(DRS_ERR_INVALID, "unknown stage %d", stage);
  }
  CHECK_ARG((int64_t)r * c <= cap, "capacity too small: need %lld floats", (long long)r * c);
  CU(cudaMemcpy2D(dst, (size_t)c * 4, src, (size_t)ld * 4, (size_t)c * 4, r,
                  cudaMemcpyDeviceToHost));
  *rows = r; *cols = c;
  return DRS_OK;
}

int drs_launch_count(const drs_engine* e, int64_t* n) {
  CHECK_ARG(e && n, "NULL argument");
  *n = e->launches;
  return DRS_OK;
}

int drs_hbm_bytes(const drs_engine* e, int64_t* bytes) {
  CHECK_ARG(e && bytes, "NULL argument");
  *bytes = e->hbm_bytes;
  return DRS_OK;
}

int drs_profile_begin(drs_engine* e, int max_steps) {
  CHECK_ARG(e && max_steps >= 1, "bad arguments");
  CU(cudaSetDevice(e->device));
  while ((int)e->prof_ev.size() < max_steps * 4) {
    cudaEvent_t ev;
    CU(cudaEventCreate(&ev));
    e->prof_ev.push_back(ev);
  }
  e->prof_cap = max_steps;
  e->prof_n = 0;
  e->prof_on = true;
  return DRS_OK;
}

int drs_profile_end(drs_engine* e, int* n_steps, float* pre_ms, float* gather_ms, float* post_ms) {
  CHECK_ARG(e && n_steps && pre_ms && gather_ms && post_ms, "NULL argument");
  CU(cudaSetDevice(e->device));
  e->prof_on = false;
  CU(cudaStreamSynchronize(e->stream));
  double a = 0, b = 0, c = 0;
  for (int i = 0; i < e->prof_n; ++i) {
    float t;
    CU(cudaEventElapsedTime(&t, e->prof_ev[(size_t)i * 4 + 0], e->prof_ev[(size_t)i * 4 + 1])); a += t;
    CU(cudaEventElapsedTime(&t, e->prof_ev[(size_t)i * 4 + 1], e->prof_ev[(size_t)i * 4 + 2])); b += t;
    CU(cudaEventElapsedTime(&t, e->prof_ev[(size_t)i * 4 + 2], e->prof_ev[(size_t)i * 4 + 3])); c += t;
  }
  *n_steps = e->prof_n;
  *pre_ms = (float)a; *gather_ms = (float)b; *post_ms = (float)c;
  return DRS_OK;
}

int drs_shard_export(drs_engine* e, void* blob_out) {
  CHECK_ARG(e && blob_out, "NULL argument");
  return fail(DRS_ERR_COMM, "table-sharded exchange is not built in this version");
}

int drs_shard_connect(drs_engine* e, int rank, int world, const void* all_blobs) {
  CHECK_ARG(e && all_blobs, "NULL argument");
  (void)rank; (void)world;
  return fail(DRS_ERR_COMM, "table-sharded exchange is not built in this version");
}

}  // extern "C"
