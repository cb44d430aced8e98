// Host-side id packer (see hostpack.h for the layout and why it exists).  Plain C++ for g++:
// GCC vector extensions, one clone per ISA level picked at load time.
#include "hostpack.h"

#include <emmintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace drs {
namespace {

typedef uint32_t v16u __attribute__((vector_size(64), aligned(4)));
typedef uint64_t v16q __attribute__((vector_size(128), aligned(8)));

// nblk whole blocks of 512 ids -> nblk * 16 * bits words.  Returns the OR of every id (as
// unsigned 64-bit: a negative id or one that needs more than `bits` bits shows in the high part).
// NT: streaming stores — the packed words are read next by the copy engine, not by a core, and
// an ordinary store would first read the line it overwrites (measured on the B200 box, RMC1
// 16,384-sample steps on 8 threads: 18.3 M inferences/s with ordinary stores, 20.8 M with
// streaming ones); needs a 64-byte aligned destination
template <typename IdT, bool NT>
__attribute__((target_clones("avx512f", "avx2", "default")))
uint64_t pack_whole_blocks(const IdT* in, int64_t nblk, int bits, uint32_t* out) {
  v16u seen = {};
  v16q seen_q = {};
  for (int64_t b = 0; b < nblk; ++b) {
    const IdT* p = in + b * kPackBlockIds;
    uint32_t* o = out + b * kPackLanes * bits;
    v16u acc = {};
    int fill = 0;
    for (int j = 0; j < kPackBlockIds / kPackLanes; ++j) {
      v16u v;
      if (sizeof(IdT) == 8) {
        v16q q;
        std::memcpy(&q, p + j * kPackLanes, sizeof(q));
        seen_q |= q;
        v = __builtin_convertvector(q, v16u);
      } else {
        std::memcpy(&v, p + j * kPackLanes, sizeof(v));
        seen |= v;
      }
      acc |= v << fill;
      fill += bits;
      if (fill >= 32) {
        if (NT) {
          __m128i part[4];
          std::memcpy(part, &acc, sizeof(acc));
          for (int k = 0; k < 4; ++k) _mm_stream_si128(reinterpret_cast<__m128i*>(o) + k, part[k]);
        } else {
          std::memcpy(o, &acc, sizeof(acc));
        }
        o += kPackLanes;
        fill -= 32;
        acc = fill ? v >> (bits - fill) : (v16u){};
      }
    }
  }
  if (NT) _mm_sfence();
  uint64_t r = 0;
  for (int k = 0; k < kPackLanes; ++k) r |= (uint64_t)seen[k] | seen_q[k];
  return r;
}

bool use_streaming_stores() {
  static const bool on = [] {
    const char* v = std::getenv("DRS_HOSTPACK_NT");       // "0": ordinary stores (A/B switch)
    return v == nullptr || v[0] != '0';
  }();
  return on;
}

template <typename IdT>
uint64_t pack_range(const IdT* ids, int64_t n, int64_t b0, int64_t b1, int bits, uint32_t* out) {
  // blocks [b0, b1) of a job with n ids; the job's last block may be partial
  const int64_t whole_end = n / kPackBlockIds;
  uint64_t seen = 0;
  const int64_t w1 = b1 < whole_end ? b1 : whole_end;
  if (w1 > b0) {
    uint32_t* o = out + b0 * kPackLanes * bits;
    const bool nt = use_streaming_stores() && (reinterpret_cast<uintptr_t>(o) & 63) == 0;
    seen |= nt ? pack_whole_blocks<IdT, true>(ids + b0 * kPackBlockIds, w1 - b0, bits, o)
               : pack_whole_blocks<IdT, false>(ids + b0 * kPackBlockIds, w1 - b0, bits, o);
  }
  if (b1 > whole_end && whole_end >= b0) {
    IdT tail[kPackBlockIds];
    std::memset(tail, 0, sizeof(tail));
    std::memcpy(tail, ids + whole_end * kPackBlockIds,
                (size_t)(n - whole_end * kPackBlockIds) * sizeof(IdT));
    seen |= pack_whole_blocks<IdT, false>(tail, 1, bits, out + whole_end * kPackLanes * bits);
  }
  return seen;
}

struct Chunk { int job; int64_t b0, b1; };
constexpr int64_t kChunkBlocks = 64;          // 32 Ki ids: 128 KB (int32) in, <= 96 KB out

}  // namespace

class HostPacker {
 public:
  explicit HostPacker(int threads) {
    for (int i = 1; i < threads; ++i) workers_.emplace_back([this] { worker(); });
  }
  ~HostPacker() {
    { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int threads() const { return (int)workers_.size() + 1; }

  void run(PackJob* jobs, int n_jobs, void (*done)(int, void*), void* ctx) {
    chunks_.clear();
    remaining_.assign(n_jobs, 0);
    for (int j = 0; j < n_jobs; ++j) {
      jobs[j].ok = 1;
      const int64_t nb = pack_blocks_of(jobs[j].n);
      for (int64_t b = 0; b < nb; b += kChunkBlocks) {
        chunks_.push_back({j, b, b + kChunkBlocks < nb ? b + kChunkBlocks : nb});
        ++remaining_[j];
      }
    }
    left_.reset(new std::atomic<int>[n_jobs > 0 ? n_jobs : 1]);
    for (int j = 0; j < n_jobs; ++j) left_[j].store(remaining_[j], std::memory_order_relaxed);
    jobs_ = jobs;
    next_.store(0, std::memory_order_relaxed);
    // few chunks: the wake-up of the pool costs more than the work
    const bool use_pool = !workers_.empty() && chunks_.size() > 2;
    if (use_pool) {
      { std::lock_guard<std::mutex> g(mu_);
        busy_.store((int)workers_.size(), std::memory_order_relaxed);
        ++gen_; }
      cv_.notify_all();
    }
    int reported = 0;
    auto report = [&] {
      while (reported < n_jobs && left_[reported].load(std::memory_order_acquire) == 0) {
        if (done) done(reported, ctx);
        ++reported;
      }
    };
    while (work_one()) report();
    while (reported < n_jobs) {
      report();
      if (reported < n_jobs) std::this_thread::yield();
    }
    if (use_pool)
      while (busy_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  }

 private:
  bool work_one() {
    const int i = next_.fetch_add(1, std::memory_order_relaxed);
    if (i >= (int)chunks_.size()) return false;
    const Chunk& c = chunks_[i];
    PackJob& jb = jobs_[c.job];
    const uint64_t seen =
        jb.is_i64 ? pack_range<uint64_t>(static_cast<const uint64_t*>(jb.ids), jb.n, c.b0, c.b1,
                                         jb.bits, jb.out)
                  : pack_range<uint32_t>(static_cast<const uint32_t*>(jb.ids), jb.n, c.b0, c.b1,
                                         jb.bits, jb.out);
    if (seen >> jb.bits) __atomic_store_n(&jb.ok, 0, __ATOMIC_RELAXED);
    left_[c.job].fetch_sub(1, std::memory_order_acq_rel);
    return true;
  }
  void worker() {
    uint64_t seen_gen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return stop_ || gen_ != seen_gen; });
        if (stop_) return;
        seen_gen = gen_;
      }
      while (work_one()) {}
      busy_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }

  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t gen_ = 0;
  bool stop_ = false;
  std::vector<Chunk> chunks_;
  std::vector<int> remaining_;
  std::unique_ptr<std::atomic<int>[]> left_;
  PackJob* jobs_ = nullptr;
  std::atomic<int> next_{0};
  std::atomic<int> busy_{0};
};

HostPacker* hostpack_create(int threads) { return new HostPacker(threads < 1 ? 1 : threads); }
void hostpack_destroy(HostPacker* p) { delete p; }
int hostpack_threads(const HostPacker* p) { return p ? p->threads() : 0; }
void hostpack_run(HostPacker* p, PackJob* jobs, int n_jobs, void (*done)(int, void*), void* ctx) {
  p->run(jobs, n_jobs, done, ctx);
}

}  // namespace drs
