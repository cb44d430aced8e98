#include "../../deeprecsys_b200/csrc/hostpack.h"
#include <vector>
#include <cstdio>
#include <cstdlib>
using namespace drs;
static int calls = 0;
static void done(int j, void* ctx) { ++calls; (void)j; (void)ctx; }
int main() {
  HostPacker* p = hostpack_create(6);
  std::vector<int32_t> ids(64 * 512 * 9 + 77);
  for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int32_t)(i * 2654435761u % 4000000u);
  std::vector<int64_t> ids64(ids.begin(), ids.end());
  std::vector<uint32_t> out(pack_block_words(ids.size(), 22) * 4 + 64);
  for (int rep = 0; rep < 200; ++rep) {
    PackJob jobs[4];
    for (int t = 0; t < 4; ++t) {
      jobs[t].ids = (t & 1) ? (const void*)ids64.data() : (const void*)ids.data();
      jobs[t].is_i64 = t & 1; jobs[t].n = (int64_t)ids.size() - t * 1000; jobs[t].bits = 22;
      jobs[t].out = out.data() + t * pack_block_words(ids.size(), 22);
    }
    if (rep == 7) ids[5] = -3;
    hostpack_run(p, jobs, 4, done, nullptr);
    if (rep == 7) { if (jobs[0].ok || !jobs[1].ok) { printf("ok flags wrong\n"); return 1; } ids[5] = 5; }
    else for (int t = 0; t < 4; ++t) if (!jobs[t].ok) { printf("unexpected !ok\n"); return 1; }
  }
  hostpack_destroy(p);
  printf("calls %d\n", calls);
  return 0;
}
