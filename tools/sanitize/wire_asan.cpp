#include "../../include/drs_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
int main() {
  FILE* f = fopen("frames.bin", "rb"); std::vector<uint8_t> buf(1 << 20); size_t n = fread(buf.data(), 1, buf.size(), f); fclose(f);
  buf.resize(n);
  std::vector<drs_wire_rec> recs(256);
  const char mods[] = "utils.packets\0deeprecsys_b200.utils.packets\0";
  int32_t order[12] = {0,1,2,3,4,5,6,7,8,9,10,11};
  const uint8_t cls[] = "\x8c\x05m.pkg\x94\x8c\x0fServiceResponse\x94\x93\x94";
  srand(7);
  long accepted = 0, emitted = 0;
  for (int it = 0; it < 20000; ++it) {
    size_t len = it == 0 ? n : (size_t)(rand() % (int)n);
    uint8_t* h = (uint8_t*)malloc(len ? len : 1);            // exact-size heap copy: ASan guards both ends
    memcpy(h, buf.data() + (it == 0 ? 0 : rand() % (int)(n - len + 1)), len);
    if (it) for (int k = rand() % 4; k > 0 && len; --k) h[rand() % len] ^= (uint8_t)(1 << (rand() % 8));
    int32_t nr = 0, stop = 0; int64_t cons = 0;
    if (drs_wire_parse_requests(h, (int64_t)len, (it & 1) ? mods : nullptr, recs.data(), 256, &nr, &cons, &stop) != 0) return 2;
    if (cons < 0 || cons > (int64_t)len || nr < 0 || nr > 256) return 3;
    accepted += nr;
    for (int r = 0; r < nr; ++r) { recs[r].out_batch_size = recs[r].batch_size; recs[r].start_time = 1.0; recs[r].end_time = 2.0; }
    int64_t cap = (it % 3 == 0) ? rand() % 500 : (int64_t)nr * 384;
    uint8_t* o = (uint8_t*)malloc(cap ? cap : 1);
    int64_t olen = -1;
    int rc = drs_wire_emit_responses(recs.data(), nr, cls, (int)sizeof(cls) - 1, order, 7, o, cap, &olen);
    if (rc == 0) { if (olen < 0 || olen > cap) return 4; emitted += nr; }
    free(o); free(h);
  }
  printf("accepted %ld emitted %ld\n", accepted, emitted);
  return 0;
}
