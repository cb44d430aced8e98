// Queue wire codec of the serving loop (host only, compiled by g++): the reference's serving
// protocol is pickled packets over multiprocessing Queues — a frame is struct.pack("!i", len) +
// pickle.dumps(obj) (/root/reference/DeepRecSys.py:44-66, loadGenerator.py:162-177,
// accelInferenceEngine.py:46-83, utils/packets.py:6-59).  A B200 answers a 256-sample RMC1 query
// in ~4 us; unpickling the ServiceRequest and building + pickling the ServiceResponse in Python
// cost more than that per packet.  These two functions move both ends below the interpreter:
//
//   drs_wire_parse_requests   a buffer of frames read from the request pipe -> fixed-size
//                             records, as long as each frame is exactly a pickled
//                             `ServiceRequest` (of a module the caller names) whose fields are
//                             int / bool / None (arrival_time: float / None); the first frame
//                             that is anything else (the final
//                             None, a request that carries tensors, an unexpected type) stops
//                             the parse and is left to pickle.loads;
//   drs_wire_emit_responses   records -> frames holding pickled `ServiceResponse` objects
//                             (protocol 4, the same opcode sequence pickle.dumps emits), ready
//                             for ONE write to the response pipe.
//
// A stock Queue.get() on the other side unpickles them into the reference's own classes:
// attribute names, types and values are what the Python path produces (checked against
// pickle.dumps / pickle.loads in tests/test_wire.py and by a self-check at engine start).
#include "../../include/drs_b200.h"

#include <cstdint>
#include <cstring>

namespace {

enum : uint8_t {
  OP_PROTO = 0x80, OP_FRAME = 0x95, OP_SHORT_BINUNICODE = 0x8c, OP_MEMOIZE = 0x94,
  OP_STACK_GLOBAL = 0x93, OP_EMPTY_TUPLE = ')', OP_NEWOBJ = 0x81, OP_EMPTY_DICT = '}',
  OP_MARK = '(', OP_BININT1 = 'K', OP_BININT2 = 'M', OP_BININT = 'J', OP_LONG1 = 0x8a,
  OP_BINFLOAT = 'G', OP_NEWTRUE = 0x88, OP_NEWFALSE = 0x89, OP_NONE = 'N', OP_SETITEMS = 'u',
  OP_BUILD = 'b', OP_STOP = '.'
};

constexpr int KIND_INT = 0, KIND_BOOL = 1, KIND_NONE = 2;      // utils/response_ring.py

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool take(uint8_t op) { if (p < end && *p == op) { ++p; return true; } return false; }
  bool have(int64_t n) const { return end - p >= n; }
};

// SHORT_BINUNICODE <len> <bytes> MEMOIZE -> (ptr, len)
bool short_str(Cursor& c, const uint8_t** s, int* n) {
  if (!c.take(OP_SHORT_BINUNICODE) || !c.have(1)) return false;
  *n = *c.p++;
  if (!c.have(*n)) return false;
  *s = c.p;
  c.p += *n;
  return c.take(OP_MEMOIZE);
}

bool is(const uint8_t* s, int n, const char* lit) {
  return (int)std::strlen(lit) == n && std::memcmp(s, lit, (size_t)n) == 0;
}

// request fields in the record's kinds order (utils/response_ring._INT_FIELDS + arrival_time)
const char* const kReqFields[7] = {"epoch", "batch_id", "batch_size", "sub_id",
                                   "total_sub_batches", "exp_packet", "arrival_time"};

// one value: int (BININT1/2, BININT, LONG1 <= 8 bytes), bool, None, float
struct Value { int kind; int64_t i; double f; bool is_float; };

bool value(Cursor& c, Value* v) {
  if (!c.have(1)) return false;
  const uint8_t op = *c.p++;
  v->is_float = false; v->kind = KIND_INT; v->i = 0; v->f = 0.0;
  switch (op) {
    case OP_BININT1: if (!c.have(1)) return false; v->i = c.p[0]; c.p += 1; return true;
    case OP_BININT2: if (!c.have(2)) return false; v->i = c.p[0] | (c.p[1] << 8); c.p += 2; return true;
    case OP_BININT: {
      if (!c.have(4)) return false;
      uint32_t u; std::memcpy(&u, c.p, 4); c.p += 4;           // little-endian host
      v->i = (int32_t)u; return true;
    }
    case OP_LONG1: {
      if (!c.have(1)) return false;
      const int n = *c.p++;
      if (n > 8 || !c.have(n)) return false;
      uint64_t u = 0;
      for (int k = 0; k < n; ++k) u |= (uint64_t)c.p[k] << (8 * k);
      if (n > 0 && n < 8 && (c.p[n - 1] & 0x80)) u |= ~uint64_t(0) << (8 * n);   // sign-extend
      c.p += n; v->i = (int64_t)u; return true;
    }
    case OP_NEWTRUE: v->kind = KIND_BOOL; v->i = 1; return true;
    case OP_NEWFALSE: v->kind = KIND_BOOL; v->i = 0; return true;
    case OP_NONE: v->kind = KIND_NONE; return true;
    case OP_BINFLOAT: {
      if (!c.have(8)) return false;
      uint64_t u = 0;
      for (int k = 0; k < 8; ++k) u = (u << 8) | c.p[k];       // big-endian on the wire
      c.p += 8; std::memcpy(&v->f, &u, 8); v->is_float = true; return true;
    }
    default: return false;
  }
}

// is (s, len) one of the NUL-separated names of `list` (terminated by an empty name)?
bool listed(const char* list, const uint8_t* s, int len) {
  for (const char* m = list; *m; m += std::strlen(m) + 1)
    if (is(s, len, m)) return true;
  return false;
}

// payload = pickle.dumps(ServiceRequest(...)) exactly?  -> record
bool parse_request(const uint8_t* p, int64_t n, const char* modules, drs_wire_rec* r) {
  Cursor c{p, p + n};
  if (!c.take(OP_PROTO) || !c.have(1) || (*c.p != 4 && *c.p != 5)) return false;
  ++c.p;
  if (c.take(OP_FRAME)) {                                      // must frame exactly the rest
    if (!c.have(8)) return false;
    uint64_t fl;
    std::memcpy(&fl, c.p, 8);
    c.p += 8;
    if (fl != (uint64_t)(c.end - c.p)) return false;
  }
  const uint8_t* s; int len;
  if (!short_str(c, &s, &len)) return false;                   // module: one the caller vouches for
  if (modules && !listed(modules, s, len)) return false;
  if (!short_str(c, &s, &len) || !is(s, len, "ServiceRequest")) return false;
  if (!c.take(OP_STACK_GLOBAL) || !c.take(OP_MEMOIZE) || !c.take(OP_EMPTY_TUPLE) ||
      !c.take(OP_NEWOBJ) || !c.take(OP_MEMOIZE) || !c.take(OP_EMPTY_DICT) ||
      !c.take(OP_MEMOIZE) || !c.take(OP_MARK))
    return false;
  int64_t* const ints[6] = {&r->epoch, &r->batch_id, &r->batch_size, &r->sub_id,
                            &r->total_sub_batches, &r->exp_packet};
  unsigned seen = 0;
  int64_t kinds = 0;
  for (int k = 0; k < 7; ++k) {
    if (!short_str(c, &s, &len)) return false;
    int f = -1;
    for (int j = 0; j < 7; ++j) if (is(s, len, kReqFields[j])) { f = j; break; }
    if (f < 0 || (seen >> f) & 1) return false;
    seen |= 1u << f;
    Value v;
    if (!value(c, &v)) return false;
    if (f < 6) {
      if (v.is_float) return false;
      *ints[f] = v.i;
    } else {
      if (v.kind == KIND_NONE) r->arrival_time = 0.0;
      else if (v.is_float) r->arrival_time = v.f;
      else return false;                                        // an int arrival_time: Python path
    }
    kinds |= (int64_t)v.kind << (2 * f);
  }
  if (seen != 0x7f) return false;
  if (!c.take(OP_SETITEMS) || !c.take(OP_BUILD) || !c.take(OP_STOP) || c.p != c.end) return false;
  r->kinds = kinds;
  r->out_batch_size = 0;
  r->start_time = r->end_time = 0.0;
  return true;
}

// ---- emitter -------------------------------------------------------------------------------
struct Out {
  uint8_t* p;
  uint8_t* end;
  bool ok = true;
  void put(uint8_t b) { if (p < end) *p++ = b; else ok = false; }
  void put(const void* s, int64_t n) {
    if (end - p >= n) { std::memcpy(p, s, (size_t)n); p += n; } else ok = false;
  }
};

void emit_int(Out& o, int64_t v) {               // pickle.save_long, protocol >= 2
  if (v >= 0 && v <= 0xff) { o.put(OP_BININT1); o.put((uint8_t)v); }
  else if (v >= 0 && v <= 0xffff) { o.put(OP_BININT2); o.put((uint8_t)v); o.put((uint8_t)(v >> 8)); }
  else if (v >= INT32_MIN && v <= INT32_MAX) {
    const uint32_t u = (uint32_t)(int32_t)v;
    o.put(OP_BININT); o.put(&u, 4);
  } else {                                       // LONG1: minimal two's-complement little-endian
    uint8_t b[8];
    std::memcpy(b, &v, 8);
    int n = 8;
    while (n > 1 && ((b[n - 1] == 0x00 && !(b[n - 2] & 0x80)) ||
                     (b[n - 1] == 0xff && (b[n - 2] & 0x80))))
      --n;
    o.put(OP_LONG1); o.put((uint8_t)n); o.put(b, n);
  }
}

void emit_float(Out& o, double f) {
  uint64_t u;
  std::memcpy(&u, &f, 8);
  o.put(OP_BINFLOAT);
  for (int k = 7; k >= 0; --k) o.put((uint8_t)(u >> (8 * k)));
}

void emit_kinded(Out& o, int kind, int64_t v) {
  if (kind == KIND_NONE) o.put(OP_NONE);
  else if (kind == KIND_BOOL) o.put(v ? OP_NEWTRUE : OP_NEWFALSE);
  else emit_int(o, v);
}

// ServiceResponse attributes (utils/packets.py:32-59; the process_start_time keyword is stored
// as queue_start_time — the reference's own quirk)
const char* const kRespFields[12] = {
    "consumer_id", "epoch", "batch_id", "batch_size", "arrival_time", "queue_start_time",
    "queue_end_time", "inference_end_time", "out_batch_size", "total_sub_batches", "exp_packet",
    "sub_id"};

}  // namespace

extern "C" {

int drs_wire_parse_requests(const uint8_t* buf, int64_t len, const char* modules,
                            drs_wire_rec* recs, int max_recs, int32_t* n_recs, int64_t* consumed,
                            int32_t* stop) {
  if (!buf || !recs || !n_recs || !consumed || !stop || len < 0 || max_recs < 0)
    return DRS_ERR_INVALID;
  int64_t pos = 0;
  int n = 0;
  *stop = DRS_WIRE_NEED_MORE;
  while (n < max_recs) {
    if (len - pos < 4) break;
    const int32_t sz = (int32_t)(((uint32_t)buf[pos] << 24) | ((uint32_t)buf[pos + 1] << 16) |
                                 ((uint32_t)buf[pos + 2] << 8) | (uint32_t)buf[pos + 3]);
    if (sz < 0) { *stop = DRS_WIRE_PYTHON; break; }            // > 2 GiB framing: not ours
    if (len - pos - 4 < sz) break;
    if (!parse_request(buf + pos + 4, sz, modules, &recs[n])) { *stop = DRS_WIRE_PYTHON; break; }
    pos += 4 + (int64_t)sz;
    ++n;
  }
  if (n == max_recs && *stop == DRS_WIRE_NEED_MORE && len - pos >= 4) *stop = DRS_WIRE_FULL;
  *n_recs = n;
  *consumed = pos;
  return DRS_OK;
}

int drs_wire_emit_responses(const drs_wire_rec* recs, int n, const uint8_t* class_ops,
                            int class_ops_len, const int32_t* field_order, int64_t consumer_id,
                            uint8_t* out, int64_t cap, int64_t* out_len) {
  if (!recs || !class_ops || !field_order || !out || !out_len || n < 0 || class_ops_len <= 0)
    return DRS_ERR_INVALID;
  for (int k = 0; k < 12; ++k)
    if (field_order[k] < 0 || field_order[k] >= 12) return DRS_ERR_INVALID;
  Out o{out, out + cap};
  for (int i = 0; i < n; ++i) {
    const drs_wire_rec& r = recs[i];
    uint8_t* const frame = o.p;
    o.put("\0\0\0\0", 4);                          // "!i" length, patched below
    o.put(OP_PROTO); o.put(4);
    o.put(OP_FRAME);
    uint8_t* const flen = o.p;
    o.put("\0\0\0\0\0\0\0\0", 8);                  // little-endian frame length, patched below
    uint8_t* const body = o.p;
    o.put(class_ops, class_ops_len);               // module, qualname, STACK_GLOBAL (+ memoize)
    o.put(OP_EMPTY_TUPLE); o.put(OP_NEWOBJ); o.put(OP_MEMOIZE);
    o.put(OP_EMPTY_DICT); o.put(OP_MEMOIZE); o.put(OP_MARK);
    auto kind = [&](int f) { return (int)((r.kinds >> (2 * f)) & 3); };
    for (int k = 0; k < 12; ++k) {
      const int f = field_order[k];
      const char* name = kRespFields[f];
      const int nl = (int)std::strlen(name);
      o.put(OP_SHORT_BINUNICODE); o.put((uint8_t)nl); o.put(name, nl); o.put(OP_MEMOIZE);
      switch (f) {
        case 0: emit_int(o, consumer_id); break;
        case 1: emit_kinded(o, kind(0), r.epoch); break;
        case 2: emit_kinded(o, kind(1), r.batch_id); break;
        case 3: emit_kinded(o, kind(2), r.batch_size); break;
        case 4: if (kind(6) == KIND_NONE) o.put(OP_NONE); else emit_float(o, r.arrival_time); break;
        case 5: emit_float(o, r.start_time); break;
        case 6: emit_float(o, r.end_time); break;
        case 7: emit_float(o, r.end_time); break;
        case 8: emit_int(o, r.out_batch_size); break;
        case 9: emit_kinded(o, kind(4), r.total_sub_batches); break;
        case 10: emit_kinded(o, kind(5), r.exp_packet); break;
        case 11: emit_kinded(o, kind(3), r.sub_id); break;
      }
    }
    o.put(OP_SETITEMS); o.put(OP_BUILD); o.put(OP_STOP);
    if (!o.ok) return DRS_ERR_INVALID;             // caller's buffer too small
    const uint64_t fl = (uint64_t)(o.p - body);
    std::memcpy(flen, &fl, 8);
    const uint32_t total = (uint32_t)(o.p - frame - 4);
    frame[0] = (uint8_t)(total >> 24); frame[1] = (uint8_t)(total >> 16);
    frame[2] = (uint8_t)(total >> 8); frame[3] = (uint8_t)total;
  }
  *out_len = o.p - out;
  return DRS_OK;
}

}  // extern "C"
