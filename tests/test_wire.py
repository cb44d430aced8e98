"""Queue wire codec (csrc/wire.cpp behind utils/wire.py): request frames are parsed and response
frames emitted in C — checked against ``pickle`` itself, which is what a stock
``multiprocessing.Queue`` on the other side runs (reference protocol: DeepRecSys.py:44-66,
loadGenerator.py:162-177, utils/packets.py:6-59)."""
import pickle
import struct

import numpy as np
from hypothesis import given, settings, strategies as st

from deeprecsys_b200.utils import wire
from deeprecsys_b200.utils.packets import ServiceRequest, ServiceResponse, TensorServiceRequest

HDR = struct.Struct("!i")
ints = st.one_of(st.integers(-2 ** 62, 2 ** 62), st.integers(-70000, 70000), st.integers(0, 300))
field = st.one_of(ints, st.booleans(), st.none())
request_fields = st.fixed_dictionaries(dict(
    epoch=field, batch_id=field, batch_size=field, sub_id=field, total_sub_batches=field,
    exp_packet=field,
    arrival_time=st.one_of(st.none(), st.floats(allow_nan=False, allow_infinity=True))))


def frame(obj):
    data = pickle.dumps(obj)
    return HDR.pack(len(data)) + data


def split(blob):
    out, pos = [], 0
    while pos < len(blob):
        (n,) = HDR.unpack_from(blob, pos)
        out.append(blob[pos + 4:pos + 4 + n])
        pos += 4 + n
    assert pos == len(blob)
    return out


def test_codec_self_check_passes_on_this_interpreter():
    assert wire.WireCodec(ServiceResponse, 5).self_check(ServiceRequest)


@settings(max_examples=150, deadline=None)
@given(cases=st.lists(request_fields, min_size=1, max_size=12), engine_id=st.integers(0, 300))
def test_requests_parse_to_their_fields_and_responses_unpickle_to_the_python_ones(cases, engine_id):
    codec = wire.WireCodec(ServiceResponse, engine_id)
    blob = b"".join(frame(ServiceRequest(**c)) for c in cases)
    recs, consumed, stop = codec.parse(blob)
    assert len(recs) == len(cases) and consumed == len(blob) and stop == wire.NEED_MORE
    recs["out_batch_size"] = np.arange(len(cases)) * 7
    recs["start_time"] = 10.5
    recs["end_time"] = 11.25
    frames = split(bytes(codec.emit(recs)))
    assert len(frames) == len(cases)
    for i, (c, payload) in enumerate(zip(cases, frames)):
        want = ServiceResponse(consumer_id=engine_id, epoch=c["epoch"], batch_id=c["batch_id"],
                               batch_size=c["batch_size"], arrival_time=c["arrival_time"],
                               process_start_time=10.5, queue_end_time=11.25,
                               inference_end_time=11.25, out_batch_size=i * 7,
                               total_sub_batches=c["total_sub_batches"],
                               exp_packet=c["exp_packet"], sub_id=c["sub_id"])
        assert payload == pickle.dumps(want)          # byte for byte what Queue.put would send
        got = pickle.loads(payload)
        assert type(got) is ServiceResponse
        assert list(got.__dict__.items()) == list(want.__dict__.items())
        assert [type(v) for v in got.__dict__.values()] == [type(v) for v in want.__dict__.values()]


def test_parse_stops_at_frames_that_need_python_and_at_partial_frames():
    codec = wire.WireCodec(ServiceResponse, 1)
    plain = ServiceRequest(batch_id=1, epoch=2, arrival_time=3.5, batch_size=4, sub_id=0,
                           total_sub_batches=1, exp_packet=False)
    f = frame(plain)
    for odd in (None,
                TensorServiceRequest(batch_id=1, epoch=2, arrival_time=3.5, batch_size=4, sub_id=0,
                                     total_sub_batches=1, exp_packet=False, ids=[np.arange(3)]),
                ServiceRequest(batch_id=1.5, epoch=2, arrival_time=3.5, batch_size=4, sub_id=0,
                               total_sub_batches=1, exp_packet=False),      # a float id
                ServiceRequest(batch_id=1, epoch=2, arrival_time=7, batch_size=4, sub_id=0,
                               total_sub_batches=1, exp_packet=False),      # an int arrival_time
                ServiceRequest(batch_id=2 ** 70, epoch=2, arrival_time=3.5, batch_size=4,
                               sub_id=0, total_sub_batches=1, exp_packet=False),   # > 64 bits
                ServiceRequest(batch_id="3", epoch=2, arrival_time=3.5, batch_size=4, sub_id=0,
                               total_sub_batches=1, exp_packet=False),
                {"batch_id": 1}, 17, "text"):
        recs, consumed, stop = codec.parse(f + f + frame(odd) + f)
        assert len(recs) == 2 and consumed == 2 * len(f) and stop == wire.PYTHON
    extra = ServiceRequest(batch_id=1, epoch=2, arrival_time=3.5, batch_size=4, sub_id=0,
                           total_sub_batches=1, exp_packet=False)
    extra.colour = 3                                         # an attribute the reference lacks
    recs, consumed, stop = codec.parse(f + frame(extra))
    assert len(recs) == 1 and stop == wire.PYTHON
    for cut in (1, 3, 4, 5, len(f) - 1):                     # a frame still on its way
        recs, consumed, stop = codec.parse(f + f[:cut])
        assert len(recs) == 1 and consumed == len(f) and stop == wire.NEED_MORE
    recs, consumed, stop = codec.parse(b"")
    assert len(recs) == 0 and consumed == 0 and stop == wire.NEED_MORE
    small = wire.WireCodec(ServiceResponse, 1, max_recs=3)
    recs, consumed, stop = small.parse(f * 5)
    assert len(recs) == 3 and consumed == 3 * len(f) and stop == wire.FULL


def test_requests_pickled_from_another_module_path_parse_too():
    """The load generator of a DeepRecSys checkout pickles ``utils.packets.ServiceRequest``."""
    blob = frame(ServiceRequest(batch_id=9, epoch=1, arrival_time=2.0, batch_size=30, sub_id=1,
                                total_sub_batches=2, exp_packet=True))
    blob = blob.replace(b"\x8c\x1ddeeprecsys_b200.utils.packets", b"\x8c\x0dutils.packets")
    body = blob[4:]
    # re-frame: the FRAME opcode's length and the "!i" header shrink with the module name
    body = body[:3] + struct.pack("<Q", len(body) - 11) + body[11:]
    recs, consumed, stop = wire.WireCodec(ServiceResponse, 1).parse(HDR.pack(len(body)) + body)
    assert len(recs) == 1 and stop == wire.NEED_MORE
    r = recs[0]
    assert (r["batch_id"], r["epoch"], r["batch_size"], r["sub_id"], r["total_sub_batches"],
            r["exp_packet"], r["arrival_time"]) == (9, 1, 30, 1, 2, 1, 2.0)


# ------------------------------------------------------------------ robustness of the C parser
class _GuardedBuffer(object):
    """Bytes placed so that the byte after the last one is an inaccessible page: a parser that
    reads past the end of its input dies here instead of passing by luck."""

    def __init__(self, max_len):
        import ctypes
        import mmap
        self.page = mmap.PAGESIZE
        self.pages = (max_len + self.page - 1) // self.page + 1
        self.map = mmap.mmap(-1, self.pages * self.page)
        self.base = ctypes.addressof(ctypes.c_char.from_buffer(self.map))
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        guard = self.base + (self.pages - 1) * self.page
        assert libc.mprotect(ctypes.c_void_p(guard), self.page, 0) == 0      # PROT_NONE
        self.limit = (self.pages - 1) * self.page

    def place(self, data):
        start = self.limit - len(data)
        self.map[start:self.limit] = data
        return self.base + start


_guard = _GuardedBuffer(1 << 16)


def _parse_guarded(codec, data):
    import ctypes as C
    from deeprecsys_b200 import _lib
    n, consumed, stop = C.c_int32(), C.c_int64(), C.c_int32()
    ptr = _guard.place(data) if data else _guard.base
    _lib.check(codec.lib.drs_wire_parse_requests(ptr, len(data), codec._modules,
                                                 codec._recs.ctypes.data, codec.max_recs,
                                                 C.byref(n), C.byref(consumed), C.byref(stop)))
    return n.value, consumed.value, stop.value


@settings(max_examples=400, deadline=None)
@given(cases=st.lists(request_fields, min_size=1, max_size=4), data=st.data())
def test_parser_never_reads_past_its_input_and_never_accepts_a_damaged_frame(cases, data):
    """Truncated, bit-flipped and spliced frames (the bytes of a pipe are not to be trusted): the
    parser stays inside its buffer (guard page), decodes only whole frames that pickle itself
    would decode to the same fields, and leaves everything else to the Python path."""
    codec = wire.WireCodec(ServiceResponse, 1, max_recs=16)
    frames = [frame(ServiceRequest(**c)) for c in cases]
    blob = bytearray(b"".join(frames))
    mode = data.draw(st.sampled_from(["truncate", "flip", "garbage", "intact"]))
    if mode == "truncate":
        blob = blob[:data.draw(st.integers(0, len(blob)))]
    elif mode == "flip":
        for _ in range(data.draw(st.integers(1, 4))):
            i = data.draw(st.integers(0, len(blob) - 1))
            blob[i] ^= 1 << data.draw(st.integers(0, 7))
    elif mode == "garbage":
        blob = bytearray(data.draw(st.binary(min_size=0, max_size=300)))
    n, consumed, stop = _parse_guarded(codec, bytes(blob))
    assert 0 <= consumed <= len(blob) and 0 <= n <= 16
    assert stop in (wire.NEED_MORE, wire.PYTHON, wire.FULL)
    # every frame the parser accepted is one pickle accepts too, with the same fields
    pos = 0
    for k in range(n):
        (sz,) = HDR.unpack_from(blob, pos)
        obj = pickle.loads(bytes(blob[pos + 4:pos + 4 + sz]))
        pos += 4 + sz
        assert type(obj) is ServiceRequest
        r = codec._recs[k]
        for i, name in enumerate(("epoch", "batch_id", "batch_size", "sub_id",
                                  "total_sub_batches", "exp_packet")):
            v, kind = getattr(obj, name), (int(r["kinds"]) >> (2 * i)) & 3
            assert kind == (2 if v is None else 1 if isinstance(v, bool) else 0)
            assert v is None or int(v) == int(r[name])
        at = obj.arrival_time
        assert (at is None and (int(r["kinds"]) >> 12) & 3 == 2) or at == r["arrival_time"] \
            or (at != at and r["arrival_time"] != r["arrival_time"])          # NaN stays NaN
    assert pos == consumed
