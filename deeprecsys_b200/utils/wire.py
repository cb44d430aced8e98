"""Queue wire codec of the serving loop: the C parser / emitter of ``csrc/wire.cpp`` behind numpy.

The reference's serving protocol is pickled packets over ``multiprocessing.Queue``s
(/root/reference/DeepRecSys.py:44-66, loadGenerator.py:162-177, accelInferenceEngine.py:46-83,
utils/packets.py:6-59).  With coalesced launches the GPU answers a 256-sample RMC1 query in a few
microseconds; what is left per packet is the interpreter: ``pickle.loads`` of the request, a
Python object per packet in the loop, constructing + pickling the ``ServiceResponse``.
``WireCodec`` keeps packets as rows of a numpy record array from the request pipe to the
response pipe:

* ``parse(buf)``: the leading frames of ``buf`` that are plain ``ServiceRequest`` pickles ->
  records (``drs_wire_parse_requests``); the first frame that is anything else — the final
  ``None``, a request that carries tensors, an unexpected field type — stops the parse and stays
  in ``buf`` for ``pickle.loads``;
* ``emit(recs)``: records -> the bytes ``Queue.put(ServiceResponse(...))`` would have put on the
  pipe for each of them (``drs_wire_emit_responses``), ready for one ``write``.

A stock ``Queue.get()`` on the other side cannot tell the difference; ``self_check()`` proves
that against ``pickle`` itself before the serving loop relies on it.
"""
from __future__ import absolute_import, division, print_function

import ctypes as C
import pickle
import struct

import numpy as np

from deeprecsys_b200 import _lib
from deeprecsys_b200.utils.response_ring import REC

NEED_MORE, PYTHON, FULL = 0, 1, 2
RESPONSE_FIELDS = ("consumer_id", "epoch", "batch_id", "batch_size", "arrival_time",
                   "queue_start_time", "queue_end_time", "inference_end_time", "out_batch_size",
                   "total_sub_batches", "exp_packet", "sub_id")
_HDR = struct.Struct("!i")


def _class_ops(cls):
    """The opcodes pickle protocol 4 uses to push ``cls``: module, qualname, STACK_GLOBAL."""
    out = b""
    for name in (cls.__module__, cls.__qualname__):
        raw = name.encode("utf-8")
        if len(raw) > 255:
            raise ValueError("class path too long for SHORT_BINUNICODE")
        out += b"\x8c" + bytes([len(raw)]) + raw + b"\x94"
    return out + b"\x93\x94"


class WireCodec(object):
    def __init__(self, response_cls, consumer_id, max_recs=1024, request_modules=None):
        """``request_modules``: module paths a ``ServiceRequest`` frame may name (default: the
        reference checkout's ``utils.packets`` and this package's); a frame naming any other
        module goes to ``pickle.loads``, which knows whether that module exists."""
        self.lib = _lib.load()
        mods = list(request_modules or ("utils.packets", "deeprecsys_b200.utils.packets"))
        self._modules = b"".join(m.encode("utf-8") + b"\0" for m in mods) + b"\0"
        self.response_cls = response_cls
        self.consumer_id = int(consumer_id)
        self.max_recs = int(max_recs)
        self._recs = np.zeros(self.max_recs, dtype=REC)
        self._out = C.create_string_buffer(self.max_recs * 384)
        self._n, self._consumed, self._stop = C.c_int32(), C.c_int64(), C.c_int32()
        self._out_len = C.c_int64()
        self._class_ops = _class_ops(response_cls)
        sample = response_cls(**{("process_start_time" if k == "queue_start_time" else k): 0
                                 for k in RESPONSE_FIELDS})
        keys = list(sample.__dict__)
        if sorted(keys) != sorted(RESPONSE_FIELDS):
            raise ValueError("%r does not have the reference's ServiceResponse attributes" %
                             (response_cls,))
        self._order = (C.c_int32 * 12)(*[RESPONSE_FIELDS.index(k) for k in keys])

    # ------------------------------------------------------------------ requests
    def parse(self, buf):
        """-> (records (a copy), bytes consumed, stop reason)."""
        view = np.frombuffer(buf, dtype=np.uint8)
        _lib.check(self.lib.drs_wire_parse_requests(
            view.ctypes.data, view.shape[0], self._modules, self._recs.ctypes.data, self.max_recs,
            C.byref(self._n), C.byref(self._consumed), C.byref(self._stop)))
        return self._recs[:self._n.value].copy(), self._consumed.value, self._stop.value

    # ------------------------------------------------------------------ responses
    def emit(self, recs):
        """-> memoryview of the frames (valid until the next ``emit``)."""
        recs = np.ascontiguousarray(recs, dtype=REC)
        n = recs.shape[0]
        if n * 384 > len(self._out):
            self._out = C.create_string_buffer(n * 384)
        _lib.check(self.lib.drs_wire_emit_responses(
            recs.ctypes.data, n, self._class_ops, len(self._class_ops), self._order,
            self.consumer_id, self._out, len(self._out), C.byref(self._out_len)))
        return memoryview(self._out)[:self._out_len.value]

    # ------------------------------------------------------------------ proof
    def self_check(self, request_cls):
        """The C codec against ``pickle`` on this interpreter: requests of ``request_cls`` parse
        to their fields; emitted frames unpickle to what the Python path would have sent
        (same class, attribute order, values and types).  -> True / False."""
        try:
            cases = [dict(epoch=3, batch_id=7, batch_size=256, sub_id=0, total_sub_batches=1,
                          exp_packet=False, arrival_time=1729000000.25),
                     dict(epoch=0, batch_id=255, batch_size=65535, sub_id=70000,
                          total_sub_batches=-2, exp_packet=True, arrival_time=0.0),
                     dict(epoch=None, batch_id=2 ** 40, batch_size=1, sub_id=True,
                          total_sub_batches=-2 ** 33, exp_packet=None, arrival_time=None)]
            frames = b""
            for c in cases:
                data = pickle.dumps(request_cls(**c))
                frames += _HDR.pack(len(data)) + data
            tail = pickle.dumps(None)
            recs, consumed, stop = self.parse(frames + _HDR.pack(len(tail)) + tail)
            if len(recs) != len(cases) or consumed != len(frames) or stop != PYTHON:
                return False
            recs["out_batch_size"] = [256, 0, 1]
            recs["start_time"] = [1729000000.5, 1.5, 2.5]
            recs["end_time"] = [1729000001.75, 3.25, 4.0]
            blob = bytes(self.emit(recs))
            pos = 0
            for c, r in zip(cases, recs):
                (sz,) = _HDR.unpack_from(blob, pos)
                got = pickle.loads(blob[pos + 4:pos + 4 + sz])
                pos += 4 + sz
                want = self.response_cls(
                    consumer_id=self.consumer_id, epoch=c["epoch"], batch_id=c["batch_id"],
                    batch_size=c["batch_size"], arrival_time=c["arrival_time"],
                    process_start_time=float(r["start_time"]), queue_end_time=float(r["end_time"]),
                    inference_end_time=float(r["end_time"]),
                    out_batch_size=int(r["out_batch_size"]),
                    total_sub_batches=c["total_sub_batches"], exp_packet=c["exp_packet"],
                    sub_id=c["sub_id"])
                if type(got) is not self.response_cls:
                    return False
                if list(got.__dict__.items()) != list(want.__dict__.items()):
                    return False
                if [type(v) for v in got.__dict__.values()] != \
                        [type(v) for v in want.__dict__.values()]:
                    return False
            return pos == len(blob)
        except Exception:
            return False
