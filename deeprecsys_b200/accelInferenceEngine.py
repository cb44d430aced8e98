"""Drop-in accelerator engine process for DeepRecSys.

Same signature and Queue protocol as the reference's
``accelInferenceEngine(args, requestQueue, engine_id, responseQueue,
inferenceEngineReadyQueue)`` (/root/reference/accelInferenceEngine.py:18-86),
which DeepRecSys.py:63-66 starts as the last engine process.  The reference
body *sleeps* for a GTX-1080-Ti latency looked up in
accelerator/predict_execution.py; here each packet is actually executed on a
B200 through libdrs_b200.so.

Protocol kept:
  1. seed numpy's global RNG with args.numpy_rand_seed            (ref :25)
  2. requestQueue is None -> print + sys.exit()                   (ref :28-31)
  3. exactly one ready token — put *after* CUDA init, parameter upload,
     batch staging and warm-up (the reference signals before its own init, so
     a failed init hangs the parent; see step 6)                  (ref :34)
  4. loop: request = requestQueue.get(); None -> responseQueue.put(None), return
                                                                  (ref :46-53)
  5. one ServiceResponse per request with the reference's kwargs  (ref :68-80)
  6. never die silently: if the backend cannot be built the engine still
     signals ready, answers every request with out_batch_size = 0 and forwards
     the sentinel, so DeepRecSys.py terminates instead of polling forever.

Like every engine process of the reference, this one regenerates the synthetic
dataset and the parameters locally from the seed (inferenceEngine.py:72-88:
datagen inputs -> datagen targets -> tables -> MLPs) and serves a request by
running the first ``batch_size`` rows of pre-generated batch ``batch_id``
(inferenceEngine.py:200-215).  CUDA is initialised inside this (forked) child,
never in the parent.
"""
from __future__ import absolute_import, division, print_function

import os
import sys
import time
import traceback

import numpy as np

try:                                   # running inside a DeepRecSys checkout
    from utils.packets import ServiceRequest, ServiceResponse
    from utils.utils import debugPrint
except ImportError:                    # standalone
    from deeprecsys_b200.utils.packets import ServiceRequest, ServiceResponse
    from deeprecsys_b200.utils.utils import debugPrint

# the reference's whitelist (accelInferenceEngine.py:37) plus ncf
SUPPORTED_MODEL_NAMES = ("wnd", "rm1", "rm2", "rm3", "din", "dien", "mtwnd", "ncf")


class AccelBackend(object):
    """Data + engine of one accel process: build once, then ``serve`` packets."""

    def __init__(self, args):
        from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
        from deeprecsys_b200.engine import AccelEngine, bind_to_device_numa_node

        self.args = args
        bind_to_device_numa_node(int(getattr(args, "accel_device", 0)))
        gen = DLRMDataGenerator(args)
        (self.nbatches, self.lX, self.lS_l, self.lS_i) = gen.generate_input_data()
        (_, self.lT) = gen.generate_output_data()          # consumes the RNG like the CPU engine
        # coalesced serving stacks several queued requests into one launch: the engine's
        # activation buffers hold accel_stack_rows samples (a staged batch still holds
        # max_mini_batch_size)
        self.coalesce = max(1, int(getattr(args, "accel_coalesce", 1) or 1))
        cap = int(args.max_mini_batch_size)
        if self.coalesce > 1:
            cap = max(cap, int(getattr(args, "accel_stack_rows", 0) or 0))
        self.engine = AccelEngine(args=args, device=int(getattr(args, "accel_device", 0)),
                                  max_batch=cap, lanes=int(getattr(args, "accel_lanes", 0) or 0),
                                  max_lookups_per_table=cap
                                  * max(1, int(args.num_indices_per_lookup)))
        self.stack_rows = cap
        # query sizes vary from packet to packet: a (lane, batch, size) step hardly ever repeats,
        # so capturing CUDA graphs of single steps would only cost; load is absorbed by coalescing
        self.engine.set_option("cuda_graphs", 0)
        self.engine.create()                               # tables -> MLPs, reference draw order
        if int(getattr(args, "accel_host_pack", 0) or 0) > 0:
            self.engine.set_option("host_pack", int(args.accel_host_pack))
        self.spec = self.engine.spec
        self.upload = bool(getattr(args, "accel_upload_inputs", False))
        T = self.spec.T
        if not self.upload:
            for j in range(self.nbatches):
                self.engine.stage_batch(
                    j, self.lX[j] if self.spec.m_den > 0 else None,
                    [self.lS_i[j][t] for t in range(T)], [self.lS_l[j][t] for t in range(T)])
        else:
            self._pin_dataset()
        self.dump_dir = os.environ.get("DRS_ACCEL_DUMP_DIR")
        if self.nbatches > 0:                              # warm-up: first-launch costs off the clock
            n = min(int(args.max_mini_batch_size), 16)
            self.serve(0, n)
            if self.can_coalesce and 2 * n <= self.stack_rows:
                for _ in range(self.depth):                # every lane's coalesced launch once
                    self.complete(self.submit_many([0, 0], [n, n]))

    def _pin_dataset(self):
        """Upload-per-query mode: keep the pre-generated batches in pinned host memory (ids already
        cast to int32, the dtype the reference's Cast produces, dlrm_s_caffe2.py:308) so that each
        packet's H2D copies are true asynchronous DMA and can overlap the previous packet's
        kernels.  The first B rows / B*L ids of a batch are a prefix of these buffers, so a packet
        needs no host-side copy at all."""
        import ctypes
        from deeprecsys_b200.engine import pinned_empty
        T, L = self.spec.T, self.spec.L
        self._pinned = []
        for j in range(self.nbatches):
            n = int(np.asarray(self.lS_l[j][0]).shape[0])
            ids, o1 = pinned_empty((T, n * L), np.int32)
            lens, o2 = pinned_empty((T, n), np.int32)
            for t in range(T):
                ids[t] = self.lS_i[j][t]
                lens[t] = self.lS_l[j][t]
            den, o3 = (None, None)
            if self.spec.m_den > 0:
                den, o3 = pinned_empty((n, self.spec.m_den), np.float32)
                den[:] = self.lX[j]
            self._pinned.append({
                "ids": ids, "lens": lens, "dense": den, "owners": (o1, o2, o3),
                "ids_ptrs": (ctypes.c_void_p * T)(*[ids[t].ctypes.data for t in range(T)]),
                "len_ptrs": (ctypes.c_void_p * T)(*[lens[t].ctypes.data for t in range(T)]),
                "dense_ptr": ctypes.c_void_p(den.ctypes.data) if den is not None else None})

    def serve(self, batch_id, batch_size):
        """Run rows [0, batch_size) of batch ``batch_id``; returns prob_click [B, out_dim]."""
        B = int(batch_size)
        if self.upload:
            T, L = self.spec.T, self.spec.L
            out, _ = self.engine.infer(
                self.lX[batch_id][:B] if self.spec.m_den > 0 else None,
                [self.lS_i[batch_id][t][:B * L] for t in range(T)],
                [self.lS_l[batch_id][t][:B] for t in range(T)], B)
        else:
            out, _ = self.engine.infer_staged(batch_id, B)
        return out

    # pipelined serving: up to ``depth`` packets in flight (engine lanes overlap the MLP kernels
    # of one packet with the gather of the next)
    @property
    def depth(self):
        return self.engine.lane_count()

    def _next_out(self):
        import ctypes
        if not hasattr(self, "_out_pool"):
            from deeprecsys_b200.engine import pinned_empty
            self._out_pool = [pinned_empty((self.engine.max_batch, self.spec.out_dim), np.float32)
                              for _ in range(self.depth + 1)]
            self._out_next = 0
        buf, _owner = self._out_pool[self._out_next]
        self._out_next = (self._out_next + 1) % len(self._out_pool)
        return buf, ctypes.c_void_p(buf.ctypes.data)

    def submit(self, batch_id, batch_size):
        buf, out_ptr = self._next_out()
        if self.upload:                    # H2D of this packet's inputs from the pinned dataset
            p = self._pinned[batch_id]
            ticket = self.engine.submit_host(p["dense_ptr"], p["ids_ptrs"], 0, p["len_ptrs"],
                                             int(batch_size), out_ptr)
        else:
            ticket = self.engine.submit_staged(batch_id, int(batch_size), out_ptr)
        return ("ticket", ticket, buf, int(batch_size))

    def submit_many(self, batch_ids, batch_sizes):
        """Several queued requests as ONE stacked launch (drs_submit_many); rows of request k
        start at sum(batch_sizes[:k]) of the completed buffer."""
        buf, out_ptr = self._next_out()
        ticket = self.engine.submit_many(batch_ids, batch_sizes, out_ptr)
        return ("ticket", ticket, buf, int(np.sum(batch_sizes)))

    @property
    def can_coalesce(self):
        return self.coalesce > 1 and not self.upload

    # requests that bring their own tensors (utils.packets.TensorServiceRequest) ----------------
    def _attach_ring(self, desc):
        """Map a producer's shared-memory ring once and page-lock it, so that the H2D copies of
        its slots are asynchronous DMA straight from the producer's bytes."""
        from deeprecsys_b200.utils.input_ring import SharedInputRing
        if not hasattr(self, "_rings"):
            self._rings = {}
        ring = self._rings.get(desc["path"])
        if ring is None:
            ring = SharedInputRing.attach(desc)
            if ring.T != self.spec.T or ring.m_den != self.spec.m_den:
                raise ValueError("ring %s is laid out for %d tables / %d dense features, the model "
                                 "has %d / %d" % (desc["path"], ring.T, ring.m_den, self.spec.T,
                                                  self.spec.m_den))
            if ring.max_batch > self.engine.max_batch:
                raise ValueError("ring %s holds queries of up to %d samples, the engine serves at "
                                 "most %d" % (desc["path"], ring.max_batch, self.engine.max_batch))
            try:
                from deeprecsys_b200 import _lib
                _lib.check(self.engine.lib.drs_host_register(ring.base_address(), ring.nbytes))
                ring.registered = True
            except Exception as exc:       # pageable memory still works: the driver stages it
                print("[Accel] ring %s not page-locked (%r); copies will be staged"
                      % (desc["path"], exc))
            ring._ptr_cache = {}
            self._rings[desc["path"]] = ring
        return ring

    def submit_inputs(self, request):
        import ctypes
        B, T = int(request.batch_size), self.spec.T
        buf, out_ptr = self._next_out()
        if request.ring is not None:
            ring = self._attach_ring(request.ring)
            slot = int(request.slot)
            # never trust the descriptor: a slot outside the ring or a query larger than a slot
            # would make the host loop and the DMA read a neighbouring slot or unmapped memory
            if not 0 <= slot < ring.n_slots:
                raise ValueError("request names slot %d of a %d-slot ring" % (slot, ring.n_slots))
            release = lambda: ring.release(slot)
            try:
                if not 0 < B <= ring.max_batch:
                    raise ValueError("batch_size %d outside the ring's (0, %d]" % (B, ring.max_batch))
                counts = request.counts
                if counts is not None and (len(counts) != T or
                                           max(int(c) for c in counts) > ring.max_lookups):
                    raise ValueError("request counts %r do not fit a slot of %d lookups x %d tables"
                                     % (counts, ring.max_lookups, T))
                if int(ring.lengths(slot)[:, :B].sum(axis=1).max()) > ring.max_lookups:
                    raise ValueError("slot %d: sum(lengths) exceeds the slot's %d lookups per table"
                                     % (slot, ring.max_lookups))
                ptrs = ring._ptr_cache.get(slot)
                if ptrs is None:
                    d, i, l = ring.pointers(slot)
                    ptrs = (ctypes.c_void_p(d) if d is not None else None,
                            (ctypes.c_void_p * T)(*i), (ctypes.c_void_p * T)(*l))
                    ring._ptr_cache[slot] = ptrs
                ticket = self.engine.submit_host(ptrs[0], ptrs[1], 0, ptrs[2], B, out_ptr)
            except Exception:
                release()                      # a refused query must not leave its slot busy
                raise
            return ("ticket", ticket, buf, B, None, release)
        dense, dptr, ids, is64, lengths = self.engine._prep(request.dense, request.ids,
                                                            request.lengths, B)
        ticket = self.engine.submit_host(
            dptr, (ctypes.c_void_p * T)(*[a.ctypes.data for a in ids]), is64,
            (ctypes.c_void_p * T)(*[a.ctypes.data for a in lengths]), B, out_ptr)
        return ("ticket", ticket, buf, B, (dense, ids, lengths))     # arrays alive until complete

    def serve_inputs(self, request):
        return self.complete(self.submit_inputs(request))

    def complete(self, handle):
        if handle[0] == "done":
            return handle[1]
        ticket, buf, B = handle[1:4]
        try:
            self.engine.wait(ticket)
        finally:
            if len(handle) > 5 and handle[5] is not None:
                handle[5]()                    # ring slot: its H2D copies are done, hand it back
        return buf[:B]

    def close(self):
        for ring in getattr(self, "_rings", {}).values():
            if ring.registered:
                self.engine.lib.drs_host_unregister(ring.base_address())
            ring.close()
        self.engine.close()


def _carries_inputs(request):
    has = getattr(request, "has_inputs", None)
    return bool(has and has())


# tests replace this with an oracle-backed fake (no GPU in the authoring container)
BACKEND_FACTORY = AccelBackend


def _direct_put(q):
    """A ``put`` for a multiprocessing.Queue that pickles and writes to the queue's pipe in the
    calling thread.

    ``Queue.put`` only appends to a buffer and wakes a feeder thread, which then has to win the
    GIL from the serving loop before the response reaches the pipe: measured with a zero-latency
    backend that hand-off adds ~1 ms to the median response and caps one engine process at
    ~12 k responses/s; written directly, the median is 0.15 ms and the cap ~2x higher.  The
    bytes on the pipe and the semaphore accounting are exactly those of ``Queue.put``
    (multiprocessing/queues.py: put + _feed), so the reader's plain ``get()`` is unaffected and
    several processes may write to the same queue (the queue's own cross-process write lock is
    taken).  A full pipe blocks the caller until the parent reads — back-pressure instead of an
    unbounded buffer.  Falls back to ``q.put`` for anything that is not a stock Queue, or when
    DRS_ACCEL_QUEUE_FEEDER=1."""
    if os.environ.get("DRS_ACCEL_QUEUE_FEEDER") == "1":
        return q.put
    try:
        from multiprocessing.reduction import ForkingPickler
        sem, writer, wlock = q._sem, q._writer, q._wlock
        import pickle
        send_bytes, dumps, plain_dumps = writer.send_bytes, ForkingPickler.dumps, pickle.dumps
        if getattr(q, "_thread", None) is not None or getattr(q, "_joincancelled", None) is None:
            return q.put                   # feeder already running (order!) / not a stock Queue
    except AttributeError:
        return q.put

    def put(obj):
        try:
            data = plain_dumps(obj)        # packets are plain python objects: no need to build a
        except Exception:                  # ForkingPickler (a Pickler + dispatch table) per call
            data = dumps(obj)
        sem.acquire()
        if wlock is None:
            send_bytes(data)
        else:
            with wlock:
                send_bytes(data)
    return put


def _die_with_parent():
    """Linux: deliver SIGTERM to this process when its parent dies (DeepRecSys.py:179 ends the
    run by terminate()-ing the engine processes; sibling GPU workers must not outlive that)."""
    try:
        import ctypes
        import signal
        ctypes.CDLL("libc.so.6", use_errno=True).prctl(1, signal.SIGTERM)     # PR_SET_PDEATHSIG
    except Exception:
        pass


def _spawn_siblings(args, requestQueue, engine_id, responseQueue):
    """--accel_gpus N > 1: the single accel slot of DeepRecSys.py fronts N GPUs.  N-1 sibling
    processes (one per extra GPU, replicas of the whole model) consume the SAME accelRequestQueue
    and answer on the SAME responseQueue.  Forked before this process touches CUDA."""
    n = int(getattr(args, "accel_gpus", 1) or 1)
    if n <= 1:
        return [], None
    if getattr(args, "accel_shard", "replica") != "replica":
        raise ValueError("--accel_shard %s: the Queue-fed engine serves replicas only (sharded "
                         "ranks must see every query in lock-step; see bench.py --shard)"
                         % args.accel_shard)
    import copy
    import multiprocessing as mp
    cur = mp.current_process()
    was_daemon = cur._config.get("daemon")
    cur._config["daemon"] = False      # DeepRecSys.py:71 marks engines daemonic; daemons may not fork
    try:
        ctx = mp.get_context("fork")
        ready = ctx.Queue()
        procs = []
        for g in range(1, n):
            a = copy.copy(args)
            a.accel_device = int(getattr(args, "accel_device", 0)) + g
            a.accel_gpus = 1
            p = ctx.Process(target=_sibling_main,
                            args=(a, requestQueue, engine_id, responseQueue, ready))
            p.start()
            procs.append(p)
    finally:
        cur._config["daemon"] = was_daemon
    return procs, ready


def _sibling_main(args, requestQueue, engine_id, responseQueue, readyQueue):
    _die_with_parent()
    _engine_process(args, requestQueue, engine_id, responseQueue, readyQueue, sibling=True)


def accelInferenceEngine(args,
                         requestQueue=None,
                         engine_id=None,
                         responseQueue=None,
                         inferenceEngineReadyQueue=None):
    if requestQueue is None:
        print("If you want to run Accel in isolation please use bench.py / "
              "deeprecsys_b200.serve directly")
        sys.stdout.flush()
        sys.exit()
    # under the reference's own cli() (a DeepRecSys checkout: utils/utils.py is theirs) the
    # namespace has none of this backend's flags: they take their defaults, as with our cli()
    from deeprecsys_b200.utils.utils import ACCEL_FLAGS
    for name, _type, default in ACCEL_FLAGS:
        if not hasattr(args, name):
            setattr(args, name, default)
    _engine_process(args, requestQueue, engine_id, responseQueue, inferenceEngineReadyQueue,
                    sibling=False)


def _engine_process(args, requestQueue, engine_id, responseQueue, readyQueue, sibling):
    """DRS_ACCEL_PROFILE=<file>: run the serve loop under cProfile and write the hottest
    functions there when the engine exits (where a packet's host time goes)."""
    prof_path = os.environ.get("DRS_ACCEL_PROFILE")
    if not prof_path:
        return _engine_loop(args, requestQueue, engine_id, responseQueue, readyQueue, sibling)
    import cProfile
    import pstats
    prof = cProfile.Profile()
    try:
        return prof.runcall(_engine_loop, args, requestQueue, engine_id, responseQueue,
                            readyQueue, sibling)
    finally:
        with open(prof_path + (".%d" % os.getpid() if sibling else ""), "w") as f:
            pstats.Stats(prof, stream=f).sort_stats("tottime").print_stats(40)


def _wait_for_siblings(siblings, requestQueue, responseQueue, engine_id, args):
    """Primary of an --accel_gpus slot whose own backend failed while siblings are healthy: serve
    nothing, let the siblings drain the queue (the sentinel circulates among them), then answer
    whatever is left with out_batch_size = 0 and forward the final sentinel."""
    import queue as _queue
    for p in siblings:
        p.join()
    put_response = _direct_put(responseQueue)
    while True:
        try:
            request = requestQueue.get(timeout=1.0)
        except _queue.Empty:
            continue
        if request is None:
            put_response(None)
            return
        now = time.time()
        put_response(ServiceResponse(consumer_id=engine_id, epoch=request.epoch,
                                     batch_id=request.batch_id, batch_size=request.batch_size,
                                     arrival_time=request.arrival_time, process_start_time=now,
                                     queue_end_time=now, inference_end_time=now, out_batch_size=0,
                                     total_sub_batches=request.total_sub_batches,
                                     exp_packet=request.exp_packet, sub_id=request.sub_id))


def _spawn_responder(args, responseQueue, engine_id, multi):
    """Fork the responder helper (utils/response_ring.py) BEFORE this process touches CUDA.
    Returns (ring, process) or (None, None): several consumers on one queue, a non-stock queue,
    or --accel_responder 0 keep the responses in-process."""
    if multi or not int(getattr(args, "accel_responder", 0) or 0):
        return None, None
    if int(getattr(args, "accel_coalesce", 1) or 1) <= 1:
        return None, None
    from deeprecsys_b200.utils.fastqueue import FrameWriter
    if not FrameWriter(responseQueue).fast:
        return None, None
    try:
        import multiprocessing as mp
        from deeprecsys_b200.utils.response_ring import ResponseRing, responder_main
        ring = ResponseRing()
        cur = mp.current_process()
        was_daemon = cur._config.get("daemon")
        cur._config["daemon"] = False          # DeepRecSys.py:71 marks engines daemonic
        try:
            proc = mp.get_context("fork").Process(
                target=responder_main, args=(ring, responseQueue, engine_id, ServiceResponse))
            proc.start()
        finally:
            cur._config["daemon"] = was_daemon
        return ring, proc
    except Exception as exc:
        print("[Accel] responder helper unavailable (%r); responses stay in-process" % (exc,))
        return None, None


def _wire_codec(args, backend, engine_id, reader, writer, multi, coalesce, pipelined):
    """The C wire codec for this engine, or None when the record path does not apply: several
    consumers on the queue, non-stock queues, one launch per request, debug prints,
    DRS_ACCEL_WIRE=0, or a codec that fails its self-check against ``pickle``."""
    if (multi or reader is None or not reader.fast or not writer.fast or coalesce <= 1
            or not pipelined or backend is None
            or getattr(args, "debug_mode", False) or os.environ.get("DRS_ACCEL_WIRE") == "0"
            or not hasattr(backend, "nbatches")):
        return None
    try:
        from deeprecsys_b200.utils.wire import WireCodec
        codec = WireCodec(ServiceResponse, engine_id)
        if not codec.self_check(ServiceRequest):
            print("[Accel] wire codec failed its self-check against pickle; packets stay Python "
                  "objects")
            return None
        return codec
    except Exception as exc:
        print("[Accel] wire codec unavailable (%r); packets stay Python objects" % (exc,))
        return None


def _serve_records(args, backend, codec, reader, requestQueue, responseQueue, depth, coalesce,
                   stack_rows):
    """Serving loop over packet RECORDS (utils/wire.py): request frames are parsed in C straight
    from the pipe's bytes into a numpy record array, groups of waiting requests are cut with
    array arithmetic and served by ONE stacked launch each (``drs_submit_many``), and the
    ServiceResponse frames of a finished group are emitted in C and written with one ``write``.
    Same protocol, same fields, same order of events as the object loop below
    (accelInferenceEngine.py:46-83 of the reference: pull, serve, respond) — without a Python
    object per packet.

    Returns when the next frame on the pipe is not a plain ServiceRequest (the final ``None``, a
    request that carries tensors, an odd field type): everything before it has been answered,
    the frame and what follows are left in ``reader._buf`` for the object loop."""
    import collections
    from deeprecsys_b200.utils.wire import PYTHON
    fd, poll, rlock = reader._fd, reader._poll, requestQueue._rlock
    release = requestQueue._sem.release
    wfd, wlock, acquire = responseQueue._writer.fileno(), responseQueue._wlock, \
        responseQueue._sem.acquire
    nbatches = int(backend.nbatches)
    max_rows = int(args.max_mini_batch_size)
    coalesce = min(int(coalesce), 32)                      # drs_submit_many: <= 32 requests
    # batch_id (field 1) and batch_size (field 2) must be numbers (a bool counts, as int(True)
    # does in the object loop), not None: kind 2 in their two-bit slots of `kinds`
    buf = reader._buf
    inflight = collections.deque()                         # (records, handle, rows)
    st = {"recs": None, "head": 0, "bail": False, "served": 0}
    dump_dir = getattr(backend, "dump_dir", None)

    def pull(block):
        nonlocal buf
        while True:
            if len(buf) >= 4:
                recs, consumed, stop = codec.parse(buf)
                if consumed:
                    buf = buf[consumed:]
                n = len(recs)
                if stop == PYTHON:
                    st["bail"] = True
                if n:
                    for _ in range(n):
                        release()                          # Queue.get: a free slot per packet
                    recs["start_time"] = time.time()
                    sizes = recs["batch_size"]
                    kinds = recs["kinds"]
                    ok = ((sizes >= 1) & (sizes <= max_rows) & (recs["batch_id"] >= 0)
                          & (recs["batch_id"] < nbatches)
                          & (((kinds >> 2) & 3) != 2) & (((kinds >> 4) & 3) != 2))
                    st["recs"], st["head"] = recs, 0
                    st["ok"] = ok
                    st["bad"] = np.flatnonzero(~ok)
                    st["cum"] = np.cumsum(np.where(ok, sizes, 0))
                    st["ids32"] = recs["batch_id"].astype(np.int32)
                    st["sizes32"] = np.where(ok, sizes, 0).astype(np.int32)
                    return True
                if st["bail"]:
                    return False
            if not block and not poll.poll(0):
                return False
            with rlock:
                chunk = os.read(fd, 1 << 16)
            if not chunk:
                raise EOFError("queue pipe closed")
            buf = buf + chunk if buf else chunk
            block = False                                  # parse what came; wait again if partial
            if len(buf) < 4:
                block = True

    def respond(recs):
        blob = codec.emit(recs)
        for _ in range(len(recs)):
            acquire()                                      # Queue.put: a slot per packet
        with wlock:
            while blob:
                blob = blob[os.write(wfd, blob):]

    def finish_oldest():
        recs, handle, rows = inflight.popleft()
        try:
            out = backend.complete(handle)
            if int(out.shape[0]) != rows:
                raise RuntimeError("coalesced launch returned %d rows for %d" % (out.shape[0], rows))
            recs["out_batch_size"] = recs["batch_size"]
            if dump_dir:                                   # the parity side channel (packets
                row = 0                                    # carry no tensors)
                for r in recs:
                    n = int(r["batch_size"])
                    np.save(os.path.join(dump_dir, "accel_e%s_b%s_n%s.npy" % (
                        int(r["epoch"]), int(r["batch_id"]), n)), np.array(out[row:row + n]))
                    row += n
        except Exception as exc:                           # keep serving: the parent counts responses
            print("[Accel] request failed: %r" % (exc,))
            sys.stdout.flush()
            recs["out_batch_size"] = 0
        recs["end_time"] = time.time()
        respond(recs)
        st["served"] += len(recs)

    while True:
        if len(inflight) >= depth:
            finish_oldest()
            continue
        recs, head = st["recs"], st["head"]
        if recs is None or head >= len(recs):
            if st["bail"]:
                break
            st["recs"] = None
            if not pull(block=not inflight):
                if inflight:
                    finish_oldest()                        # nothing waiting: complete the oldest
                continue
            recs, head = st["recs"], 0
        if not st["ok"][head]:
            bad = recs[head:head + 1]                      # refused: answered with 0 rows at once
            print("[Accel] request refused: batch_id %d, batch_size %d"
                  % (int(bad["batch_id"][0]), int(bad["batch_size"][0])))
            sys.stdout.flush()
            bad["out_batch_size"] = 0
            bad["end_time"] = time.time()
            respond(bad)
            st["head"] = head + 1
            continue
        # every valid request that is already here: up to `coalesce` requests / stack_rows samples
        cum, end = st["cum"], min(len(recs), head + coalesce)
        nb = st["bad"]
        if len(nb):
            k = int(np.searchsorted(nb, head))
            if k < len(nb):
                end = min(end, int(nb[k]))
        base = int(cum[head - 1]) if head else 0
        end = min(end, int(np.searchsorted(cum, base + stack_rows, side="right")))
        end = max(end, head + 1)
        group = recs[head:end]
        rows = int(cum[end - 1]) - base
        try:
            if end - head == 1:                            # a lone request: the ordinary launch
                handle = backend.submit(int(st["ids32"][head]), rows)
            else:
                handle = backend.submit_many(st["ids32"][head:end], st["sizes32"][head:end])
            inflight.append((group, handle, rows))
        except Exception as exc:
            print("[Accel] coalesced launch refused: %r" % (exc,))
            sys.stdout.flush()
            group["out_batch_size"] = 0
            group["end_time"] = time.time()
            respond(group)
        st["head"] = end
    while inflight:
        finish_oldest()
    reader._buf = buf
    if os.environ.get("DRS_ACCEL_WIRE_REPORT") == "1":
        print("[Accel] record loop answered %d packets" % st["served"])
        sys.stdout.flush()


def _engine_loop(args, requestQueue, engine_id, responseQueue, readyQueue, sibling):
    np.random.seed(args.numpy_rand_seed)
    np.set_printoptions(precision=args.print_precision)

    siblings, sibling_ready = ([], None) if sibling else _spawn_siblings(
        args, requestQueue, engine_id, responseQueue)
    multi = sibling or bool(siblings)
    resp_ring, responder = _spawn_responder(args, responseQueue, engine_id, multi)

    backend, failure = None, None
    if args.model_name not in SUPPORTED_MODEL_NAMES:
        failure = "Model %r not found in ones supported %s" % (args.model_name,
                                                               list(SUPPORTED_MODEL_NAMES))
    else:
        try:
            backend = BACKEND_FACTORY(args)
        except BaseException as exc:       # incl. SystemExit from spec checks
            failure = "accel backend init failed: %r\n%s" % (exc, traceback.format_exc())
    if failure:
        print("[Accel] " + failure)
        sys.stdout.flush()

    if sibling:
        # A replica that could not be built must not touch the shared accelRequestQueue: it would
        # answer every packet instantly with out_batch_size = 0 and, doing no work, drain the
        # queue faster than the healthy replicas.  It reports its state and leaves; "never die
        # silently" (answer with 0, forward the sentinel) is the primary's job alone.
        readyQueue.put(backend is not None)
        if backend is None:
            return
    else:
        healthy = sum(1 for _ in siblings if sibling_ready.get())
        if backend is None and healthy > 0:
            print("[Accel] primary replica unavailable; %d sibling replica(s) serve the queue"
                  % healthy)
            sys.stdout.flush()
            readyQueue.put(True)           # one ready token for the whole accel slot
            return _wait_for_siblings(siblings, requestQueue, responseQueue, engine_id, args)
        readyQueue.put(True)

    # multiprocessing.Queue.put hands the response to a feeder thread; while this loop is busy
    # (requests always pending) that thread only gets the GIL at a switch interval, 5 ms by
    # default -> responses would leave up to 5 ms late exactly when the engine is loaded
    sys.setswitchinterval(1e-4)

    # Packets are served in arrival order with up to ``depth`` launches in flight on the GPU; a
    # launch serves ONE request, or — coalesced serving — every plain request that was already
    # waiting when the loop came back to the queue (up to --accel_coalesce of them, at most
    # --accel_stack_rows samples).  A response is stamped when its own result has landed in host
    # memory.
    import collections
    from deeprecsys_b200.utils.fastqueue import FrameReader, FrameWriter
    from deeprecsys_b200.utils.response_ring import KIND_SENTINEL, encode_request
    writer = FrameWriter(responseQueue)
    reader = None if multi else FrameReader(requestQueue)      # sole consumer: batched pipe reads
    inflight = collections.deque()         # (entries [(request, start_time)], handle, sizes | None)
    pending = collections.deque()          # pulled from the queue, not yet submitted
    pipelined = backend is not None and hasattr(backend, "submit") and backend.depth > 1
    depth = backend.depth if pipelined else 1
    coalesce = 1
    if (pipelined and not multi and hasattr(backend, "submit_many")
            and getattr(backend, "can_coalesce", True)):
        coalesce = max(1, int(getattr(args, "accel_coalesce", 1) or 1))
    stack_rows = int(getattr(backend, "stack_rows", 0) or getattr(args, "max_mini_batch_size", 0))

    # Packets as records: frames are parsed / emitted in C and the loop works on numpy record
    # arrays (utils/wire.py) until a frame needs the Python path — the final None at the latest —
    # which the object loop below then handles (everything before it has been answered).
    codec = _wire_codec(args, backend, engine_id, reader, writer, multi, coalesce, pipelined)
    if codec is not None:
        _serve_records(args, backend, codec, reader, requestQueue, responseQueue, depth, coalesce,
                       stack_rows)

    def respond_many(entries, out_sizes):
        """One response per request; out_sizes[i] = rows answered for entries[i] (0 = failed)."""
        end_time = time.time()
        debugPrint(args, "Accel", "Sending response back")
        rows, plain = [], []
        ring_ok = resp_ring is not None and responder.is_alive()
        for (request, start_time), out_bs in zip(entries, out_sizes):
            enc = encode_request(request) if ring_ok else None
            if enc is None:
                plain.append(ServiceResponse(consumer_id=engine_id,
                                             epoch=request.epoch,
                                             batch_id=request.batch_id,
                                             batch_size=request.batch_size,
                                             arrival_time=request.arrival_time,
                                             process_start_time=start_time,
                                             queue_end_time=end_time,
                                             inference_end_time=end_time,
                                             out_batch_size=out_bs,
                                             total_sub_batches=request.total_sub_batches,
                                             exp_packet=request.exp_packet,
                                             sub_id=request.sub_id))
            else:
                vals, kinds, at = enc
                rows.append(vals[:5] + (int(out_bs), vals[5], kinds, at, start_time, end_time))
        if rows:
            resp_ring.push(rows)
        if plain:
            writer.put_many(plain)

    def dump(request, out):
        if backend.dump_dir:
            np.save(os.path.join(backend.dump_dir, "accel_e%s_b%s_n%s.npy" % (
                request.epoch, request.batch_id, request.batch_size)), np.array(out))

    def finish_oldest():
        entries, handle, sizes = inflight.popleft()
        try:
            out = backend.complete(handle)
            if sizes is None:              # one request
                out_sizes = [int(out.shape[0])]
                dump(entries[0][0], out)
            else:                          # coalesced: rows of request k start at sum(sizes[:k])
                out_sizes, row = list(sizes), 0
                if int(out.shape[0]) != sum(sizes):
                    raise RuntimeError("coalesced launch returned %d rows for %d"
                                       % (out.shape[0], sum(sizes)))
                if backend.dump_dir:
                    for (request, _), n in zip(entries, sizes):
                        dump(request, out[row:row + n])
                        row += n
        except Exception as exc:           # keep serving: the parent counts responses
            print("[Accel] request failed: %r" % (exc,))
            sys.stdout.flush()
            out_sizes = [0] * len(entries)
        respond_many(entries, out_sizes)

    def pull(block):
        """Move what is waiting on the queue into ``pending``; -> True if anything arrived."""
        debugPrint(args, "Accel", "Trying to pull request")
        if reader is not None:
            got = reader.drain(block=block)
        else:
            # several processes on one queue: a request seen by empty() may be taken by
            # somebody else, so never block while a packet is in flight, and take one at a time
            import queue as _queue
            try:
                got = [requestQueue.get() if block else requestQueue.get(block=False)]
            except _queue.Empty:
                got = []
        if not got:
            return False
        now = time.time()
        pending.extend((r, now) for r in got)
        debugPrint(args, "Accel", "Pulled request")
        return True

    while True:
        if len(inflight) >= depth:
            finish_oldest()
            continue
        if not pending and not pull(block=not inflight):
            finish_oldest()                # nothing waiting: complete the oldest launch instead
            continue
        request, start_time = pending[0]

        if request is None:
            # With several processes on one queue the sentinel keeps circulating: whoever takes it
            # puts it back for the others (requests precede it in the FIFO, so the queue has room)
            if multi:
                requestQueue.put(None)
            while inflight:
                finish_oldest()
            if backend is not None:
                backend.close()
            if sibling:
                return                     # responses are flushed when the process exits
            for p in siblings:             # their responses must precede the final sentinel
                p.join(timeout=120)
            debugPrint(args, "Accel", "Sending final done signal")
            if resp_ring is not None and responder.is_alive():
                resp_ring.push([(0, 0, 0, 0, 0, 0, 0, KIND_SENTINEL, 0.0, 0.0, 0.0)])
                responder.join(timeout=120)    # it forwards the sentinel after the last response
            else:
                writer.put(None)
            return

        if backend is None:
            pending.popleft()
            respond_many([(request, start_time)], [0])
            continue

        # ---- coalesced launch: every plain request that is already here
        if coalesce > 1 and not _carries_inputs(request):
            group, sizes, rows = [], [], 0
            while pending and len(group) < coalesce:
                r, st = pending[0]
                if r is None or _carries_inputs(r):
                    break
                try:
                    n = int(r.batch_size)
                except (TypeError, ValueError):
                    break
                if n < 1 or (group and rows + n > stack_rows):
                    break
                group.append((r, st))
                sizes.append(n)
                rows += n
                pending.popleft()
            if len(group) > 1:
                try:
                    handle = backend.submit_many([int(r.batch_id) for r, _ in group], sizes)
                    inflight.append((group, handle, sizes))
                except Exception as exc:
                    print("[Accel] coalesced launch refused (%r); serving one by one" % (exc,))
                    sys.stdout.flush()
                    for entry in reversed(group):
                        pending.appendleft(entry)
                    coalesce = 1
                continue
            if group:
                pending.appendleft(group[0])   # a single request: the ordinary path below
                request, start_time = group[0]

        pending.popleft()
        try:
            if pipelined:
                inflight.append(([(request, start_time)],
                                 backend.submit_inputs(request) if _carries_inputs(request)
                                 else backend.submit(request.batch_id, request.batch_size), None))
                continue
            if _carries_inputs(request):
                out = backend.serve_inputs(request)
            else:
                out = backend.serve(request.batch_id, request.batch_size)
            out_batch_size = int(out.shape[0])
            dump(request, out)
        except Exception as exc:           # keep serving: the parent counts responses
            print("[Accel] request failed: %r" % (exc,))
            sys.stdout.flush()
            out_batch_size = 0
        respond_many([(request, start_time)], [out_batch_size])
