"""Host-side handle on the CUDA engine (thin: spec -> C-ABI -> kernels).

Plays the role of the reference's ``*_Wrapper`` + ``*_Net`` pair
(/root/reference/models/dlrm_s_caffe2.py:79-211 ``DLRM_Wrapper``,
:213-569 ``DLRM_Net``) for the accel engine: ``create`` loads parameters,
``run_queues(ids, lengths, fc, batch_size)`` / ``run(X, S_lengths, S_indices)``
serve one query.  numpy arrays are the host-side containers; all arithmetic is
in libdrs_b200.so.
"""
from __future__ import absolute_import, division, print_function

import ctypes as C

import numpy as np

from . import _lib
from .spec import ModelSpec, iter_parameters


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _ptr_array(arrays):
    arr = (C.c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data
    return arr


class AccelEngine(object):
    def __init__(self, args=None, spec=None, device=0, max_batch=None, precision=None,
                 lanes=0, max_lookups_per_table=None, shard_mode="table", shard_world=1):
        self.lib = _lib.load()
        self.spec = spec if spec is not None else ModelSpec(args)
        s = self.spec
        if max_batch is None:
            max_batch = int(getattr(args, "max_mini_batch_size", 1) or 1)
        if precision is None:
            precision = getattr(args, "accel_precision", "fp32") if args is not None else "fp32"
        self.max_batch = int(max_batch)
        self.precision = precision
        self._rows = (C.c_int64 * s.T)(*s.table_rows)
        self._bot = (C.c_int32 * max(1, len(s.ln_bot)))(*s.ln_bot)
        self._top = (C.c_int32 * len(s.ln_top))(*s.ln_top)
        cs = _lib.Spec()
        cs.model_type = _lib.MODEL_TYPES[s.model_type]
        cs.n_tables = s.T
        cs.table_rows = self._rows
        cs.emb_dim = s.D
        cs.n_bot = len(s.ln_bot)
        cs.ln_bot = self._bot
        cs.n_top = len(s.ln_top)
        cs.ln_top = self._top
        cs.interaction_op = _lib.INTERACT_DOT if s.interaction_op == "dot" else _lib.INTERACT_CAT
        cs.interaction_itself = int(s.interaction_itself)
        cs.max_batch = self.max_batch
        cs.max_lookups_per_table = int(max_lookups_per_table or self.max_batch * max(1, s.L))
        self._max_lookups = cs.max_lookups_per_table
        cs.precision = _lib.PRECISIONS[precision]
        cs.n_lanes = int(lanes)
        cs.hidden_size = int(getattr(s, "hidden_size", 0))
        if shard_mode not in ("table", "row"):
            raise ValueError("shard_mode must be 'table' or 'row'")
        self.shard_mode = shard_mode
        cs.shard_mode = _lib.SHARD_ROW if shard_mode == "row" else _lib.SHARD_TABLE
        cs.shard_world = int(shard_world) if shard_mode == "row" else 0
        if s.model_type == "mtwnd":
            self._task = (C.c_int32 * len(s.ln_task))(*s.ln_task)
            cs.n_tasks, cs.n_task, cs.ln_task = s.n_tasks, len(s.ln_task), self._task
            cs.task_sigmoid_layer = s.task_sigmoid_layer
        self._handle = C.c_void_p()
        _lib.check(self.lib.drs_create(C.byref(cs), int(device), C.byref(self._handle)))
        self.out_dim = s.out_dim
        self._keep = []           # host arrays referenced by async calls

    # ------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.drs_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------ parameters
    def load_table(self, t, W):
        W = np.ascontiguousarray(W, dtype=np.float32)
        _lib.check(self.lib.drs_load_table(self._handle, int(t), _ptr(W), W.shape[0], W.shape[1]))

    def load_fc(self, mlp_id, layer, W, b):
        W = np.ascontiguousarray(W, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        _lib.check(self.lib.drs_load_fc(self._handle, int(mlp_id), int(layer), _ptr(W), _ptr(b),
                                        W.shape[0], W.shape[1]))

    def create(self, on_parameter=None):
        """Draw every parameter from numpy's global RNG in the reference's order
        and load it (mirror of ``DLRM_Net.create_sequential_forward_ops``)."""
        for item in iter_parameters(self.spec):
            if item[0] == "table":
                self.load_table(item[1], item[2])
            else:
                self.load_fc(item[1], item[2], item[3], item[4])
            if on_parameter is not None:
                on_parameter(item)

    def load_parameters(self, tables, fcs):
        for t, W in enumerate(tables):
            self.load_table(t, W)
        for (mlp_id, layer), (W, b) in sorted(fcs.items()):
            self.load_fc(mlp_id, layer, W, b)

    # ------------------------------------------------------------ inputs
    def _prep(self, dense, ids, lengths, batch_size):
        s = self.spec
        B = int(batch_size)
        ids = [np.ascontiguousarray(a) for a in ids]
        is64 = ids[0].dtype == np.int64
        want = np.int64 if is64 else np.int32
        ids = [a if a.dtype == want else a.astype(want) for a in ids]
        lengths = [np.ascontiguousarray(l, dtype=np.int32) for l in lengths]
        if len(ids) != s.T or len(lengths) != s.T:
            raise ValueError("expected %d tables, got %d ids / %d lengths"
                             % (s.T, len(ids), len(lengths)))
        for t in range(s.T):
            if lengths[t].shape[0] < B:
                raise ValueError("table %d: %d lengths for batch %d" % (t, lengths[t].shape[0], B))
            need = int(lengths[t][:B].sum())
            if ids[t].shape[0] < need:
                # Caffe2 SparseLengthsSum enforce: sum(lengths) == len(indices)
                raise _lib.DrsError(_lib.ERR_INDEX, "table %d: sum(lengths)=%d > len(indices)=%d"
                                    % (t, need, ids[t].shape[0]))
        if s.m_den > 0:
            dense = np.ascontiguousarray(dense, dtype=np.float32)
            if dense.shape[0] < B or dense.shape[1] != s.m_den:
                raise ValueError("dense must be [>=%d, %d]" % (B, s.m_den))
            dptr = _ptr(dense)
        else:
            dense, dptr = None, None
        return dense, dptr, ids, int(is64), lengths

    def stage_batch(self, batch_id, dense, ids, lengths, n_samples=None):
        n = int(n_samples if n_samples is not None else lengths[0].shape[0])
        dense, dptr, ids, is64, lengths = self._prep(dense, ids, lengths, n)
        _lib.check(self.lib.drs_stage_batch(self._handle, int(batch_id), dptr, _ptr_array(ids),
                                            is64, _ptr_array(lengths), n))

    # ------------------------------------------------------------ serving
    def infer_staged(self, batch_id, batch_size, want_output=True):
        out = np.empty((int(batch_size), self.out_dim), dtype=np.float32) if want_output else None
        tm = _lib.Timing()
        _lib.check(self.lib.drs_infer_staged(self._handle, int(batch_id), int(batch_size),
                                             _ptr(out) if want_output else None, C.byref(tm)))
        return out, tm

    def infer(self, dense, ids, lengths, batch_size):
        dense, dptr, ids, is64, lengths = self._prep(dense, ids, lengths, batch_size)
        out = np.empty((int(batch_size), self.out_dim), dtype=np.float32)
        tm = _lib.Timing()
        _lib.check(self.lib.drs_infer(self._handle, dptr, _ptr_array(ids), is64,
                                      _ptr_array(lengths), int(batch_size), _ptr(out),
                                      C.byref(tm)))
        return out, tm

    # reference-shaped entry points ------------------------------------------
    def run_queues(self, ids, lengths, fc, batch_size):
        """``DLRM_Wrapper.run_queues(ids, lengths, fc, batch_size)``
        (dlrm_s_caffe2.py:162-174): ids [T, B*L], lengths [T, B], fc [B, m_den]."""
        out, _ = self.infer(fc, list(ids), list(lengths), batch_size)
        return out

    def run(self, X, S_lengths, S_indices):
        """``DLRM_Net.run(X, S_lengths, S_indices)`` (dlrm_s_caffe2.py:549-569)."""
        lengths = [np.asarray(l, dtype=np.int32) for l in S_lengths]
        ids = [np.asarray(i) for i in S_indices]
        out, _ = self.infer(X, ids, lengths, lengths[0].shape[0])
        return out

    # async halves ------------------------------------------------------------
    def enqueue_staged(self, batch_id, batch_size):
        _lib.check(self.lib.drs_enqueue_staged(self._handle, int(batch_id), int(batch_size)))

    def enqueue_host(self, dense_ptr, ids_ptrs, is64, len_ptrs, batch_size, out_ptr):
        """Raw-pointer variant for the bench/serving loops (no per-call numpy work)."""
        _lib.check(self.lib.drs_enqueue_host(self._handle, dense_ptr, ids_ptrs, is64, len_ptrs,
                                             int(batch_size), out_ptr))

    def enqueue_host_packed(self, dense_ptr, packed_ptrs, bits, len_ptrs, batch_size, out_ptr):
        """``enqueue_host`` with ids in the packed wire format (``pack_ids``): ``bits`` is a
        ctypes int32 array, one width per table."""
        _lib.check(self.lib.drs_enqueue_host_packed(self._handle, dense_ptr, packed_ptrs, bits,
                                                    len_ptrs, int(batch_size), out_ptr))

    def submit_staged(self, batch_id, batch_size, out_ptr=None):
        """Asynchronous serve of a staged batch; returns a ticket for ``wait``."""
        t = C.c_int64()
        _lib.check(self.lib.drs_submit_staged(self._handle, int(batch_id), int(batch_size),
                                              out_ptr, C.byref(t)))
        return t.value

    def submit_host(self, dense_ptr, ids_ptrs, is64, len_ptrs, batch_size, out_ptr=None):
        """Asynchronous serve of a query held in (pinned) host buffers: H2D, kernels and D2H are
        enqueued on the next lane; returns a ticket for ``wait``.  Raw pointers, no numpy work."""
        t = C.c_int64()
        _lib.check(self.lib.drs_submit_host(self._handle, dense_ptr, ids_ptrs, int(is64), len_ptrs,
                                            int(batch_size), out_ptr, C.byref(t)))
        return t.value

    MAX_COALESCE = 32

    def submit_many(self, batch_ids, batch_sizes, out_ptr=None):
        """Coalesced serve of up to 32 queued requests (request k = first ``batch_sizes[k]``
        samples of staged batch ``batch_ids[k]``) as ONE stacked query: one gather launch, one
        pass of the MLPs, one D2H copy.  Row r of request k lands at row
        ``sum(batch_sizes[:k]) + r`` of the output.  Returns a ticket for ``wait``."""
        n = len(batch_ids)
        if isinstance(batch_ids, np.ndarray) and isinstance(batch_sizes, np.ndarray):
            # the record loop hands int32 arrays: no per-request Python work
            keep = (np.ascontiguousarray(batch_ids, dtype=np.int32),
                    np.ascontiguousarray(batch_sizes, dtype=np.int32))
            ids, sizes = (a.ctypes.data_as(C.POINTER(C.c_int32)) for a in keep)
        else:
            ids = (C.c_int32 * n)(*batch_ids)
            sizes = (C.c_int32 * n)(*batch_sizes)
        t = C.c_int64()
        _lib.check(self.lib.drs_submit_many(self._handle, n, ids, sizes, out_ptr, C.byref(t)))
        return t.value

    def wait(self, ticket):
        _lib.check(self.lib.drs_wait(self._handle, int(ticket)))

    def lane_count(self):
        n = C.c_int()
        _lib.check(self.lib.drs_lane_count(self._handle, C.byref(n)))
        return n.value

    def sync(self):
        _lib.check(self.lib.drs_sync(self._handle))

    def timer_start(self):
        _lib.check(self.lib.drs_timer_start(self._handle))

    def timer_stop(self):
        ms = C.c_float()
        _lib.check(self.lib.drs_timer_stop(self._handle, C.byref(ms)))
        return ms.value

    # operator + taps -----------------------------------------------------------
    def sparse_lengths_sum(self, table, indices, lengths):
        """SparseLengthsSum([table, indices, lengths]) on a loaded table."""
        indices = np.ascontiguousarray(indices)
        if indices.dtype not in (np.int32, np.int64):
            indices = indices.astype(np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        out = np.empty((lengths.shape[0], self.spec.D), dtype=np.float32)
        _lib.check(self.lib.drs_sparse_lengths_sum(
            self._handle, int(table), _ptr(indices), int(indices.dtype == np.int64),
            indices.shape[0], _ptr(lengths), lengths.shape[0], _ptr(out)))
        return out

    def fetch_stage(self, stage, layer=0):
        cap = self.max_batch * max(self.spec.T * self.spec.D, max(self.spec.ln_top),
                                   max(self.spec.ln_bot or [1]), self.max_batch + 1) + 16
        if stage == _lib.STAGE_IDS:
            cap = max(cap, int(self._max_lookups) + 16)
        buf = np.empty(cap, dtype=np.float32)
        rows, cols = C.c_int(), C.c_int()
        _lib.check(self.lib.drs_fetch_stage(self._handle, int(stage), int(layer), _ptr(buf), cap,
                                            C.byref(rows), C.byref(cols)))
        n = rows.value * cols.value
        view = buf[:n].reshape(rows.value, cols.value)
        if stage in (_lib.STAGE_OFFSETS, _lib.STAGE_IDS):
            return view.view(np.int32).copy()
        return view.copy()

    def set_option(self, key, value):
        _lib.check(self.lib.drs_set_option(self._handle, key.encode(), int(value)))

    def profile_begin(self, max_steps):
        _lib.check(self.lib.drs_profile_begin(self._handle, int(max_steps)))

    def profile_end(self):
        """-> (n_steps, pre_gather_ms, gather_ms, post_gather_ms), summed over the steps."""
        n, a, b, c = C.c_int(), C.c_float(), C.c_float(), C.c_float()
        _lib.check(self.lib.drs_profile_end(self._handle, C.byref(n), C.byref(a), C.byref(b),
                                            C.byref(c)))
        return n.value, a.value, b.value, c.value

    # multi-GPU, one process per GPU ------------------------------------------
    def shard_export(self):
        """64-byte CUDA-IPC handle + layout of this rank's exchange slab."""
        blob = (C.c_ubyte * _lib.SHARD_BLOB_BYTES)()
        _lib.check(self.lib.drs_shard_export(self._handle, blob))
        return bytes(blob)

    def shard_connect(self, rank, world, blobs):
        """``blobs``: the exported blob of every rank, in rank order."""
        assert len(blobs) == world
        buf = b"".join(blobs)
        arr = (C.c_ubyte * len(buf)).from_buffer_copy(buf)
        _lib.check(self.lib.drs_shard_connect(self._handle, int(rank), int(world), arr))
        self.rank, self.world = int(rank), int(world)

    def exchange_begin(self, batch_id, batch_size):
        """Unfused-exchange baseline, first half (``set_option("exchange", 1)``):
        -> (send address, recv address, floats per peer, cudaStream_t)."""
        send, recv, stream, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self.lib.drs_exchange_begin(self._handle, int(batch_id), int(batch_size),
                                               C.byref(send), C.byref(recv), C.byref(n),
                                               C.byref(stream)))
        return send.value, recv.value, n.value, stream.value

    def exchange_end(self):
        _lib.check(self.lib.drs_exchange_end(self._handle))

    def launch_count(self):
        n = C.c_int64()
        _lib.check(self.lib.drs_launch_count(self._handle, C.byref(n)))
        return n.value

    def hbm_bytes(self):
        n = C.c_int64()
        _lib.check(self.lib.drs_hbm_bytes(self._handle, C.byref(n)))
        return n.value


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_device_numa_node(device):
    """Pin this process to the CPUs of the NUMA node GPU ``device`` is attached to, so that the
    pinned staging buffers it allocates afterwards (first touch) and the threads that fill them
    sit next to that GPU's PCIe root port.  With one serving process per GPU this keeps H2D
    copies off the inter-socket link.  Returns the node, or -1 if nothing was changed."""
    import os
    node = C.c_int(-1)
    _lib.check(_lib.load().drs_device_numa_node(int(device), C.byref(node)))
    if node.value < 0:
        return -1
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node.value) as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return -1
        os.sched_setaffinity(0, cpus)
    except (OSError, ValueError):
        return -1
    return node.value


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc'ed memory (kept alive by the returned owner)."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _lib.check(lib.drs_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p


def id_bits(n_rows):
    """Bits an id of an ``n_rows``-row table needs in the packed wire format."""
    return max(1, int(n_rows - 1).bit_length())


def pack_ids(ids, bits, out=None):
    """int32 ids -> packed uint32 stream (``drs_pack_ids``).  ``out``: optional preallocated
    (pinned) uint32 array of at least ``drs_packed_words(len(ids), bits)`` words."""
    lib = _lib.load()
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    words = int(lib.drs_packed_words(ids.shape[0], int(bits)))
    if out is None:
        out = np.empty(words, dtype=np.uint32)
    elif out.shape[0] < words or out.dtype != np.uint32:
        raise ValueError("out must hold %d uint32 words" % words)
    _lib.check(lib.drs_pack_ids(_ptr(ids), ids.shape[0], int(bits), _ptr(out)))
    return out[:words]


def hostpack_ids(ids, bits, threads=1):
    """The engine's own host packer (``drs_hostpack_ids``; lane-interleaved 512-id blocks,
    csrc/hostpack.h) on int32 or int64 ids -> (uint32 words, ok).  For tests and timing: the
    serving path packs inside ``drs_enqueue_host`` once ``set_option("host_pack", n)`` is on."""
    lib = _lib.load()
    ids = np.ascontiguousarray(ids)
    if ids.dtype not in (np.int32, np.int64):
        ids = ids.astype(np.int64)
    out = np.empty(int(lib.drs_hostpack_words(ids.shape[0], int(bits))), dtype=np.uint32)
    ok = C.c_int(1)
    _lib.check(lib.drs_hostpack_ids(_ptr(ids), int(ids.dtype == np.int64), ids.shape[0], int(bits),
                                    int(threads), _ptr(out), C.byref(ok)))
    return out, bool(ok.value)
