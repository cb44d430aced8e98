"""Packed id wire format (drs_pack_ids / drs_enqueue_host_packed): the upload path with
ceil(log2(rows))-bit ids gives bit-for-bit what the int32 upload path gives."""
import ctypes as C

import numpy as np
import pytest

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import AccelEngine, id_bits, pack_ids, pinned_empty
from tests.helpers import make_case, request

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["dlrm_cat", "wnd", "din_many", "dlrm_dot_itself"])
def test_packed_upload_equals_int32_upload(name):
    c = make_case(name, num_batches=1, max_batch=64)
    T = c.spec.T
    with AccelEngine(args=c.args, max_batch=64) as eng:
        eng.load_parameters(c.tables, c.fcs)
        for B in (64, 17, 1):
            dense, ids, lengths = request(c, 0, B)
            want, _ = eng.infer(dense, ids, lengths, B)
            want_pooled = eng.fetch_stage(_lib.STAGE_POOLED)
            bits = [id_bits(r) for r in c.spec.table_rows]
            packed = [pack_ids(ids[t], bits[t]) for t in range(T)]
            out, _o = pinned_empty((64, c.spec.out_dim), np.float32)
            out[:] = np.nan
            dptr = C.c_void_p(dense.ctypes.data) if dense is not None else None
            eng.enqueue_host_packed(
                dptr, (C.c_void_p * T)(*[p.ctypes.data for p in packed]),
                (C.c_int32 * T)(*bits),
                (C.c_void_p * T)(*[l.ctypes.data for l in lengths]), B, C.c_void_p(out.ctypes.data))
            eng.sync()
            assert np.array_equal(np.array(out[:B]).view(np.uint32), want.view(np.uint32))
            assert np.array_equal(eng.fetch_stage(_lib.STAGE_POOLED).view(np.uint32),
                                  want_pooled.view(np.uint32))
            assert sum(p.nbytes for p in packed) < sum(i.astype(np.int32).nbytes for i in ids) \
                or B == 1
            # equally spaced streams in one [T, words] buffer: one copy + one widening kernel
            if len({len(i) for i in ids}) == 1 and len(set(bits)) == 1 and T > 1:
                slab, _s = pinned_empty((T, max(len(p) for p in packed)), np.uint32)
                for t in range(T):
                    slab[t, :len(packed[t])] = packed[t]
                out[:] = np.nan
                eng.enqueue_host_packed(
                    dptr, (C.c_void_p * T)(*[slab[t].ctypes.data for t in range(T)]),
                    (C.c_int32 * T)(*bits),
                    (C.c_void_p * T)(*[l.ctypes.data for l in lengths]), B,
                    C.c_void_p(out.ctypes.data))
                eng.sync()
                assert np.array_equal(np.array(out[:B]).view(np.uint32), want.view(np.uint32))


def test_packed_ids_are_validated():
    c = make_case("dlrm_cat", num_batches=1, max_batch=64)
    T = c.spec.T
    with pytest.raises(_lib.DrsError):
        pack_ids(np.array([5, 70000], np.int32), 16)            # 70000 needs 17 bits
    with AccelEngine(args=c.args, max_batch=64) as eng:
        eng.load_parameters(c.tables, c.fcs)
        dense, ids, lengths = request(c, 0, 8)
        packed = [pack_ids(ids[t], 8) for t in range(T)] if max(int(i.max()) for i in ids) < 256 \
            else [pack_ids(ids[t] & 255, 8) for t in range(T)]
        with pytest.raises(_lib.DrsError):                      # 8 bits cannot address 5000+ rows
            eng.enqueue_host_packed(
                C.c_void_p(dense.ctypes.data), (C.c_void_p * T)(*[p.ctypes.data for p in packed]),
                (C.c_int32 * T)(*([8] * T)), (C.c_void_p * T)(*[l.ctypes.data for l in lengths]),
                8, None)
