"""drs_set_option(e, "host_pack", n): the host path packs the caller's int32 / int64 ids on n
host threads (csrc/hostpack.cpp) and widens them on the GPU (k_unpack_ids_blk) — results are
bit-for-bit those of the plain upload, and the SparseLengthsSum enforce still fires."""
import numpy as np
import pytest

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import AccelEngine
from tests.helpers import make_case, request

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["dlrm_cat", "dlrm_dot", "wnd", "din_many"])
def test_host_packed_upload_equals_plain_upload(name):
    c = make_case(name, num_batches=1, max_batch=64)
    with AccelEngine(args=c.args, max_batch=64) as eng:
        eng.load_parameters(c.tables, c.fcs)
        for B in (64, 17, 1):
            dense, ids, lengths = request(c, 0, B)
            eng.set_option("host_pack", 0)
            want, _ = eng.infer(dense, ids, lengths, B)
            want_pooled = eng.fetch_stage(_lib.STAGE_POOLED)
            want_ids = [eng.fetch_stage(_lib.STAGE_IDS, t) for t in range(c.spec.T)]
            eng.set_option("host_pack", 3)
            eng.set_option("host_pack_min_ids", 0)
            for wide in (False, True):
                cast = [i.astype(np.int64 if wide else np.int32) for i in ids]
                l0 = eng.launch_count()
                got, _ = eng.infer(dense, cast, lengths, B)
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
                assert np.array_equal(eng.fetch_stage(_lib.STAGE_POOLED).view(np.uint32),
                                      want_pooled.view(np.uint32))
                for t in range(c.spec.T):
                    assert np.array_equal(eng.fetch_stage(_lib.STAGE_IDS, t), want_ids[t])
                assert eng.launch_count() > l0


def test_host_packed_upload_at_a_size_that_spans_chunks_and_a_ragged_tail():
    c = make_case("dlrm_cat", num_batches=1, max_batch=2048)
    with AccelEngine(args=c.args, max_batch=2048) as eng:
        eng.load_parameters(c.tables, c.fcs)
        B = 2039
        dense, ids, lengths = request(c, 0, B)
        want, _ = eng.infer(dense, ids, lengths, B)
        eng.set_option("host_pack", 4)
        eng.set_option("host_pack_min_ids", 0)
        for _ in range(3):                              # every lane's staging gets used
            got, _ = eng.infer(dense, [i.astype(np.int64) for i in ids], lengths, B)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("bad", [-1, 10 ** 9])
def test_enforce_fires_through_the_packed_path(bad):
    c = make_case("dlrm_cat", num_batches=1, max_batch=64)
    with AccelEngine(args=c.args, max_batch=64) as eng:
        eng.load_parameters(c.tables, c.fcs)
        eng.set_option("host_pack", 2)
        eng.set_option("host_pack_min_ids", 0)
        dense, ids, lengths = request(c, 0, 32)
        ids = [i.copy() for i in ids]
        ids[1][5] = bad
        with pytest.raises(_lib.DrsError) as ei:
            eng.infer(dense, ids, lengths, 32)
        assert ei.value.code == _lib.ERR_INDEX
        dense, ids, lengths = request(c, 0, 32)         # and the engine keeps serving
        out, _ = eng.infer(dense, ids, lengths, 32)
        assert np.isfinite(out).all()
