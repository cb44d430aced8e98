"""Requests that carry their own tensors: the shared-memory input ring, the
TensorServiceRequest packet and their way through the accel engine's Queue protocol."""
import glob
import multiprocessing as mp
import os
import pickle

import numpy as np
import pytest

from deeprecsys_b200 import accelInferenceEngine as accel_mod
from deeprecsys_b200.utils.input_ring import SharedInputRing, slots_needed
from deeprecsys_b200.utils.packets import ServiceRequest, TensorServiceRequest
from deeprecsys_b200.utils.utils import cli

SMALL_SYSTEM = ["--arch_embedding_size", "500-600-700-800", "--arch_sparse_feature_size", "16",
                "--arch_mlp_bot", "13-24-16", "--arch_mlp_top", "32-8-1",
                "--arch_interaction_op", "cat", "--num_indices_per_lookup", "10",
                "--model_name", "rm1", "--num_batches", "2", "--max_mini_batch_size", "64"]
ROWS = (500, 600, 700, 800)


def _query(rng, B, ragged=False):
    dense = rng.rand(B, 13).astype(np.float32)
    if ragged:
        lengths = [rng.randint(0, 11, size=B).astype(np.int32) for _ in ROWS]
    else:
        lengths = [np.full(B, 10, np.int32) for _ in ROWS]
    ids = [rng.randint(0, rows, size=int(l.sum())).astype(np.int64)
           for rows, l in zip(ROWS, lengths)]
    return dense, ids, lengths


def test_ring_layout_write_read_and_pointers():
    ring = SharedInputRing.create(n_slots=3, n_tables=4, m_den=13, max_batch=64, max_lookups=640)
    try:
        assert os.path.exists(ring.path) and os.path.getsize(ring.path) == ring.nbytes
        assert ring.slot_bytes % 64 == 0
        rng = np.random.RandomState(0)
        queries = [_query(rng, B, ragged=(B % 2 == 1)) for B in (64, 7, 1)]
        counts = [ring.write(s, d, i, l, l[0].shape[0]) for s, (d, i, l) in enumerate(queries)]
        for s, (d, i, l) in enumerate(queries):
            B = l[0].shape[0]
            assert counts[s] == [int(x.sum()) for x in l]
            rd, ri, rl = ring.read(s, B, counts[s])
            assert np.array_equal(rd, d)
            for t in range(4):
                assert ri[t].dtype == np.int32 and np.array_equal(ri[t], i[t])
                assert np.array_equal(rl[t], l[t])
            # counts are optional: recomputed from the lengths
            assert [len(x) for x in ring.read(s, B)[1]] == counts[s]
            # the raw addresses the C-ABI receives are the addresses of those views
            dp, ip, lp = ring.pointers(s)
            assert dp == rd.ctypes.data
            assert ip == [ring.ids(s)[t].ctypes.data for t in range(4)]
            assert lp == [ring.lengths(s)[t].ctypes.data for t in range(4)]
            assert all(p % 4 == 0 for p in ip + lp) and dp % 64 == ring.base_address() % 64
        with pytest.raises(IndexError):
            ring.dense(3)
        d, i, l = _query(rng, 64)
        with pytest.raises(ValueError):
            ring.write(0, d, i, l, 65)                       # more samples than a slot holds
        with pytest.raises(ValueError):
            ring.write(0, d, [x[:5] for x in i], l, 64)      # sum(lengths) > len(indices)
        big = [np.full(64, 11, np.int32)] * 4
        with pytest.raises(ValueError):
            ring.write(0, d, [np.zeros(704, np.int64)] * 4, big, 64)   # 704 lookups > 640
    finally:
        path = ring.path
        ring.close()
    assert not os.path.exists(path)


def test_ring_slots_are_handed_back_explicitly():
    """Slot reuse does not rely on counting what the queue and the engine may hold: a slot is
    busy from ``acquire`` until the consumer ``release``s it, in whichever process."""
    ring = SharedInputRing.create(n_slots=2, n_tables=4, m_den=13, max_batch=16, max_lookups=160)
    try:
        assert ring.base_address() % 64 == 0 and ring.header_bytes % 64 == 0
        ring.acquire(0)
        assert ring.is_busy(0) and not ring.is_busy(1)
        with pytest.raises(TimeoutError):
            ring.acquire(0, timeout_s=0.05)                  # still owned by the consumer
        other = SharedInputRing.attach(ring.descriptor())    # the flags live in the mapping
        assert other.is_busy(0)
        other.release(0)
        other.close()
        ring.acquire(0, timeout_s=0.05)
        for bad in (-1, 2):
            with pytest.raises(IndexError):
                ring.pointers(bad)
            with pytest.raises(IndexError):
                ring.release(bad)
        assert slots_needed(32, 3) == 37 and slots_needed(32, 3, consumers=8) == 65
    finally:
        ring.close()


def test_ring_is_shared_between_processes():
    ring = SharedInputRing.create(n_slots=2, n_tables=4, m_den=13, max_batch=16, max_lookups=160)
    try:
        d, i, l = _query(np.random.RandomState(1), 16)
        counts = ring.write(1, d, i, l, 16)
        desc = pickle.loads(pickle.dumps(ring.descriptor()))

        def child(q):
            r = SharedInputRing.attach(desc)
            rd, ri, rl = r.read(1, 16, counts)
            q.put((rd.copy(), [x.copy() for x in ri], [x.copy() for x in rl]))
            r.dense(0)[:] = 7.0                              # visible to the creator
            r.close()

        ctx = mp.get_context("fork")
        q = ctx.Queue()
        p = ctx.Process(target=child, args=(q,))
        p.start()
        rd, ri, rl = q.get(timeout=30)
        p.join(timeout=30)
        assert np.array_equal(rd, d) and all(np.array_equal(a, b) for a, b in zip(ri, i))
        assert all(np.array_equal(a, b) for a, b in zip(rl, l))
        assert (ring.dense(0) == 7.0).all()
        assert os.path.exists(ring.path)                     # attach/close never unlinks
        with pytest.raises(ValueError):                      # a descriptor larger than the file
            SharedInputRing.attach(dict(desc, n_slots=5))
    finally:
        ring.close()


def test_tensor_request_keeps_the_reference_fields_and_stays_small():
    ref = ServiceRequest(batch_id=3, epoch=1, batch_size=17, sub_id=0, total_sub_batches=1,
                         exp_packet=False, arrival_time=1.5)
    ring_desc = {"path": "/dev/shm/x", "n_slots": 40, "n_tables": 8, "m_den": 128,
                 "max_batch": 1024, "max_lookups": 81920}
    req = TensorServiceRequest(ring=ring_desc, slot=5, counts=[1360] * 8, batch_id=3, epoch=1,
                               batch_size=17, sub_id=0, total_sub_batches=1, exp_packet=False,
                               arrival_time=1.5)
    for k, v in vars(ref).items():
        assert getattr(req, k) == v
    assert req.has_inputs() and not TensorServiceRequest(batch_id=0).has_inputs()
    assert len(pickle.dumps(req)) < 700                      # vs ~0.8 MB of arrays for RMC1, B=256
    assert slots_needed(32, 3) == 37


ENGINE_SCRIPT = r"""
import glob, multiprocessing as mp, os, pickle, sys
import numpy as np
from deeprecsys_b200 import accelInferenceEngine as accel_mod
from deeprecsys_b200.utils.input_ring import SharedInputRing
from deeprecsys_b200.utils.packets import ServiceRequest, TensorServiceRequest
from deeprecsys_b200.utils.utils import cli
from oracle.cpu_inference_engine import OracleBackend        # CPU test tier only: no GPU here
from tests.test_input_ring import SMALL_SYSTEM, _query

out_dir = sys.argv[1]
rng = np.random.RandomState(42)
ring = SharedInputRing.create(n_slots=4, n_tables=4, m_den=13, max_batch=64, max_lookups=640)
fields = dict(sub_id=0, total_sub_batches=1, exp_packet=False, arrival_time=0.0)
queries, requests = {}, []
for n, (B, ragged) in enumerate([(64, False), (9, True), (1, False)]):
    d, i, l = _query(rng, B, ragged)
    counts = ring.write(n, d, i, l, B)
    queries[(10 + n, 0, B)] = (d, i, l)
    requests.append(TensorServiceRequest(ring=ring.descriptor(), slot=n, counts=counts,
                                         batch_id=0, epoch=10 + n, batch_size=B, **fields))
d, i, l = _query(rng, 5, True)
queries[(20, 1, 5)] = (d, i, l)
requests.append(TensorServiceRequest(dense=d, ids=i, lengths=l, batch_id=1, epoch=20,
                                     batch_size=5, **fields))
requests.append(ServiceRequest(batch_id=1, epoch=30, batch_size=6, **fields))
requests.append(TensorServiceRequest(ring=dict(ring.descriptor(), path=ring.path + ".gone"),
                                     slot=0, counts=[0] * 4, batch_id=0, epoch=40, batch_size=3,
                                     **fields))
os.environ["DRS_ACCEL_DUMP_DIR"] = out_dir
accel_mod.BACKEND_FACTORY = OracleBackend
ctx = mp.get_context("fork")
rq, resp, ready = ctx.Queue(), ctx.Queue(), ctx.Queue()
p = ctx.Process(target=accel_mod.accelInferenceEngine, args=(cli(SMALL_SYSTEM), rq, 1, resp, ready))
p.daemon = True
p.start()
assert ready.get(timeout=120) is True
for r in requests:
    rq.put(r)
rq.put(None)
responses = []
while True:
    item = resp.get(timeout=120)
    if item is None:
        break
    responses.append((item.epoch, item.batch_size, item.out_batch_size))
p.join(timeout=30)
ring.close()
with open(os.path.join(out_dir, "run.pkl"), "wb") as f:
    pickle.dump({"responses": responses, "queries": queries,
                 "sent": [(q.epoch, q.batch_size) for q in requests]}, f)
"""


def test_engine_serves_requests_that_carry_inputs(tmp_path):
    """Ring-backed, inline and plain (batch_id) requests through one engine process: each
    response answers its request, and the outputs are what the oracle computes for THOSE
    inputs (which differ from the pre-generated dataset).  Run in a fresh interpreter: the
    engine process is forked, and a child forked from this (multi-threaded, torch-using) pytest
    process can deadlock in torch's thread pool."""
    import subprocess
    import sys
    from deeprecsys_b200.spec import ModelSpec, draw_all_parameters
    from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
    from oracle import numpy_oracle as O
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, "-c", ENGINE_SCRIPT, str(tmp_path)], cwd=root,
                          env=dict(os.environ, PYTHONPATH=root, PYTHONWARNINGS="ignore"),
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-3000:]
    with open(str(tmp_path / "run.pkl"), "rb") as f:
        run = pickle.load(f)
    dumps = {os.path.basename(f): np.load(f) for f in glob.glob(str(tmp_path / "*.npy"))}
    assert [r[:2] for r in run["responses"]] == run["sent"]
    assert [r[2] for r in run["responses"]] == [64, 9, 1, 5, 6, 0]    # bad ring: answered, 0
    # the engine process drew the dataset and the parameters from the seed; redo it here
    args = cli(SMALL_SYSTEM)
    gen = DLRMDataGenerator(args)
    _, lX, lS_l, lS_i = gen.generate_input_data()
    gen.generate_output_data()
    spec = ModelSpec(args)
    tables, fcs = draw_all_parameters(spec)
    for (epoch, batch_id, B), (d, i, l) in run["queries"].items():
        got = dumps["accel_e%d_b%d_n%d.npy" % (epoch, batch_id, B)]
        want = O.forward(spec, tables, fcs, d, i, l)["out"]
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    plain = O.forward(spec, tables, fcs, lX[1][:6], [x[:60] for x in lS_i[1]],
                      [x[:6] for x in lS_l[1]])["out"]
    np.testing.assert_allclose(dumps["accel_e30_b1_n6.npy"], plain, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_gpu_backend_serves_ring_and_inline_requests(tmp_path):
    """The real backend: slots of a page-locked shared-memory ring and inline arrays through
    drs_submit_host, against the oracle on the same inputs (outputs atol 1e-5)."""
    from oracle import numpy_oracle as O
    from deeprecsys_b200.spec import draw_all_parameters
    args = cli(SMALL_SYSTEM)
    np.random.seed(args.numpy_rand_seed)
    backend = accel_mod.AccelBackend(args)
    # same seed -> same draws as the backend made: datagen inputs, targets, then parameters
    np.random.seed(args.numpy_rand_seed)
    from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
    gen = DLRMDataGenerator(args)
    gen.generate_input_data()
    gen.generate_output_data()
    tables, fcs = draw_all_parameters(backend.spec)
    rng = np.random.RandomState(3)
    n_slots = slots_needed(2, backend.depth)
    ring = SharedInputRing.create(n_slots=n_slots, n_tables=4, m_den=13, max_batch=64,
                                  max_lookups=640)
    try:
        fields = dict(batch_id=0, epoch=0, sub_id=0, total_sub_batches=1, exp_packet=False)
        inflight = []
        for n in range(3 * n_slots):                         # slots are reused
            B = (64, 9, 1, 33)[n % 4]
            d, i, l = _query(rng, B, ragged=(n % 2 == 1))
            want = O.forward(backend.spec, tables, fcs, d, i, l)["out"]
            if n % 5 == 4:
                req = TensorServiceRequest(dense=d, ids=i, lengths=l, batch_size=B, **fields)
            else:
                slot = n % n_slots
                counts = ring.write(slot, d, i, l, B)
                req = TensorServiceRequest(ring=ring.descriptor(), slot=slot, counts=counts,
                                           batch_size=B, **fields)
            inflight.append((backend.submit_inputs(req), want))
            if len(inflight) >= backend.depth:
                handle, ref = inflight.pop(0)
                np.testing.assert_allclose(backend.complete(handle), ref, rtol=1e-4, atol=1e-5)
        for handle, ref in inflight:
            np.testing.assert_allclose(backend.complete(handle), ref, rtol=1e-4, atol=1e-5)
        d, i, l = _query(rng, 12)
        ring.write(0, d, i, l, 12)
        out = backend.serve_inputs(TensorServiceRequest(ring=ring.descriptor(), slot=0,
                                                        batch_size=12, **fields))
        np.testing.assert_allclose(out, O.forward(backend.spec, tables, fcs, d, i, l)["out"],
                                   rtol=1e-4, atol=1e-5)
        assert backend.engine.launch_count() > 0
    finally:
        backend.close()
        ring.close()
