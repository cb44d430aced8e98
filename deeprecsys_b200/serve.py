"""System-level serving harness for the accel engine (the "inferences/sec at p99"
half of the metric) for boxes where the reference checkout is not available.

It drives ``accelInferenceEngine`` through the SAME process/Queue protocol as
the reference orchestrator (/root/reference/DeepRecSys.py:44-138): a bounded
``accelRequestQueue`` (maxsize 32, DeepRecSys.py:45), one response Queue, one
ready Queue, ``ServiceRequest`` in, ``ServiceResponse`` out, ``None`` sentinel.
The load generator follows loadGenerator.py: query sizes from
``model_batch_size_distribution`` (:20-43: normal / lognormal / fixed, clipped
to [1, max_mini_batch_size]), Poisson arrivals (:198) — except that
inter-arrival times are continuous exponentials (sub-millisecond resolution)
instead of integer milliseconds, because a B200 serves a 1024-item RMC1 query
in ~0.2 ms.  Latency = inference_end_time - arrival_time, as
DeepRecSys.py:122-123 computes it.

    python -m deeprecsys_b200.serve --config_file deeprecsys_b200/models/configs/dlrm_rm1.json \
        --num_batches 32 --max_mini_batch_size 1024 --batch_size_distribution normal \
        --avg_mini_batch_size 256 --var_mini_batch_size 32 --avg_arrival_rate 0.25 --nepochs 40

``--serve_rates_ms a,b,c`` sweeps mean inter-arrival times, ``--serve_queries N`` sets the
queries per rate, ``--serve_generators G`` splits the arrival stream over G generator processes,
``--serve_ring`` makes every request carry its own tensors through a
shared-memory ring (utils/input_ring.py) instead of naming a pre-generated batch.
"""
from __future__ import absolute_import, division, print_function

import json
import multiprocessing as mp
import sys
import time

import numpy as np

from deeprecsys_b200.accelInferenceEngine import accelInferenceEngine
from deeprecsys_b200.utils.packets import ServiceRequest
from deeprecsys_b200.utils.utils import cli


def query_sizes(args, n, rng):
    """loadGenerator.py:20-43 (file-based distribution omitted)."""
    dist = args.batch_size_distribution
    if dist == "normal":
        s = rng.normal(args.avg_mini_batch_size, args.var_mini_batch_size, n)
    elif dist == "lognormal":
        s = rng.lognormal(args.avg_mini_batch_size, args.var_mini_batch_size, n)
    elif dist == "fixed":
        s = np.full(n, args.avg_mini_batch_size, dtype=np.float64)
    else:
        raise ValueError("batch_size_distribution %r not supported by this harness" % dist)
    return np.clip(s, 1, args.max_mini_batch_size).astype(np.int64)


def _sleep_until(t):
    while True:
        dt = t - time.time()
        if dt <= 0:
            return
        if dt > 0.002:
            time.sleep(dt - 0.001)      # coarse sleep, then spin (loadGenerator.py:57-64 does the same)
        else:
            time.sleep(0)               # yield the core between polls


def _generate_load(args, accel_q, report_q, n_total, warmup_queries, arrival_ms, seed,
                   ring=None, dataset=None):
    """Load-generator process (the reference runs loadGenerator() as its own process too,
    DeepRecSys.py:74-81): Poisson arrivals of ServiceRequests into the bounded accel queue.
    With ``ring``: every request carries its inputs — the generator writes rows [0, B) of batch
    ``batch_id`` of its own copy of the dataset into the next ring slot and sends a
    TensorServiceRequest naming the slot (the engine then reads nothing from its own dataset)."""
    from deeprecsys_b200.accelInferenceEngine import _direct_put
    from deeprecsys_b200.utils.packets import TensorServiceRequest
    sys.setswitchinterval(1e-4)
    put = _direct_put(accel_q)          # same bytes as Queue.put, written without a feeder thread
    rng = np.random.RandomState(seed)
    sizes = query_sizes(args, n_total, rng)
    gaps = rng.exponential(arrival_ms / 1000.0, n_total)
    nb = max(1, args.num_batches)
    ring_desc = ring.descriptor() if ring is not None else None
    t_next = time.time()
    t_first = None
    for i in range(n_total):
        t_next += gaps[i]
        _sleep_until(t_next)
        if ring is None:
            req = ServiceRequest(batch_id=i % nb, epoch=i // nb, batch_size=int(sizes[i]),
                                 sub_id=0, total_sub_batches=1, exp_packet=(i < warmup_queries))
        else:
            # the query "arrives" with its tensors: writing them into the slot is the client's
            # work and happens before the arrival stamp
            lX, lS_l, lS_i = dataset
            b, B, slot = i % nb, int(sizes[i]), i % ring.n_slots
            ring.acquire(slot)              # released by the engine when the slot's query is done
            counts = ring.write(slot, lX[b] if ring.m_den > 0 else None, lS_i[b], lS_l[b], B)
            req = TensorServiceRequest(ring=ring_desc, slot=slot, counts=counts, batch_id=b,
                                       epoch=i // nb, batch_size=B, sub_id=0,
                                       total_sub_batches=1, exp_packet=(i < warmup_queries))
        req.arrival_time = time.time()
        if i == warmup_queries:
            t_first = req.arrival_time
        put(req)                        # blocks when the engine falls behind (bounded queue)
    report_q.put((t_first, time.time()))


class SystemHarness(object):
    """One accel engine process (started once: dataset + parameters are regenerated inside it,
    like every engine process of the reference) driven through the reference's Queue protocol;
    ``run`` performs one experiment at a fixed mean inter-arrival time.  Process layout as in
    DeepRecSys.py: the load generator is its own process, this (parent) process only collects
    responses (DeepRecSys.py:92-123)."""

    def __init__(self, args, ring=False):
        self.args = args
        self.ring, self.dataset = None, None
        if ring:
            # the load generator's own copy of the queries (same seed -> the data every engine
            # process regenerates) and the shared-memory ring their tensors travel through
            from deeprecsys_b200.data_generator.dlrm_data import DLRMDataGenerator
            from deeprecsys_b200.spec import ModelSpec
            from deeprecsys_b200.utils.input_ring import SharedInputRing, slots_needed
            np.random.seed(args.numpy_rand_seed)
            _, lX, lS_l, lS_i = DLRMDataGenerator(args).generate_input_data()
            self.dataset = (lX, lS_l, lS_i)
            spec = ModelSpec(args)
            mb = int(args.max_mini_batch_size)
            consumers = int(getattr(args, "accel_gpus", 1) or 1)
            depth = max(4, int(getattr(args, "accel_coalesce", 0) or 0) * 2)
            self.ring = SharedInputRing.create(slots_needed(32, depth, consumers), spec.T,
                                               spec.m_den, mb, mb * max(1, spec.L))
        self.ctx = ctx = mp.get_context("fork")   # DeepRecSys.py uses the default (fork) start method
        self.accel_q = ctx.Queue(maxsize=32)                  # DeepRecSys.py:45
        self.resp_q, self.ready_q = ctx.Queue(), ctx.Queue()
        self.proc = ctx.Process(target=accelInferenceEngine,
                                args=(args, self.accel_q, 0, self.resp_q, self.ready_q))
        self.proc.daemon = True
        self.proc.start()
        from deeprecsys_b200.utils.fastqueue import FrameReader
        self.resp_reader = FrameReader(self.resp_q)
        self.ready_q.get()                                    # loadGenerator.py:76-78

    def run(self, n_queries, arrival_ms, seed=0, warmup_queries=20, timeout_s=300.0,
            generators=1):
        """``generators`` > 1: that many load-generator processes, each a Poisson stream of
        1/generators of the rate (their superposition is the Poisson stream of the full rate) —
        one Python process tops out near 80 k requests/s, below what a B200 engine can take."""
        import queue as _queue
        generators = 1 if self.ring is not None else max(1, int(generators))
        per_gen = -(-n_queries // generators)
        warm_gen = -(-warmup_queries // generators)
        n_queries, warmup_queries = per_gen * generators, warm_gen * generators
        n_total = n_queries + warmup_queries
        report_q = self.ctx.Queue()
        gens = []
        for g in range(generators):
            gen = self.ctx.Process(target=_generate_load,
                                   args=(self.args, self.accel_q, report_q, per_gen + warm_gen,
                                         warm_gen, arrival_ms * generators, seed * 131 + g,
                                         self.ring, self.dataset))
            gen.daemon = True
            gen.start()
            gens.append(gen)
        got = []
        deadline = time.time() + timeout_s
        # the collector must keep up with ~100 k responses/s: batched pipe reads (same protocol
        # as resp_q.get(), utils/fastqueue.py); latencies come from the stamps inside the
        # responses, not from when the collector sees them
        ended = False
        while len(got) < n_total and time.time() < deadline and not ended:
            batch = self.resp_reader.drain(block=True, timeout=1.0)
            if not batch:
                if not self.proc.is_alive():
                    break
                continue
            for r in batch:
                if r is None:
                    ended = True
                    break
                got.append(r)
        t_first = None
        try:
            firsts = [report_q.get(timeout=10.0)[0] for _ in gens]
            t_first = min(f for f in firsts if f is not None)
        except (_queue.Empty, ValueError):
            t_first = None
        for gen in gens:
            gen.join(timeout=10)
            if gen.is_alive():
                gen.terminate()
        meas = [r for r in got if not r.exp_packet]
        if not meas or t_first is None:
            raise RuntimeError("serving run produced %d of %d responses" % (len(got), n_total))
        lat = np.array([r.inference_end_time - r.arrival_time for r in meas]) * 1e3
        items = int(sum(r.batch_size for r in meas))
        failed = sum(1 for r in meas if r.out_batch_size != r.batch_size)
        failed += n_queries - len(meas)
        span = max(r.inference_end_time for r in meas) - t_first
        return {
            "arrival_ms": arrival_ms, "generators": generators, "queries": len(meas),
            "failed": failed, "items": items,
            "offered_qps": 1000.0 / arrival_ms, "qps": len(meas) / span,
            "inferences_per_s": items / span,
            "p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)),
            "p99_ms": float(np.percentile(lat, 99)), "max_ms": float(lat.max()),
            "mean_query_size": float(np.mean([r.batch_size for r in meas])),
        }

    def close(self):
        import queue as _queue
        self.accel_q.put(None)
        deadline = time.time() + 120
        ended = False
        while time.time() < deadline and not ended:
            batch = self.resp_reader.drain(block=True, timeout=1.0)
            if not batch and not self.proc.is_alive():
                break
            ended = any(r is None for r in batch)
        self.proc.join(timeout=60)
        if self.proc.is_alive():
            self.proc.terminate()
        if self.ring is not None:
            self.ring.close()


def run_system(args, n_queries, arrival_ms, seed=0, warmup_queries=20):
    """One experiment at a fixed mean inter-arrival time.  Returns a dict of stats."""
    h = SystemHarness(args)
    try:
        return h.run(n_queries, arrival_ms, seed=seed, warmup_queries=warmup_queries)
    finally:
        h.close()


def sweep(args, rates_ms, n_queries, target_p99_ms=None, ring=False, generators=1,
          min_seconds=0.0):
    """Run the experiment at each mean inter-arrival time against ONE engine process; report
    the best throughput whose p99 meets ``target_p99_ms`` (args.target_latency by default).
    ``ring``: requests carry their inputs through a shared-memory ring (TensorServiceRequest)."""
    target = float(target_p99_ms if target_p99_ms is not None else args.target_latency)
    h = SystemHarness(args, ring=ring)
    try:
        # at least n_queries per rate, and at least min_seconds of arrivals (a fast rate would
        # otherwise be over before the queues reach their steady state)
        rows = [h.run(max(n_queries, int(1000.0 * min_seconds / r)), r, seed=17 + i,
                      generators=generators) for i, r in enumerate(rates_ms)]
    finally:
        h.close()
    ok = [r for r in rows if r["p99_ms"] <= target and r["failed"] == 0]
    best = max(ok, key=lambda r: r["inferences_per_s"]) if ok else None
    return {"target_p99_ms": target, "runs": rows, "best": best,
            "inputs": "shared-memory ring (TensorServiceRequest)" if ring else "batch_id",
            "engine": {"accel_coalesce": int(getattr(args, "accel_coalesce", 1)),
                       "accel_stack_rows": int(getattr(args, "accel_stack_rows", 0)),
                       "accel_responder": int(getattr(args, "accel_responder", 0)),
                       "accel_lanes": int(getattr(args, "accel_lanes", 0)),
                       "accel_gpus": int(getattr(args, "accel_gpus", 1))}}


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    rates, n_queries = None, 2000
    if "--serve_rates_ms" in argv:
        i = argv.index("--serve_rates_ms")
        rates = [float(x) for x in argv[i + 1].split(",")]
        del argv[i:i + 2]
    if "--serve_queries" in argv:
        i = argv.index("--serve_queries")
        n_queries = int(argv[i + 1])
        del argv[i:i + 2]
    generators = 1
    if "--serve_generators" in argv:
        i = argv.index("--serve_generators")
        generators = int(argv[i + 1])
        del argv[i:i + 2]
    ring = "--serve_ring" in argv
    if ring:
        argv.remove("--serve_ring")
    args = cli(argv)
    if not args.model_name:
        args.model_name = {"dlrm": "rm1"}.get(args.model_type, args.model_type)
    rates = rates or [args.avg_arrival_rate]
    out = sweep(args, rates, n_queries, ring=ring, generators=generators, min_seconds=0.3)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
