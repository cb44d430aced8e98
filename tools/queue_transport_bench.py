#!/usr/bin/env python
"""One-way latency of ServiceRequest packets through a multiprocessing.Queue (maxsize 32, as
DeepRecSys.py:45) with Poisson arrivals: stock ``Queue.put`` (buffer + feeder thread) against
``accelInferenceEngine._direct_put`` (pickle + pipe write in the calling thread).  CPU only.

    python tools/queue_transport_bench.py [--n 8000] > profiles/rNN_queue_transport.txt
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprecsys_b200.accelInferenceEngine import _direct_put      # noqa: E402
from deeprecsys_b200.serve import _sleep_until                    # noqa: E402
from deeprecsys_b200.utils.packets import ServiceRequest          # noqa: E402


def generate(q, n, gap_s, direct):
    put = _direct_put(q) if direct else q.put
    gaps = np.random.RandomState(0).exponential(gap_s, n)
    t = time.time()
    for i in range(n):
        t += gaps[i]
        _sleep_until(t)
        r = ServiceRequest(batch_id=1, epoch=0, batch_size=256, sub_id=0, total_sub_batches=1,
                           exp_packet=False)
        r.arrival_time = time.time()
        put(r)
    put(None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8000)
    opts = ap.parse_args()
    ctx = mp.get_context("fork")
    print("# host cores: %d" % len(os.sched_getaffinity(0)))
    for direct in (0, 1):
        for gap_us in (200, 80, 50, 30):
            q = ctx.Queue(maxsize=32)
            p = ctx.Process(target=generate, args=(q, opts.n, gap_us * 1e-6, direct))
            p.start()
            lat, t0 = [], time.time()
            while True:
                r = q.get()
                if r is None:
                    break
                lat.append(time.time() - r.arrival_time)
            dt = time.time() - t0
            p.join()
            lat = np.array(lat) * 1e3
            print("%-12s offered %6.0f/s delivered %6.0f/s  one-way p50 %.3f ms  p99 %.3f ms"
                  % ("direct write" if direct else "feeder put", 1e6 / gap_us, len(lat) / dt,
                     np.percentile(lat, 50), np.percentile(lat, 99)))


if __name__ == "__main__":
    main()
