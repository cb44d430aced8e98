"""Locality-aware synthetic id traces (LRU stack-distance model).

Host-side counterpart of the reference's
/root/reference/data_generator/trace_generator.py: an id
trace for one embedding table is produced from a cumulative stack-distance
distribution — with probability P(sd = 0) the next id is one never seen before,
otherwise it is the id that sits ``sd`` places below the top of an LRU stack —
so the gather kernel can be driven with realistic reuse instead of
uniform-random rows.

Same functions, arguments and results as the reference module
(``read_dist_from_file`` :32-44, ``generate_stack_distance`` :47-70,
``trace_generate_lru`` :73-100, ``write_trace_to_file`` :102-110) and the same
consumption of both random streams (python's ``random`` for the order of first
references, numpy's legacy global RNG for one uniform draw per reference), so a
seeded run is value-for-value identical (tests/test_trace.py, golden fixtures
recorded from the reference functions).  What differs is the cost: the
reference keeps the whole table as one python list and removes from its middle
on every reuse (O(table) per reference); here only the already-referenced ids
live in the stack, first references are bulk-copied from the permutation and
stack distances are looked up with one vectorised search per call
(O(stack distance) per reuse, ~0 per first reference).
"""
from __future__ import absolute_import, division, print_function

import bisect
import random

import numpy as np

# the reference's global (trace_generator.py:30): ids are not grouped into lines
cache_line_size = 1


def read_dist_from_file(file_path):
    """Two-line profile (ref :32-44): stack distances, then their cumulative probabilities.
    The three-line variant of data_generator/dlrm_data_caffe2.py:365-377 (first line = the ids in
    order of first reference) is recognised too and returned as a 3-tuple."""
    with open(file_path, "r") as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    if len(lines) == 2:
        return ([int(el) for el in lines[0].split(", ")],
                [float(el) for el in lines[1].split(", ")])
    if len(lines) == 3:
        return ([int(el) for el in lines[0].split(", ")],
                [int(el) for el in lines[1].split(", ")],
                [float(el) for el in lines[2].split(", ")])
    raise ValueError("%s: expected 2 or 3 lines of ', '-separated values, found %d"
                     % (file_path, len(lines)))


def write_dist_to_file(file_path, list_sd, cumm_sd, unique_accesses=None):
    """ref trace_profile.py:61-71 (two lines) / dlrm_data_caffe2.py:380-393 (three lines)."""
    with open(file_path, "w") as f:
        for seq in ([unique_accesses] if unique_accesses is not None else []) + [list_sd, cumm_sd]:
            s = str([v.item() if hasattr(v, "item") else v for v in seq])
            f.write(s[1:len(s) - 1] + "\n")


def write_trace_to_file(file_path, syn_trace):
    """One line of ', '-separated ids (ref :102-110)."""
    with open(file_path, "w") as f:
        f.write(", ".join(str(int(v)) for v in syn_trace) + "\n")


def read_trace_from_file(file_path):
    with open(file_path, "r") as f:
        return np.array([int(v) for v in f.readline().split(", ")], dtype=np.uint64)


def generate_stack_distance(cumm_val, cumm_dist, max_i, i, enable_padding=False):
    """One stack distance (ref :47-70): a uniform draw, with the support cut to distances
    <= the number of first references made so far while that is below the largest distance."""
    u = np.random.rand(1)
    if i < max_i:
        j = bisect.bisect(cumm_val, i) - 1
        u *= cumm_dist[j]
    elif enable_padding:
        fi = cumm_dist[0]
        u = (1.0 - fi) * u + fi
    for (j, f) in enumerate(cumm_dist):
        if u <= f:
            return cumm_val[j]
    return None                       # u above the last cumulative value (a cdf not reaching 1)


class LruTraceGenerator(object):
    """Resumable form of ``trace_generate_lru``: ``generate(k)`` returns the next k ids of the
    trace, so a dataset can be produced batch by batch while reuse carries across batches.

    ``line_accesses``: the ids in the order in which they are referenced for the first time.
    """

    def __init__(self, line_accesses, list_sd, cumm_sd, enable_padding=False):
        self.fresh = np.asarray(line_accesses, dtype=np.int64)   # not yet referenced: fresh[head:]
        self.head = 0
        self.stack = []               # referenced ids, least recently used first
        self.list_sd = [int(v) for v in list_sd]
        self.cumm_sd = [float(v) for v in cumm_sd]
        self._sd = np.asarray(self.list_sd, dtype=np.int64)
        self._cdf = np.asarray(self.cumm_sd, dtype=np.float64)
        self.max_sd = self.list_sd[-1]
        self.enable_padding = bool(enable_padding)
        self.new_refs = 0             # the reference's ``i``

    def _first_reference(self):
        if self.head < len(self.fresh):
            ref = int(self.fresh[self.head])
            self.head += 1
        else:                         # every id has been seen: the reference's pop(0) now takes
            ref = self.stack.pop(0)   # the least recently used one
        self.stack.append(ref)
        self.new_refs += 1
        return ref

    def _reuse(self, sd):
        if sd > len(self.stack):
            raise ValueError("stack distance %d exceeds the %d ids referenced so far (table "
                             "smaller than the profile's largest distance?)" % (sd, len(self.stack)))
        ref = self.stack.pop(len(self.stack) - sd)
        self.stack.append(ref)
        return ref

    def _lookup(self, u):
        j = int(np.searchsorted(self._cdf, u, side="left"))     # first j with u <= cdf[j]
        # a running float sum may end a few ulps below 1 (the reference returns None for a draw
        # above it and then fails); such a draw takes the largest distance
        return self.list_sd[min(j, len(self.list_sd) - 1)]

    def generate(self, k):
        k = int(k)
        out = np.empty(k, dtype=np.uint64)
        u = np.random.rand(k)         # the same doubles as k calls of rand(1)
        pos = 0
        # while fewer first references than the largest distance have been made the support of
        # every draw depends on the running count: one at a time
        while pos < k and self.new_refs < self.max_sd:
            j = bisect.bisect(self.list_sd, self.new_refs) - 1
            sd = self._lookup(u[pos] * self.cumm_sd[j])
            out[pos] = self._first_reference() if sd == 0 else self._reuse(sd)
            pos += 1
        if pos == k:
            return out
        rest = u[pos:]
        if self.enable_padding:
            f0 = self.cumm_sd[0]
            rest = (1.0 - f0) * rest + f0
        j = np.minimum(np.searchsorted(self._cdf, rest, side="left"), len(self.list_sd) - 1)
        sds = self._sd[j]
        reuse_at = np.flatnonzero(sds)
        stack = self.stack
        prev = 0
        for p, sd in zip(reuse_at.tolist() + [len(sds)], sds[reuse_at].tolist() + [0]):
            run = p - prev                                        # first references up to p
            if run > 0:
                bulk = min(run, len(self.fresh) - self.head)
                if bulk > 0:
                    ids = self.fresh[self.head:self.head + bulk]
                    out[pos:pos + bulk] = ids
                    stack.extend(ids.tolist())
                    self.head += bulk
                    self.new_refs += bulk
                    pos += bulk
                for _ in range(run - bulk):
                    out[pos] = self._first_reference()
                    pos += 1
            if sd:
                if sd > len(stack):
                    self._reuse(sd)                               # raises
                ref = stack.pop(-sd)
                stack.append(ref)
                out[pos] = ref
                pos += 1
            prev = p + 1
        return out


def trace_generate_lru(table_size, list_sd, cumm_sd, out_trace_len, enable_padding=False):
    """ref :73-100.  ``table_size`` ids, first referenced in an order drawn with python's
    ``random.sample`` exactly as the reference does; returns the trace as uint64 (the reference
    returns a python list of numpy.uint64 with the same values).  The three-line profile variant
    (dlrm_data_caffe2.py:251-276) passes the first-reference order itself instead of a size."""
    if isinstance(table_size, (int, np.integer)):
        line_accesses = random.sample(range(int(table_size)), int(table_size))
    else:
        line_accesses = list(table_size)
    gen = LruTraceGenerator(line_accesses, list_sd, cumm_sd, enable_padding)
    return gen.generate(out_trace_len)


def main(argv=None):
    """Same flags as the reference's script (:113-148)."""
    import argparse
    import time
    parser = argparse.ArgumentParser(description="Trace Generator")
    parser.add_argument("--cumm_file", type=str, default="./profile/sd_cumm")
    parser.add_argument("--gen_trace_file", type=str, default="./syn_traces/tbl1")
    parser.add_argument("--numpy_rand_seed", type=int, default=123)
    parser.add_argument("--print_precision", type=int, default=5)
    parser.add_argument("--table_size", type=int, default=1000000)
    parser.add_argument("--num_batches", type=int, default=1)
    parser.add_argument("--mini_batch_size", type=int, default=32)
    parser.add_argument("--pooling_factor", type=int, default=80)
    args = parser.parse_args(argv)
    np.random.seed(args.numpy_rand_seed)
    np.set_printoptions(precision=args.print_precision)
    print("cumm filepath = ", args.cumm_file)
    list_sd, cumm_sd = read_dist_from_file(args.cumm_file)[-2:]
    gen_len = args.num_batches * args.mini_batch_size * args.pooling_factor
    print("Begin gen traces ....")
    start = time.time()
    trace = trace_generate_lru(args.table_size, list_sd, cumm_sd, gen_len, False)
    print("syn_trace time cost (min) = ", (time.time() - start) / 60)
    print("max of syn_trace = ", trace.max())
    write_trace_to_file(args.gen_trace_file, trace)


if __name__ == "__main__":
    main()
