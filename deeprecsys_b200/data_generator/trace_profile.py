"""Stack-distance profile of an id trace — the input of trace_generator.

Host-side counterpart of /root/reference/data_generator/trace_profile.py: the
reference rescans a ``max_stack_distance`` window of the trace for every
reference (``np.where`` + ``set``, :39-60); here the position of every id's most
recent reference is kept, so one reference costs a dictionary lookup and, for a
reuse, one sum over at most ``max_stack_distance`` flags.  Results are identical
(tests/test_trace.py, golden fixtures recorded from the reference function).
"""
from __future__ import absolute_import, division, print_function

import collections

import numpy as np

from deeprecsys_b200.data_generator.trace_generator import write_dist_to_file


def trace_profile(trace, max_stack_distance):
    """ref :39-60.  For reference i of id x: if x occurs among the previous
    ``max_stack_distance`` references, the distance is the number of distinct ids referenced
    since (and including) its latest occurrence; otherwise 0, and x is appended to
    ``line_accesses`` (also when x was seen earlier, outside the window — as the reference).
    Returns (stack_distances, line_accesses) as python lists."""
    trace = np.asarray(trace)
    n = len(trace)
    window = int(max_stack_distance)
    latest = {}                                   # id -> index of its most recent reference
    is_latest = np.zeros(n, dtype=np.int32)       # 1 where trace[q] has not been referenced since
    stack_distances, line_accesses = [], []
    for i in range(n):
        x = trace[i]
        key = x.item() if hasattr(x, "item") else x
        p = latest.get(key)
        if p is not None and p >= i - window:
            stack_distances.append(int(is_latest[p:i].sum()))
        else:
            stack_distances.append(0)
            line_accesses.append(x)
        if p is not None:
            is_latest[p] = 0
        is_latest[i] = 1
        latest[key] = i
    return stack_distances, line_accesses


def stack_distance_distribution(stack_distances):
    """(list_sd, prob_sd, cumm_sd) exactly as the reference's script accumulates them
    (:120-136: running float sum of count / len, ascending distance)."""
    total = float(len(stack_distances))
    dc = sorted(collections.Counter(stack_distances).items())
    list_sd = [d for d, _ in dc]
    prob_sd, cumm_sd = [], []
    for i, (_, k) in enumerate(dc):
        prob_sd.append(k / total)
        cumm_sd.append(k / total if i == 0 else cumm_sd[i - 1] + k / total)
    return list_sd, prob_sd, cumm_sd


def read_trace_from_file(file_path):
    """Space-separated ids, any number of lines (the reference reads it with pandas, :32-36)."""
    with open(file_path, "r") as f:
        return np.array(f.read().split(), dtype=np.int64)


def zipf_trace(n, table_size, alpha):
    """n ids whose popularity follows Zipf(alpha): rank r -> row (r * 2654435761) mod table_size,
    so popular rows are scattered over the table rather than packed at its start."""
    ranks = np.random.zipf(alpha, int(n)).astype(np.uint64)
    return ((ranks * np.uint64(2654435761)) % np.uint64(table_size)).astype(np.int64)


def main(argv=None):
    """Same flags as the reference's script (:74-147)."""
    import argparse
    parser = argparse.ArgumentParser(description="Profile traces stack distance distribution")
    parser.add_argument("--profile_trace_file", type=str, default="./original_traces/tbl1")
    parser.add_argument("--profile_trace_random", type=int, default=1)
    parser.add_argument("--profile_len", type=int, default=1000000)
    parser.add_argument("--max_stack_distance", type=int, default=10000)
    parser.add_argument("--sd_file_prob", type=str, default="./profile/sd_prob")
    parser.add_argument("--sd_file_cumm", type=str, default="./profile/sd_cumm")
    parser.add_argument("--numpy_rand_seed", type=int, default=123)
    parser.add_argument("--print_precision", type=int, default=5)
    # not in the reference: a popularity-skewed source trace (Zipf ranks scattered over the table)
    parser.add_argument("--profile_trace_zipf", type=float, default=0.0)
    parser.add_argument("--table_size", type=int, default=10000000)
    args = parser.parse_args(argv)
    np.random.seed(args.numpy_rand_seed)
    np.set_printoptions(precision=args.print_precision)
    if args.profile_trace_zipf > 1.0:
        trace = zipf_trace(args.profile_len, args.table_size, args.profile_trace_zipf)
    elif args.profile_trace_random == 0:
        raw = read_trace_from_file(args.profile_trace_file)
        trace = raw[:min(args.profile_len, len(raw))]
    else:
        trace = np.random.randint(0, args.table_size, args.profile_len, np.int32)
    stack_distances, line_accesses = trace_profile(trace, args.max_stack_distance)
    print("len of stack_distances = ", len(stack_distances))
    print("len of line accesses = ", len(line_accesses))
    list_sd, prob_sd, cumm_sd = stack_distance_distribution(stack_distances)
    write_dist_to_file(args.sd_file_prob, list_sd, prob_sd)
    write_dist_to_file(args.sd_file_cumm, list_sd, cumm_sd)


if __name__ == "__main__":
    main()
