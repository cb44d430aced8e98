"""Randomised properties of the host-side pieces around the path (hypothesis): the LRU trace
generator against a literal list-based restatement of the reference loop
(/root/reference/data_generator/trace_generator.py:73-100), and the shared-memory input ring."""
import bisect
import random

import numpy as np
from hypothesis import given, settings, strategies as st

from deeprecsys_b200.data_generator import trace_generator as tg
from deeprecsys_b200.utils.input_ring import SharedInputRing


def literal_trace(table_size, list_sd, cumm_sd, n, enable_padding):
    """The reference's algorithm, line by line: one python list holds the whole table, new
    references pop its head, reuses pop ``l - sd``; ``u`` is drawn per reference."""
    line_accesses = random.sample(range(table_size), table_size)
    max_sd, l, i, out = list_sd[-1], table_size, 0, []
    for _ in range(n):
        u = np.random.rand(1)
        if i < max_sd:
            u *= cumm_sd[bisect.bisect(list_sd, i) - 1]
        elif enable_padding:
            u = (1.0 - cumm_sd[0]) * u + cumm_sd[0]
        sd = next(list_sd[j] for j, f in enumerate(cumm_sd) if u <= f)
        if sd == 0:
            ref = line_accesses.pop(0)
            i += 1
        else:
            ref = line_accesses.pop(l - sd)
        line_accesses.append(ref)
        out.append(ref)
    return out


@st.composite
def distributions(draw):
    k = draw(st.integers(1, 6))
    dists = sorted(draw(st.sets(st.integers(1, 40), min_size=k - 1, max_size=k - 1)))
    list_sd = [0] + dists
    w = np.array(draw(st.lists(st.floats(0.05, 1.0), min_size=k, max_size=k)))
    cumm = np.cumsum(w / w.sum()).tolist()
    cumm[-1] = 1.0
    return list_sd, cumm


@settings(max_examples=60, deadline=None)
@given(distributions(), st.integers(41, 300), st.integers(0, 700), st.booleans(),
       st.integers(0, 2 ** 31 - 1))
def test_generator_equals_the_literal_reference_loop(dist, table_size, n, padding, seed):
    list_sd, cumm = dist
    random.seed(seed)
    np.random.seed(seed)
    want = literal_trace(table_size, list_sd, cumm, n, padding)
    tail = np.random.rand(1)
    random.seed(seed)
    np.random.seed(seed)
    got = tg.trace_generate_lru(table_size, list_sd, cumm, n, padding)
    assert got.tolist() == want and np.random.rand(1) == tail
    assert all(0 <= v < table_size for v in want)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 5), st.integers(0, 7), st.integers(1, 24), st.integers(0, 9),
       st.integers(0, 2 ** 31 - 1), st.booleans())
def test_ring_round_trip_of_ragged_queries(T, m_den, max_batch, max_len, seed, as_i64):
    rng = np.random.RandomState(seed)
    ring = SharedInputRing.create(n_slots=3, n_tables=T, m_den=m_den, max_batch=max_batch,
                                  max_lookups=max_batch * max(1, max_len))
    try:
        written = {}
        for slot in (2, 0, 1):
            B = int(rng.randint(0, max_batch + 1))
            lengths = [rng.randint(0, max_len + 1, size=B).astype(np.int32) for _ in range(T)]
            ids = [rng.randint(0, 2 ** 31 - 1, size=int(l.sum())).astype(
                np.int64 if as_i64 else np.int32) for l in lengths]
            dense = rng.rand(B, m_den).astype(np.float32)
            counts = ring.write(slot, dense, ids, lengths, B)
            written[slot] = (B, dense, ids, lengths, counts)
        for slot, (B, dense, ids, lengths, counts) in written.items():   # slots do not overlap
            rd, ri, rl = ring.read(slot, B, counts)
            assert (rd is None) == (m_den == 0)
            if m_den:
                assert np.array_equal(rd, dense)
            for t in range(T):
                assert counts[t] == len(ids[t])
                assert np.array_equal(ri[t], ids[t]) and np.array_equal(rl[t], lengths[t])
    finally:
        ring.close()
