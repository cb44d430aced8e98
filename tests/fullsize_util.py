"""Dyadic table values that make full-size pooled vectors predictable by integer arithmetic.

Row r of table t holds, in column c,  (h(r) + 2^hb * (c + D * t)) * 2^-23  with h a hb-bit hash
of the row id.  hb is the largest width for which a sum of L such numerators stays below 2^24,
so every L-row sum is exact in fp32 in ANY order and numpy integer arithmetic on the ids predicts
every pooled element bit for bit: a wrong row, column, table, segment boundary, a dropped or a
duplicated lookup changes the sum.  (tests/test_zfullsize_gpu.py is the RMC1 instance, hb = 9.)
"""
import numpy as np

SCALE = np.float32(2.0 ** -23)


def hash_bits(T, D, L):
    hb = 0
    while max(1, L) * ((1 << (hb + 1)) * T * D) < 2 ** 24 and hb < 12:
        hb += 1
    if hb == 0 or max(1, L) * ((1 << hb) * T * D) >= 2 ** 24:
        raise ValueError("T*D*L too large for an exact fp32 checksum")
    return hb


def row_hash(ids, hb):
    return ((np.asarray(ids).astype(np.uint64) * np.uint64(2654435761)) >> np.uint64(7)) \
        & np.uint64((1 << hb) - 1)


def table_values(rows, t, D, hb):
    """float32 [len(rows), D] for the given row ids of table t."""
    num = row_hash(rows, hb)[:, None] + np.uint64(1 << hb) * (
        np.arange(D, dtype=np.uint64) + np.uint64(D * t))
    return num.astype(np.float32) * SCALE


def expected_pooled(ids_per_table, T, B, D, L, hb):
    """[B, T*D] float32 pooled vectors predicted from the ids alone (fixed L per segment)."""
    out = np.empty((B, T * D), dtype=np.float32)
    for t in range(T):
        h_sum = row_hash(ids_per_table[t][:B * L], hb).reshape(B, L).sum(axis=1)     # exact ints
        col = np.uint64(L << hb) * (np.arange(D, dtype=np.uint64) + np.uint64(D * t))
        assert int(h_sum.max()) + int(col.max()) < 2 ** 24                           # exactness premise
        out[:, t * D:(t + 1) * D] = (h_sum[:, None] + col[None, :]).astype(np.float32) * SCALE
    return out


def compact_tables(ids_per_table, lengths_per_table, pick, spec, hb):
    """The oracle's view of the picked samples: per table the rows they touch (dyadic values),
    ids renumbered into that compact table, lengths."""
    L = spec.L
    ids, tables, lengths = [], [], []
    for t in range(spec.T):
        seg = np.asarray(ids_per_table[t]).reshape(-1, L)[pick].reshape(-1)
        uniq, inv = np.unique(seg, return_inverse=True)
        tables.append(table_values(uniq.astype(np.int64), t, spec.D, hb))
        ids.append(inv.astype(np.int64))
        lengths.append(np.ascontiguousarray(np.asarray(lengths_per_table[t])[pick]))
    return tables, ids, lengths
