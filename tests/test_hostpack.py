"""The engine's host packer (csrc/hostpack.cpp, exposed as drs_hostpack_ids): lane-interleaved
blocks of 512 ids — id r of a block is bits [(r>>4)*bits, +bits) of lane r&15, word w of lane l
at block[w*16 + l].  Checked against a numpy decoder that mirrors k_unpack_ids_blk."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import hostpack_ids


def unpack_blk(words, bits, n):
    i = np.arange(n, dtype=np.int64)
    r = i & 511
    lane, bit = r & 15, (r >> 4) * bits
    w, sh = bit >> 5, (bit & 31).astype(np.uint64)
    base = (i >> 9) * (16 * bits)
    w64 = np.concatenate([words, np.zeros(16, np.uint32)]).astype(np.uint64)
    lo = w64[base + w * 16 + lane]
    nxt = np.minimum(w + 1, bits - 1)            # the window never needs a word past the lane's last
    hi = np.where(bit + bits > (w + 1) * 32, w64[base + nxt * 16 + lane], 0).astype(np.uint64)
    win = lo | (hi << np.uint64(32))
    return ((win >> sh) & np.uint64((1 << bits) - 1)).astype(np.int64)


@settings(max_examples=80, deadline=None)
@given(bits=st.integers(1, 31), n=st.integers(0, 3000), seed=st.integers(0, 2 ** 31 - 1),
       wide=st.booleans(), threads=st.sampled_from([1, 3]))
def test_block_pack_then_unpack_is_the_identity(bits, n, seed, wide, threads):
    rng = np.random.RandomState(seed)
    hi = (1 << bits) - 1
    ids = rng.randint(0, hi + 1, size=n, dtype=np.int64)
    if n:
        ids[rng.randint(n)] = hi
    words, ok = hostpack_ids(ids if wide else ids.astype(np.int32), bits, threads)
    assert ok and words.dtype == np.uint32 and len(words) == (n + 511) // 512 * 16 * bits
    assert np.array_equal(unpack_blk(words, bits, n), ids)


def test_many_chunks_over_the_pool_and_both_id_widths_agree():
    rng = np.random.RandomState(3)
    n = 64 * 512 * 5 + 77                         # five whole chunks and a ragged tail block
    ids = rng.randint(0, 4000000, size=n, dtype=np.int64)
    a, ok_a = hostpack_ids(ids.astype(np.int32), 22, threads=4)
    b, ok_b = hostpack_ids(ids, 22, threads=2)
    c, ok_c = hostpack_ids(ids, 22, threads=1)
    assert ok_a and ok_b and ok_c
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(unpack_blk(a, 22, n), ids)


@pytest.mark.parametrize("bad,dtype", [(1 << 22, np.int32), (-1, np.int32), (-5, np.int64),
                                       (1 << 40, np.int64), ((1 << 32) + 7, np.int64)])
def test_ids_that_do_not_fit_are_reported_not_truncated_silently(bad, dtype):
    ids = np.arange(2000, dtype=dtype)
    ids[1234] = bad
    _, ok = hostpack_ids(ids, 22)
    assert not ok
    _, ok = hostpack_ids(np.arange(2000, dtype=dtype), 22)
    assert ok


def test_arguments_are_checked():
    with pytest.raises(_lib.DrsError):
        hostpack_ids(np.arange(4, dtype=np.int32), 0)
    with pytest.raises(_lib.DrsError):
        hostpack_ids(np.arange(4, dtype=np.int32), 32)
