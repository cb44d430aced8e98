"""Packed id wire format, host side (drs_pack_ids runs on the CPU): id i occupies bits
[i*bits, (i+1)*bits) of a little-endian uint32 stream — checked against a numpy decoder that
mirrors the GPU's k_unpack_ids window arithmetic."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from deeprecsys_b200 import _lib
from deeprecsys_b200.engine import id_bits, pack_ids


def unpack(words, bits, n):
    w64 = words.astype(np.uint64)
    win = w64[:-1] | (w64[1:] << np.uint64(32))
    bit = np.arange(n, dtype=np.int64) * bits
    return ((win[bit >> 5] >> (bit & 31).astype(np.uint64)) & np.uint64((1 << bits) - 1)).astype(np.int32)


@settings(max_examples=60, deadline=None)
@given(bits=st.integers(1, 31), n=st.integers(0, 300), seed=st.integers(0, 2 ** 31 - 1))
def test_pack_then_unpack_is_the_identity(bits, n, seed):
    rng = np.random.RandomState(seed)
    hi = (1 << bits) - 1
    ids = rng.randint(0, hi + 1, size=n, dtype=np.int64).astype(np.int32)
    if n:
        ids[rng.randint(n)] = hi                         # the largest value that fits
    words = pack_ids(ids, bits)
    assert words.dtype == np.uint32 and len(words) == (n * bits + 31) // 32 + 1
    assert np.array_equal(unpack(words, bits, n), ids)


def test_ids_that_do_not_fit_are_refused_and_widths_follow_the_table():
    with pytest.raises(_lib.DrsError):
        pack_ids(np.array([1, 2, 1 << 22], np.int32), 22)
    with pytest.raises(_lib.DrsError):
        pack_ids(np.array([-1], np.int32), 31)               # a negative id never fits
    assert [id_bits(n) for n in (1, 2, 3, 4, 5, 4000000, 1 << 22, (1 << 22) + 1)] == \
        [1, 1, 2, 2, 3, 22, 22, 23]
    out = np.zeros(2, np.uint32)
    with pytest.raises(ValueError):
        pack_ids(np.arange(100, dtype=np.int32), 22, out=out)  # caller's buffer too small
