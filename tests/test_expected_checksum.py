"""The host closed form of the probe checksum (csrc/pattern_math.hpp, exported as b2dp_expected_checksum) against the
oracle's numpy summation: the reference value of every health verdict is computed on the host, never on the GPU under
test, so it is checked here without a GPU."""
import ctypes as C

import pytest
from hypothesis import given, settings, strategies as st

from oracle import probe as oprobe


def _abi(pkg, n_words, seed):
    out = C.c_uint64(0)
    assert pkg._native.lib.b2dp_expected_checksum(n_words, seed, C.byref(out)) == 0
    return out.value


@pytest.mark.parametrize("n_words", [0, 1, 2, 3, 4, 5, 1023, 1024, 4096 + 4, (1 << 20) + 12, 3 * (1 << 20) + 4 * 37,
                                     (8 << 20) // 4 + 4 * 7, 1 << 24, (1 << 24) + 3])
@pytest.mark.parametrize("seed", [0, 0x5EED0000, 0x5EED0007, 0xFFFFFFFF, 0x80000001])
def test_closed_form_equals_the_oracle_sum(pkg, n_words, seed):
    assert _abi(pkg, n_words, seed) == oprobe.expected_checksum(n_words, seed)


@settings(max_examples=60, deadline=None)
@given(st.integers(min_value=0, max_value=1 << 22), st.integers(min_value=0, max_value=(1 << 32) - 1))
def test_closed_form_random_sizes_and_seeds(pkg, n_words, seed):
    assert _abi(pkg, n_words, seed) == oprobe.expected_checksum(n_words, seed)


def test_full_size_golden_vectors(pkg):
    """The committed 1 GiB golden checksums (tests/golden/probe_vectors.json, made from the oracle)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "probe_vectors.json")))
    for seed_s, want in g["full_size"]["seeds"].items():
        assert _abi(pkg, g["full_size"]["n_words"], int(seed_s)) == want


def test_word_index_wraps_at_2_to_32(pkg):
    """One full period of (uint32(i) * K) ^ seed visits every 32-bit value once: sum = 2^31 * (2^32 - 1) for any
    seed; past 2^32 words the pattern starts over (the kernels cast the word index to uint32)."""
    period = (1 << 31) * ((1 << 32) - 1)
    for seed in (0, 0x5EED0000, 0xDEADBEEF):
        assert _abi(pkg, 1 << 32, seed) == period & oprobe.NO_BAD
        extra = (16 << 20) // 4
        assert _abi(pkg, (1 << 32) + extra, seed) == (period + oprobe.expected_checksum(extra, seed)) & oprobe.NO_BAD
        assert _abi(pkg, 3 << 32, seed) == (3 * period) & oprobe.NO_BAD
    assert pkg._native.lib.b2dp_expected_checksum(1, 0, None) == pkg._native.E_INVAL
