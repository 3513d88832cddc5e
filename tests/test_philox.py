"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for BOTH implementations: the
product's (host-compiled from csrc/philox.cuh, exported as a test hook) and the oracle's."""
import ctypes as C

import numpy as np
import pytest

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _run(fn, ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    fn(c, k, o)
    return tuple(o)


@pytest.mark.parametrize("ctr,key,expect", KATS)
def test_product_philox_kat(chaos_lib, ctr, key, expect):
    assert _run(chaos_lib.cl_philox4x32_10, ctr, key) == expect


@pytest.mark.parametrize("ctr,key,expect", KATS)
def test_oracle_philox_kat(oracle_api, ctr, key, expect):
    assert _run(oracle_api.lib().orc_philox4x32_10, ctr, key) == expect


def test_product_and_oracle_agree_on_random_counters(chaos_lib, oracle_api):
    rng = np.random.default_rng(0)
    for _ in range(200):
        ctr = tuple(int(x) for x in rng.integers(0, 2**32, 4))
        key = tuple(int(x) for x in rng.integers(0, 2**32, 2))
        assert _run(chaos_lib.cl_philox4x32_10, ctr, key) == _run(oracle_api.lib().orc_philox4x32_10, ctr, key)


def test_uniform53_matches_numpy_construction(chaos_lib):
    rng = np.random.default_rng(1)
    for _ in range(200):
        a, b = (int(x) for x in rng.integers(0, 2**32, 2))
        u = ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0
        assert chaos_lib.cl_uniform53(a, b, -30.0, 30.0) == -30.0 + 60.0 * u
        assert 0.0 <= chaos_lib.cl_uniform53(a, b, 0.0, 1.0) < 1.0
