"""Philox4x32-10 known-answer vectors (Random123 v1.09 kat_vectors) and the derived-draw contract."""
import numpy as np
import pytest

from oracle import draws, oracle

KAT = [  # (counter, key, expected) -- Random123 kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_kat_numpy(ctr, key, want):
    got = draws.philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))
    assert [int(x) for x in got] == want


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_kat_c_oracle(ctr, key, want):
    assert [int(x) for x in oracle.philox(ctr, key)] == want


def test_c_and_numpy_draws_agree():
    rng = np.random.default_rng(0)
    for _ in range(200):
        seed = int(rng.integers(0, 2 ** 63))
        env = int(rng.integers(0, 2 ** 40))
        ep, st, blk = int(rng.integers(0, 1000)), int(rng.integers(0, 3000)), int(rng.integers(0, 18))
        raw_c, uni_c, nrm_c = oracle.draws(seed, env, ep, st, blk)
        raw_n = draws.block(seed, env, ep, st, blk)
        assert (raw_c == raw_n).all()
        assert (uni_c == draws.u01(raw_n)).all()  # uniforms are exact
        np.testing.assert_allclose(nrm_c, draws.normals(raw_n), rtol=3e-6, atol=3e-7)


def test_uniform_ranges():
    x = np.array([0, 0xff, 0x100, 0xffffffff], np.uint32)
    u = draws.u01(x)
    assert u[0] == 0.0 and u[1] == 0.0 and u[-1] < 1.0
    uo = draws.u01_open(x)
    assert uo[0] > 0.0 and uo[-1] == 1.0


def test_normal_moments():
    z = draws.normals(draws.block(1234, np.arange(100000), 0, 1, draws.BLK_EVADE))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert np.isfinite(z).all()
