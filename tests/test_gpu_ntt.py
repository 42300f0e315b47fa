"""GPU parity: b2p_ntt (fft.Domain.FFT / FFTInverse) against the oracles, bit-exact."""
import random

import pytest

import helpers as H
from algoplonk_b200 import api
from oracle import cpu_oracle as co
from oracle import plonk_oracle as po

pytestmark = pytest.mark.gpu
CURVES = ("BN254", "BLS12_381")


def _le_to_mont(curve, raw_le: bytes) -> bytes:
    vals = [int.from_bytes(raw_le[i:i + 32], "little") for i in range(0, len(raw_le), 32)]
    return api.fr_to_mont_bytes(curve, vals)


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("logn", [0, 1, 2, 3, 5, 8, 9, 10, 11, 12, 13])
def test_ntt_small_vs_bigint_oracle(gpu, curve, logn):
    cv = po.CURVES[curve]
    n = 1 << logn
    a = H.scalars_uniform(cv.r, n, 100 + logn)
    w = po.domain_generator(cv, n)
    if logn <= 10:
        fwd, cos = po.ntt(cv, a, w), po.coset_ntt(cv, a, w, cv.coset_shift)
    else:
        fwd, cos = co.ntt(cv.cid, a), co.ntt(cv.cid, a, coset=True)
    got = api.ntt(curve, a)
    assert got == fwd
    assert api.ntt(curve, got, inverse=True) == a
    gotc = api.ntt(curve, a, coset=True)
    assert gotc == cos
    assert api.ntt(curve, gotc, inverse=True, coset=True) == a


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("logn", [16, 17, 19, 20, 22, 23])
def test_ntt_large_vs_cpp_oracle(gpu, curve, logn):
    """Config sizes (2^17 / 2^20-row circuits use domains 2^17..2^22, the 2^21-row BLS12-381 config 2^23):
    multi-pass path, bit-exact against the C++ oracle."""
    cv = po.CURVES[curve]
    if logn == 23 and curve == "BN254":
        pytest.skip("4n = 2^23 only occurs for the 2^21-row BLS12-381 config")
    n = 1 << logn
    rng = random.Random(logn)
    raw = b"".join(rng.randrange(cv.r).to_bytes(32, "little") for _ in range(n))
    vals = [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]
    for coset in (False, True):
        exp = co.ntt_bytes(cv.cid, raw, coset=coset)
        got = api.ntt(curve, vals, coset=coset)
        assert b"".join(v.to_bytes(32, "little") for v in got) == exp
        assert api.ntt(curve, got, inverse=True, coset=coset) == vals


@pytest.mark.parametrize("curve", CURVES)
def test_ntt_properties_at_full_size(gpu, curve):
    """2^22 (= 4n for the 2^20-row configs): round trip, linearity, and the delta / constant pair,
    none of which need an oracle run."""
    cv = po.CURVES[curve]
    n = 1 << 22
    rng = random.Random(22)
    seed_vals = [rng.randrange(cv.r) for _ in range(4096)]
    a = [seed_vals[(i * 7919) & 4095] for i in range(n)]
    b = [seed_vals[(i * 104729 + 13) & 4095] for i in range(n)]
    fa, fb = api.ntt(curve, a), api.ntt(curve, b)
    assert api.ntt(curve, fa, inverse=True) == a
    s = [(x + y) % cv.r for x, y in zip(a, b)]
    assert api.ntt(curve, s) == [(x + y) % cv.r for x, y in zip(fa, fb)]
    delta = [0] * n
    delta[0] = 5
    assert api.ntt(curve, delta) == [5] * n
    delta[0], delta[1] = 0, 1
    w = po.domain_generator(cv, n)
    f1 = api.ntt(curve, delta)
    assert f1[0] == 1 and f1[1] == w and f1[n // 2] == cv.r - 1 and f1[3] == pow(w, 3, cv.r)


def test_ntt_rejects_bad_sizes(gpu):
    with pytest.raises(gpu.B200PlonkError):
        api.ntt("BN254", [1, 2, 3])
    with pytest.raises(gpu.B200PlonkError):
        api.ntt("BN254", [])
