"""GPU parity: b2p_msm_g2 (G2Affine.MultiExp -- the G2 half of "MSM over G1/G2"; AlgoPlonk reaches it only through
kzg.NewSRS, setup/setup.go:124) against the oracle's affine group law over Fp2, bit-exact affine results."""
import random

import pytest

import helpers as H
from algoplonk_b200 import _lib, api
from oracle import pairing as opair
from oracle import plonk_oracle as po

pytestmark = pytest.mark.gpu
CURVES = ("BN254", "BLS12_381")
REAL = {"BN254": "PerpetualPowersOfTauBN254", "BLS12_381": "DuskBLS12_381"}


def _gen(curve):
    return api.g2_from_mont_bytes(curve, api.g2_unsafe(curve, 1))[0]


def _multiples(cv, gen, n, seed):
    """n points k_i * G2 with known k_i: a random start, then steps of small known multiples (one addition each)."""
    rng = random.Random(seed)
    k = rng.randrange(1, cv.r)
    P = opair.g2_mul(cv, gen, k)
    steps = [opair.g2_mul(cv, gen, j) for j in range(1, 5)]
    ks, pts = [], []
    for _ in range(n):
        ks.append(k)
        pts.append(P)
        j = rng.randrange(1, 5)
        k = (k + j) % cv.r
        P = opair.g2_add(cv, P, steps[j - 1])
    return ks, pts


def _run(curve, pts, sc):
    out = api.msm_g2_raw(curve, api.g2_to_mont_bytes(curve, pts), sc)
    return api.g2_from_mont_bytes(curve, out)[0]


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n", [1, 2, 3, 33, 200])
def test_msm_g2_small_vs_naive_oracle(gpu, curve, n):
    cv = po.CURVES[curve]
    ks, pts = _multiples(cv, _gen(curve), n, n)
    rng = random.Random(n + 1)
    cases = [H.scalars_uniform(cv.r, n, 1), [0] * n, [1] * n, [cv.r - 1] * n, H.scalars_witness_like(cv.r, n, 2),
             [1 << ((17 * i) % 250) for i in range(n)]]
    for sc in cases:
        exp = opair.g2_msm_naive(cv, pts, sc)
        assert _run(curve, pts, sc) == exp
        # and the closed form: sum s_i k_i G2
        assert exp == opair.g2_mul(cv, _gen(curve), sum(s * k for s, k in zip(sc, ks)) % cv.r)
    assert _run(curve, [], []) is None                        # empty input -> infinity


@pytest.mark.parametrize("curve", CURVES)
def test_msm_g2_degenerate_inputs(gpu, curve):
    """The same point many times (bucket collisions -> doublings), P and -P (cancellation), infinity among the bases."""
    cv = po.CURVES[curve]
    gen = _gen(curve)
    ks, pts = _multiples(cv, gen, 8, 5)
    P = pts[0]
    same = [P] * 64
    assert _run(curve, same, [3] * 64) == opair.g2_mul(cv, P, 192)
    assert _run(curve, same, list(range(64))) == opair.g2_mul(cv, P, 64 * 63 // 2)
    assert _run(curve, [P, opair.g2_neg(cv, P)], [7, 7]) is None
    assert _run(curve, [P, opair.g2_neg(cv, P), pts[1]], [cv.r - 2, cv.r - 2, 5]) == opair.g2_mul(cv, pts[1], 5)
    withinf = [pts[0], None, pts[1], None]
    sc = H.scalars_uniform(cv.r, 4, 9)
    assert _run(curve, withinf, sc) == opair.g2_msm_naive(cv, withinf, sc)
    assert _run(curve, [None, None], [1, 2]) is None


@pytest.mark.parametrize("curve", CURVES)
def test_msm_g2_on_the_ceremony_points(gpu, curve):
    """Bases = the two G2 points of the reference's vk.bin ([1]_2, [tau]_2, tau unknown): a [1]_2 + b [tau]_2 against
    the oracle, and through the pairing: e(G1, a [1]_2 + b [tau]_2) == e(a G1 + b [tau]_1, [1]_2)."""
    cv = po.CURVES[curve]
    g2 = H.real_srs_g2(REAL[curve])
    a, b = H.scalars_uniform(cv.r, 2, 3)
    got = _run(curve, list(g2), [a, b])
    assert got == opair.g2_add(cv, opair.g2_mul(cv, g2[0], a), opair.g2_mul(cv, g2[1], b))
    pts = H.real_srs_points(REAL[curve])                     # [1]_1, [tau]_1, ...
    lhs = po.g1_neg(cv, pts[0])
    rhs = po.g1_add(cv, po.g1_mul(cv, pts[0], a), po.g1_mul(cv, pts[1], b))
    assert api.pairing_check(curve, api.points_to_mont_bytes(curve, [lhs, rhs]),
                             api.g2_to_mont_bytes(curve, [got, g2[0]]))


@pytest.mark.parametrize("curve,n", [("BN254", 1 << 14), ("BLS12_381", 1 << 13)])
def test_msm_g2_mid_size_closed_form(gpu, curve, n):
    """sum s_i (k_i G2) == (sum s_i k_i) G2 at a size that fills the GPU; every window width gives the same element."""
    cv = po.CURVES[curve]
    gen = _gen(curve)
    ks, pts = _multiples(cv, gen, n, 77)
    sc = H.scalars_uniform(cv.r, n, 4)
    exp = opair.g2_mul(cv, gen, sum(s * k for s, k in zip(sc, ks)) % cv.r)
    assert _run(curve, pts, sc) == exp
    sw = H.scalars_witness_like(cv.r, n, 6)
    assert _run(curve, pts, sw) == opair.g2_mul(cv, gen, sum(s * k for s, k in zip(sw, ks)) % cv.r)


def test_msm_g2_argument_errors(gpu):
    with pytest.raises(ValueError):
        api.msm_g2_raw("BN254", b"\0" * 100, [1])
    import ctypes as C
    out = C.create_string_buffer(128)
    assert _lib.load().b2p_msm_g2(0, None, None, 3, out) != 0
    assert _lib.load().b2p_msm_g2(7, None, None, 0, out) != 0
