"""GPU parity: b2p_msm_g1 (G1Affine.MultiExp via kzg.Commit) against the oracles, bit-exact
affine results, plus linearity properties at the benchmark size."""
import random

import pytest

import helpers as H
from algoplonk_b200 import _lib, api
from oracle import cpu_oracle as co
from oracle import plonk_oracle as po

pytestmark = pytest.mark.gpu
CURVES = ("BN254", "BLS12_381")


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("n", [1, 2, 3, 33, 300, 1000])
def test_msm_small_vs_bigint_oracle(gpu, curve, n):
    cv = po.CURVES[curve]
    srs = api.SRS.unsafe(curve, n, H.TAU)
    pts = po.srs_from_tau(cv, H.TAU, n) if n <= 300 else co.srs_from_tau(cv.cid, H.TAU, n)
    assert srs.points(0, n) == pts                      # unsafekzg.NewSRS parity
    rng = random.Random(n)
    cases = [H.scalars_uniform(cv.r, n, 1), [0] * n, [1] * n, [cv.r - 1] * n, H.scalars_witness_like(cv.r, n, 2),
             [rng.randrange(cv.r) for _ in range(max(1, n // 2))],     # fewer scalars than points
             [1 << ((17 * i) % 250) for i in range(n)]]
    for sc in cases:
        exp = po.msm_naive(cv, pts, sc) if n <= 300 else co.msm(cv.cid, pts[: len(sc)], sc)
        assert srs.msm(sc) == exp
    assert srs.msm([]) is None                           # empty input -> infinity
    srs.free()


@pytest.mark.parametrize("curve", CURVES)
def test_msm_window_sweep(gpu, curve, monkeypatch):
    """Every window width the planner can pick must give the same group element."""
    cv = po.CURVES[curve]
    n = 200
    pts = po.srs_from_tau(cv, H.TAU, n)
    sc = H.scalars_witness_like(cv.r, n, 5)
    exp = po.msm_naive(cv, pts, sc)
    for c in (2, 3, 7, 11, 13, 16):
        monkeypatch.setenv("B2P_MSM_C", str(c))
        srs = api.SRS.from_points(curve, pts)
        assert srs.msm_params()[0] == c
        assert srs.msm(sc) == exp
        srs.free()


@pytest.mark.parametrize("name", ["PerpetualPowersOfTauBN254", "DuskBLS12_381"])
def test_msm_on_real_srs_slice(gpu, name):
    """Bases = points of the reference's embedded trusted setups (b2p_srs_load path)."""
    ent = H.srs_kat()[name]
    curve = ent["curve"]
    cv = po.CURVES[curve]
    pts = H.real_srs_points(name)
    srs = api.SRS.from_points(curve, pts)
    assert srs.size == len(pts) and srs.points(0, len(pts)) == pts
    for seed in (1, 2):
        sc = H.scalars_uniform(cv.r, len(pts), seed)
        assert srs.msm(sc) == po.msm_naive(cv, pts, sc)
    srs.free()


@pytest.mark.parametrize("curve", CURVES)
def test_msm_lagrange_basis(gpu, curve):
    """MSM(Lagrange SRS, v) == commit(iNTT(v)) (kzg.ToLagrangeG1, setup/setup.go:124,138)."""
    cv = po.CURVES[curve]
    n = 64
    srs = api.SRS.unsafe(curve, n + 3, H.TAU)
    pts = po.srs_from_tau(cv, H.TAU, n)
    v = H.scalars_uniform(cv.r, n, 8)
    coeffs = po.intt(cv, v, po.domain_generator(cv, n))
    assert srs.msm(v, basis=_lib.BASIS_LAGRANGE) == po.msm_naive(cv, pts, coeffs)
    srs.free()


@pytest.mark.parametrize("curve,logn", [("BN254", 14), ("BN254", 17), ("BLS12_381", 15)])
def test_msm_mid_size_vs_cpp_oracle(gpu, curve, logn):
    cv = po.CURVES[curve]
    n = (1 << logn) + 3
    srs = api.SRS.unsafe(curve, n, H.TAU)
    pts_le = co.srs_from_tau_bytes(cv.cid, H.TAU, n)
    got_pts = srs.points(0, n)
    assert co.points_le(cv.cid, got_pts) == pts_le
    for dist, seed in ((H.scalars_uniform, 3), (H.scalars_witness_like, 4)):
        sc = dist(cv.r, n, seed)
        assert srs.msm(sc) == co.msm_bytes(cv.cid, pts_le, co.scalars_le(sc))
    srs.free()


@pytest.mark.parametrize("curve", CURVES)
def test_msm_benchmark_size_vs_cpp_oracle(gpu, curve):
    """2^20 + 3 points (the SRS of the 2^20-row configs), uniform and witness-like scalars, against the C++ oracle's
    Pippenger -- bit-exact affine result; the generated SRS itself is compared point by point as well."""
    cv = po.CURVES[curve]
    n = (1 << 20) + 3
    srs = api.SRS.unsafe(curve, n, H.TAU)
    pts_le = co.srs_from_tau_bytes(cv.cid, H.TAU, n)
    out = bytearray()
    for first in range(0, n, 1 << 18):
        out += co.points_le(cv.cid, srs.points(first, min(1 << 18, n - first)))
    assert bytes(out) == pts_le
    for dist, seed in ((H.scalars_uniform, 31), (H.scalars_witness_like, 32)):
        sc = dist(cv.r, n, seed)
        assert srs.msm(sc) == co.msm_bytes(cv.cid, pts_le, co.scalars_le(sc))
    srs.free()


def test_msm_properties_at_benchmark_size(gpu):
    """2^20 + 3 points (the north-star size): linearity and unit vectors, no oracle run needed."""
    curve = "BN254"
    cv = po.BN254
    n = (1 << 20) + 3
    srs = api.SRS.unsafe(curve, n, H.TAU)
    rng = random.Random(20)
    R = 1 << 256
    s = [rng.randrange(cv.r) for _ in range(n)]
    t = [rng.randrange(cv.r) if i % 3 else 0 for i in range(n)]
    st = [(a + b) % cv.r for a, b in zip(s, t)]
    ps, pt, pst = srs.msm(s), srs.msm(t), srs.msm(st)
    assert po.g1_add(cv, ps, pt) == pst
    # a geometric scalar vector: sum_j x^j tau^j G = ((x tau)^n - 1)/(x tau - 1) G
    x = 0xDEADBEEF
    geo = [pow(x, j, cv.r) for j in range(n)]
    q = x * H.TAU % cv.r
    closed = (pow(q, n, cv.r) - 1) * pow(q - 1, -1, cv.r) % cv.r
    assert srs.msm(geo) == po.g1_mul(cv, cv.g1, closed)
    # unit vector picks the last point
    e = [0] * n
    e[n - 1] = 1
    assert srs.msm(e) == srs.points(n - 1, 1)[0]
    srs.free()


@pytest.mark.parametrize("curve", CURVES)
def test_msm_repeated_and_opposite_points(gpu, curve):
    """Collisions: an SRS made of few distinct points, their negatives and infinity, with equal scalars, so
    buckets add P + P (doubling branch), P + (-P) (back to infinity) and infinity itself, in the accumulation
    and in every level of the bucket reduction."""
    cv = po.CURVES[curve]
    g2, g3 = po.g1_mul(cv, cv.g1, 2), po.g1_mul(cv, cv.g1, 3)
    pts = ([cv.g1] * 70 + [po.g1_neg(cv, cv.g1)] * 50 + [g2] * 33 + [po.g1_neg(cv, g2)] * 33 + [None] * 9 + [g3] * 5)
    rng = random.Random(9)
    rng.shuffle(pts)
    srs = api.SRS.from_points(curve, pts)
    n = len(pts)
    cases = [[7] * n, [1] * n, [cv.r - 1] * n, [rng.choice((1, 2, 3, cv.r - 2)) for _ in range(n)],
             H.scalars_uniform(cv.r, n, 6), [5 if P == cv.g1 else 0 for P in pts]]
    for sc in cases:
        assert srs.msm(sc) == po.msm_naive(cv, pts, sc)
    srs.free()


def test_msm_errors(gpu):
    srs = api.SRS.unsafe("BN254", 8, H.TAU)
    with pytest.raises(gpu.B200PlonkError, match="more scalars"):
        srs.msm([1] * 9)
    with pytest.raises(gpu.B200PlonkError, match="power-of-two"):
        srs.msm([1] * 3, basis=_lib.BASIS_LAGRANGE)
    srs.free()
