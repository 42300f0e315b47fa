"""GPU parity: b2p_srs_load_compressed -- the embedded trusted setups as they sit on disk
(/root/reference/setup/<name>/pk.bin, loaded by setup/setup.go:165-228) decompressed on the GPU, against the
oracle's decoder and the reference's own known answers (setup/trusted_setup_test.go:53-59,132,183-189,256;
committed as tests/golden/srs_kat.json together with the compressed bytes of the leading points)."""
import pytest

import helpers as H
from algoplonk_b200 import _lib, api
from oracle import plonk_oracle as po

pytestmark = pytest.mark.gpu


def _pk_bin(ent, declared=None):
    raw = bytes.fromhex(ent["first"])
    return (int(ent["declared_count"]) if declared is None else declared).to_bytes(4, "big") + raw


@pytest.mark.parametrize("name", sorted(H.srs_kat()))
def test_decompress_matches_oracle_and_reference_kats(gpu, name):
    ent = H.srs_kat()[name]
    curve, count = ent["curve"], int(ent["count"])
    cv = po.CURVES[curve]
    srs = api.SRS.from_pk_bin(curve, _pk_bin(ent), count)
    got = srs.points(0, count)
    assert got == H.real_srs_points(name)                      # oracle decoder, every committed point
    assert got[0] == cv.g1                                     # trusted_setup_test.go:33-39,84-90,211-217
    for P in got:
        assert P is not None and (P[1] * P[1] - P[0] ** 3 - cv.b) % cv.p == 0
    # the decompressed points are usable MSM bases
    sc = H.scalars_uniform(cv.r, count, 4)
    assert srs.msm(sc) == po.msm_naive(cv, got, sc)
    # a shorter prefix (setup.Run truncates to nextPow2 + 3 points, setup.go:113-114)
    sub = api.SRS.from_pk_bin(curve, _pk_bin(ent), 5)
    assert sub.size == 5 and sub.points(0, 5) == got[:5]
    sub.free()
    srs.free()


@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_decompress_signs_infinity_and_errors(gpu, curve):
    cv = po.CURVES[curve]
    pts = [po.g1_mul(cv, cv.g1, k) for k in (1, 2, 3, 5, 7, 11)]
    pts += [po.g1_neg(cv, P) for P in pts] + [None]          # both roots of every x, and infinity
    blob = b"".join(po.g1_compress(cv, P) for P in pts)
    pk = len(pts).to_bytes(4, "big") + blob
    srs = api.SRS.from_pk_bin(curve, pk, len(pts))
    assert srs.points(0, len(pts)) == pts
    srs.free()
    # setup.go:219-223: asking for more points than the file holds
    with pytest.raises(_lib.B200PlonkError, match="pk.bin too small for 99 elements"):
        api.SRS.from_pk_bin(curve, pk, 99)
    with pytest.raises(_lib.B200PlonkError, match="pk.bin too small"):
        api.SRS.from_pk_bin(curve, (99).to_bytes(4, "big") + blob, 99)     # header lies about the payload
    # x with no point on the curve
    x = 1
    while po.fp_sqrt(cv, (x ** 3 + cv.b) % cv.p) is not None:
        x += 1
    bad = bytearray(x.to_bytes(cv.fp_bytes, "big"))
    bad[0] |= 0x80
    with pytest.raises(_lib.B200PlonkError, match="not on the curve"):
        api.SRS.from_pk_bin(curve, (2).to_bytes(4, "big") + po.g1_compress(cv, cv.g1) + bytes(bad), 2)
    # uncompressed flag inside a compressed stream
    with pytest.raises(_lib.B200PlonkError, match="invalid point flag"):
        api.SRS.from_pk_bin(curve, (1).to_bytes(4, "big") + bytes(cv.fp_bytes), 1)
    # coordinate >= p
    big = bytearray(cv.p.to_bytes(cv.fp_bytes, "big"))
    big[0] |= 0x80
    with pytest.raises(_lib.B200PlonkError, match="not reduced"):
        api.SRS.from_pk_bin(curve, (1).to_bytes(4, "big") + bytes(big), 1)


@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
@pytest.mark.parametrize("n", (1, 2, 8, 64, 1024))
def test_to_lagrange_known_tau(gpu, curve, n):
    """b2p_srs_to_lagrange = kzg.ToLagrangeG1 (setup/setup.go:124,138) on a TestOnly SRS, where tau is known:
    [L_j(tau)]_1 with L_j(tau) = (tau^n - 1) w^j / (n (tau - w^j)), computed with integers and one scalar
    multiplication per point by the oracle."""
    cv = po.CURVES[curve]
    srs = api.SRS.unsafe(curve, n + 3, H.TAU)
    got = srs.to_lagrange(n)
    w = po.domain_generator(cv, n) if n > 1 else 1
    tn = (pow(H.TAU, n, cv.r) - 1) % cv.r
    ninv = pow(n, -1, cv.r)
    for j in ([0, 1, n // 2, n - 1] if n > 8 else range(n)):
        wj = pow(w, j, cv.r)
        lj = tn * wj % cv.r * ninv % cv.r * pow((H.TAU - wj) % cv.r, -1, cv.r) % cv.r
        assert got[j] == po.g1_mul(cv, cv.g1, lj), (n, j)
    srs.free()


def test_to_lagrange_on_the_real_ppot_points_commits_like_the_prover(gpu):
    """On ceremony points (tau unknown) the Lagrange SRS is pinned through its purpose: MSM(Lagrange points, v) must be
    the commitment the prover makes for the column v, i.e. MSM(canonical points, iNTT(v)) -- both sides on the oracle's
    big integers from the GPU's table, and against b2p_msm_g1(B2P_BASIS_LAGRANGE)."""
    name, curve, n = "PerpetualPowersOfTauBN254", "BN254", 128
    cv = po.CURVES[curve]
    pts = H.real_srs_points(name)[: n + 3]
    srs = api.SRS.from_points(curve, pts)
    lag = srs.to_lagrange(n)
    v = H.scalars_uniform(cv.r, n, 5)
    coeffs = po.intt(cv, v, po.domain_generator(cv, n))
    want = po.msm_naive(cv, pts[:n], coeffs)
    assert po.msm_naive(cv, lag, v) == want == srs.msm(v, basis=_lib.BASIS_LAGRANGE)
    with pytest.raises(_lib.B200PlonkError, match="power of two"):
        srs.to_lagrange(96)
    with pytest.raises(_lib.B200PlonkError, match="power of two"):
        srs.to_lagrange(256)              # more than the SRS holds
    srs.free()
