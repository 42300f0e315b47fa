"""CPU tests of the oracles (no GPU): pins the restatement against everything the
reference holds for this path -- SRS known-answer points
(/root/reference/setup/trusted_setup_test.go), proof layout
(/root/reference/bsb22_test.go:46-123), accept / reject behaviour of the verifier
templates (/root/reference/testutils/verifier_integration_test.go:175-230) -- and the
C++ oracle against the big-integer one byte for byte."""
import random

import pytest

import helpers as H
from oracle import cpu_oracle as co
from oracle import plonk_oracle as po

CURVES = ("BN254", "BLS12_381")

# /root/reference/setup/trusted_setup_test.go:53-59 (Dusk, first five G1), :132 (Dusk G1[32767]),
# :184-189 (Ethereum KZG ceremony, first five G1), :256 (Ethereum G1[32767])
REF_KAT = {
    "DuskBLS12_381": (
        ["97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
         "b00634601c69f919549b3a284249cee53e7be4cf96f05a8ceceec6c74a0a2bc5cb21a6b060bff4c11eb3382c0ca98325",
         "b74090b6c7ad1daee5dcea5a70e0e0dd774b8308fe7e0084031ee84457de0d6f665cd55d81cbb5f650b66eccf6e4cd31",
         "81704091e4770cdf58699eca0569d99d6d001b4191d380e8dd26feddcacafa9d2425f834fe61061af349fd18facc3744",
         "b3f9536dba87c1b6bf29142b58f1760ebdbea7a5a5c92c2e2a221e903937d3da6098ebf1125cc1c8de82f234ad698001"],
        "872d03410917bae2f3536a750ad10f7b7173d89fa27026afe5c0b2e08697eed244a6798ed833826198628058e3c8b0d9"),
    "EethereumKzgCeremonyBLS12_381": (
        ["97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
         "abb83706b7f96c1ef21649124cd01ac58ec3cf19fbe7ba8e172b5f9e0facb354f3da4877946c24f17411cb551e0c24df",
         "a15cb49e7b66d0c94e46613780adcbe141adf7e2c16ec29e996a6be41c92bfc11bfee4188cbb6bdfe90ef4eb8268f1db",
         "8c5e0672d24677f430d729fc8e96cae3a62b1c67997e88d71600d8e1f1954ec04742d79f804345f8e60d11873d18d0d4",
         "b0feedf1a6c84c6470dcecf26cd95c1258c6c744eb3556ae9e864545d4d4e1c1cb9aaf52265e0df4e0c726b2e9d00045"],
        "b2cd3d87b1af48bb6f3c23d765d6ef21a7c6ca2e5e23b0c4feb20559aaf8b06f69d5a0ff7df5f90f7e3aa0225e7ddff6"),
}


# ---- constants ---------------------------------------------------------------------
@pytest.mark.parametrize("curve", CURVES)
def test_domain_constants(curve):
    cv = po.CURVES[curve]
    assert pow(cv.root, 1 << cv.two_adicity, cv.r) == 1
    assert pow(cv.root, 1 << (cv.two_adicity - 1), cv.r) == cv.r - 1
    assert (cv.r - 1) % (1 << cv.two_adicity) == 0 and ((cv.r - 1) >> cv.two_adicity) & 1
    # the coset shift must be a quadratic non-residue (cosets u<w>, u^2<w> disjoint from <w>)
    assert pow(cv.coset_shift, (cv.r - 1) // 2, cv.r) == cv.r - 1
    assert po.is_on_curve(cv, cv.g1)


# ---- SRS decoding (MSM bases) ---------------------------------------------------------
@pytest.mark.parametrize("name", sorted(REF_KAT))
def test_srs_known_answers(name):
    """trusted_setup_test.go: the embedded files start with the pinned compressed points."""
    ent = H.srs_kat()[name]
    cv = po.CURVES[ent["curve"]]
    first, last = REF_KAT[name]
    raw = bytes.fromhex(ent["first"])
    for i, hx in enumerate(first):
        assert raw[i * 48:(i + 1) * 48].hex() == hx
        P = po.g1_decompress(cv, bytes.fromhex(hx))
        assert po.is_on_curve(cv, P)
        # trusted_setup_test.go:76-80,290-303: X is the encoding with the 3 flag bits cleared
        masked = bytearray.fromhex(hx)
        masked[0] &= 0x1F
        assert P[0] == int.from_bytes(masked, "big")
        assert po.g1_compress(cv, P).hex() == hx
    assert ent["index_32767"] == last
    assert po.is_on_curve(cv, po.g1_decompress(cv, bytes.fromhex(last)))
    # trusted_setup_test.go:84-90,211-217: G1[0] is the generator
    assert po.g1_decompress(cv, bytes.fromhex(first[0])) == cv.g1


def test_srs_bn254_generator_and_slice():
    """trusted_setup_test.go:33-39: PPoT G1[0] is the BN254 generator; every committed point decodes."""
    ent = H.srs_kat()["PerpetualPowersOfTauBN254"]
    assert ent["declared_count"] == 524287
    pts = H.real_srs_points("PerpetualPowersOfTauBN254")
    cv = po.BN254
    assert pts[0] == (1, 2)
    assert all(po.is_on_curve(cv, P) for P in pts)
    raw = bytes.fromhex(ent["first"])
    assert all(po.g1_compress(cv, P) == raw[32 * i:32 * i + 32] for i, P in enumerate(pts))


def test_srs_loader_size_check():
    """setup/setup.go:219-223: a file that is too short is rejected."""
    ent = H.srs_kat()["DuskBLS12_381"]
    blob = (67).to_bytes(4, "big") + bytes.fromhex(ent["first"])
    assert len(po.load_srs_g1(po.BLS12_381, blob, 67)) == 67
    with pytest.raises(ValueError, match="too small"):
        po.load_srs_g1(po.BLS12_381, blob, 68)


def test_srs_real_slice_is_geometric():
    """The slice really is [tau^j]G: e-free check P_{j+1} = tau*P_j cannot be done without tau, but
    the C++ oracle's MSM over it must agree with the big-integer MSM (bases used by the parity tests)."""
    pts = H.real_srs_points("DuskBLS12_381")
    sc = H.scalars_uniform(po.BLS12_381.r, len(pts), 3)
    assert co.msm(1, pts, sc) == po.msm_naive(po.BLS12_381, pts, sc)


# ---- C++ oracle vs big-integer oracle ------------------------------------------------------
@pytest.mark.parametrize("curve", CURVES)
def test_cpp_fields(curve):
    cv = po.CURVES[curve]
    rng = random.Random(11)
    for fid, mod in ((0 if cv.cid == 0 else 2, cv.r), (1 if cv.cid == 0 else 3, cv.p)):
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 64) - 1, 1 << 64, (1 << 128) + 1]
        vals = edge + [rng.randrange(mod) for _ in range(40)]
        for a in vals:
            b = rng.choice(vals)
            assert co.field_mul(fid, a, b) == a * b % mod


@pytest.mark.parametrize("curve", CURVES)
def test_cpp_srs_and_msm(curve):
    cv = po.CURVES[curve]
    n = 70
    srs = po.srs_from_tau(cv, H.TAU, n)
    assert co.srs_from_tau(cv.cid, H.TAU, n) == srs
    rng = random.Random(5)
    cases = [[], [7], [0] * n, [1] * n, [cv.r - 1] * n, H.scalars_uniform(cv.r, n, 1),
             H.scalars_witness_like(cv.r, n, 2), [rng.randrange(cv.r) for _ in range(33)],
             [1 << (16 * i % 250) for i in range(n)], [(1 << 15)] * n, [(1 << 15) + 1] * n]
    for sc in cases:
        assert co.msm(cv.cid, srs[: len(sc)], sc) == po.msm_naive(cv, srs[: len(sc)], sc)
    # with points at infinity among the bases
    pts = list(srs[:10])
    pts[3] = None
    sc = H.scalars_uniform(cv.r, 10, 9)
    assert co.msm(cv.cid, pts, sc) == po.msm_naive(cv, pts, sc)


def test_cpp_msm_chunked_path():
    """n >= 1024 takes the (window, chunk) job split."""
    cv = po.BN254
    n = 1100
    srs = co.srs_from_tau(0, H.TAU, n)
    sc = H.scalars_witness_like(cv.r, n, 4)
    assert co.msm(0, srs, sc) == po.msm_naive(cv, srs, sc)


@pytest.mark.parametrize("curve", CURVES)
def test_cpp_ntt(curve):
    cv = po.CURVES[curve]
    for logn in (0, 1, 2, 5, 9):
        n = 1 << logn
        a = H.scalars_uniform(cv.r, n, logn)
        w = po.domain_generator(cv, n)
        f = co.ntt(cv.cid, a)
        assert f == po.ntt(cv, a, w)
        assert co.ntt(cv.cid, f, inverse=True) == a
        fc = co.ntt(cv.cid, a, coset=True)
        assert fc == po.coset_ntt(cv, a, w, cv.coset_shift)
        assert co.ntt(cv.cid, fc, inverse=True, coset=True) == a


# ---- pairing (oracle/pairing.py: the ec_pairing_check of the generated verifiers) ---------
# setup/trusted_setup_test.go:93-95: the compressed G2 points of the Dusk verifying key
DUSK_G2_HEX = (
    "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
    "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8",
    "8fd840491fe66a0cc60f45930d88a9b562136137f78260648ce6a4bf5d31849f18de090e2644780d2bf6b42e20842276"
    "0fabe7238383b48bd61f25125a0d093306ef5511550312e2c1a9fb985e21ce1bf71b1fb0565c3b54836463eb1f043d48",
)


def test_dusk_vk_bin_matches_the_reference_known_answers():
    from oracle import pairing
    ent = H.srs_kat()["DuskBLS12_381"]
    vk_bin = bytes.fromhex(ent["vk_bin"])
    assert vk_bin[:96].hex() == DUSK_G2_HEX[0] and vk_bin[96:192].hex() == DUSK_G2_HEX[1]
    assert vk_bin[192:].hex() == ent["first"][:96]                    # Vk.G1 == Pk.G1[0] (:118-120)
    cv = po.BLS12_381
    for i, hx in enumerate(DUSK_G2_HEX):
        Q = pairing.g2_decompress(cv, bytes.fromhex(hx))
        assert pairing.g2_on_curve(cv, Q) and Q == H.real_srs_g2("DuskBLS12_381")[i]
        # the X the test rebuilds from the hex: A1 = first half with the 3 flag bits cleared, A0 = second half
        raw = bytes.fromhex(hx)
        assert Q[0] == (int.from_bytes(raw[48:], "big"), int.from_bytes(bytes([raw[0] & 0x1F]) + raw[1:48], "big"))


@pytest.mark.parametrize("name", sorted(H.srs_kat()))
def test_pairing_on_the_ceremony_files(name):
    """e([tau]_1, [1]_2) == e([1]_1, [tau]_2) with every operand read from the reference's pk.bin / vk.bin:
    pins the pairing (and the G2 decoder) on data this repo did not produce; plus bilinearity."""
    from oracle import pairing
    ent = H.srs_kat()[name]
    cv = po.CURVES[ent["curve"]]
    g2_one, g2_tau = H.real_srs_g2(name)
    g1 = H.real_srs_points(name)
    assert pairing.pairing_product_is_one(cv, [(g1[1], g2_one), (po.g1_neg(cv, g1[0]), g2_tau)])
    assert pairing.pairing_product_is_one(cv, [(g1[2], g2_one), (po.g1_neg(cv, g1[1]), g2_tau)])
    assert not pairing.pairing_product_is_one(cv, [(g1[1], g2_one), (g1[0], g2_tau)])
    assert not pairing.pairing_product_is_one(cv, [(g1[2], g2_one), (po.g1_neg(cv, g1[0]), g2_tau)])
    # e(a P, Q) e(-P, Q)^a == 1, as e(aP, Q) * e(-a P, Q)
    a = 0xC0FFEE
    aP = po.g1_mul(cv, g1[0], a)
    assert pairing.pairing_product_is_one(cv, [(aP, g2_tau), (po.g1_neg(cv, po.g1_mul(cv, g1[1], a)), g2_one)])


@pytest.mark.parametrize("case", [c for c in H.golden_proofs() if c["srs"] != "tau"], ids=H.case_id)
def test_real_srs_golden_proofs_pass_the_pairing_check(case):
    """The proofs over the reference's REAL trusted setups (tau unknown) are accepted by the restated verifier
    with an actual pairing against the G2 points of the setup's vk.bin, and rejected when tampered with."""
    c = H.build_case(case)
    cv = c["cv"]
    def vk_point(raw):                          # G1Affine.Marshal(): infinity carries gnark's 0x40 flag on BLS12-381
        return None if raw == bytes([0x40]) + bytes(len(raw) - 1) or not any(raw) else po.g1_from_raw_bytes(cv, raw)
    vkb, nb2 = bytes.fromhex(case["vk"]), 2 * cv.fp_bytes
    vk_pts = [vk_point(vkb[i * nb2:(i + 1) * nb2]) for i in range(8 + case["k"])]
    vk = H.vk_from_points(c["tc"], vk_pts, c["srs"][0], tau=None, g2=H.real_srs_g2(case["srs"]))
    blob, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    assert po.verify_proof(vk, blob, pub)
    bad = bytearray(pub)
    bad[31] ^= 1
    assert not po.verify_proof(vk, blob, bytes(bad))
    nb = 2 * cv.fp_bytes
    swapped = blob[nb:2 * nb] + blob[nb:]                 # first G1 point overwritten with the second
    assert not po.verify_proof(vk, swapped, pub)


# ---- proofs -----------------------------------------------------------------------------
@pytest.mark.parametrize("case", H.golden_proofs(), ids=H.case_id)
def test_golden_proof_python_oracle(case):
    """The committed proofs are what the big-integer oracle produces, and (known-tau cases)
    what the restated reference verifier accepts."""
    c = H.build_case(case)
    if case["n"] > 64:
        pytest.skip("big-integer prover is too slow for this size; covered by the C++ oracle test")
    tr = H.oracle_trace(c["tc"])
    vk = po.setup(tr, c["srs"], tau=c["tau"])
    assert po.vk_transcript_bytes(vk).hex() == case["vk"]
    pf = po.prove(tr, vk, c["srs"], c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"])
    blob = po.marshal_proof(c["cv"], pf)
    assert blob.hex() == case["proof"]
    assert len(blob) == po.proof_size(c["cv"], case["k"])
    if c["tau"] is not None:
        assert po.verify_proof(vk, blob, bytes.fromhex(case["public_inputs"]))


@pytest.mark.parametrize("case", H.golden_proofs(), ids=H.case_id)
def test_golden_proof_cpp_oracle(case):
    c = H.build_case(case)
    tc, cid = c["tc"], c["cv"].cid
    circ = co.Circuit(cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, tc.qcp,
                      tc.commitment_constraint_indexes, co.points_le(cid, c["srs"]))
    vk_bytes = b"".join(po.g1_raw_bytes(c["cv"], P, gnark_infinity_flag=True) for P in circ.vk_points())
    assert vk_bytes.hex() == case["vk"]
    blob = circ.prove(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"])
    circ.free()
    assert blob.hex() == case["proof"]


@pytest.mark.parametrize("case", [c for c in H.golden_proofs() if c["srs"] == "tau"], ids=H.case_id)
def test_proof_layout_and_rejects(case):
    """bsb22_test.go:65-119 (length and tail layout) and verifier_integration_test.go:188-228
    (flip a public-input byte; overwrite the first G1 point with the second) on the restated verifier."""
    c = H.build_case(case)
    cv, k = c["cv"], case["k"]
    blob = bytes.fromhex(case["proof"])
    pub = bytes.fromhex(case["public_inputs"])
    base_words, point_bytes = (24, 64) if cv.cid == 0 else (33, 96)
    assert len(blob) == base_words * 32 + k * 32 + k * point_bytes
    for i in range(k):
        start = (base_words + k) * 32 + i * point_bytes
        assert blob[start:start + point_bytes].hex() == case["bsb22"][i]
    tr = H.oracle_trace(c["tc"])
    vk = po.setup(tr, c["srs"], tau=c["tau"])
    assert po.verify_proof(vk, blob, pub)
    bad_pub = bytearray(pub)
    bad_pub[31] ^= 1
    assert not po.verify_proof(vk, blob, bytes(bad_pub))
    bad = bytearray(blob)
    bad[0:point_bytes] = blob[point_bytes:2 * point_bytes]
    assert not po.verify_proof(vk, bytes(bad), pub)
    # any single claimed value off by one is rejected as well
    off = 6 * point_bytes
    bad = bytearray(blob)
    bad[off + 31] ^= 1
    assert not po.verify_proof(vk, bytes(bad), pub)


def test_unsatisfied_witness_is_detected():
    """A wrong witness makes the numerator indivisible by Z_H: both oracles refuse."""
    from algoplonk_b200 import frontend as fe
    B = fe.basic_circuit("BN254", 3, 4, 5)
    cs, values = B.build(), list(B.values)
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    O[3] = (O[3] + 1) % po.BN254.r
    srs = po.srs_from_tau(po.BN254, H.TAU, tc.n + 3)
    tr = H.oracle_trace(tc)
    vk = po.setup(tr, srs, tau=H.TAU)
    with pytest.raises(ArithmeticError):
        po.prove(tr, vk, srs, L, R, O, list(range(1, 10)))
    circ = co.Circuit(0, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (),
                      co.points_le(0, srs))
    with pytest.raises(ArithmeticError):
        circ.prove(L, R, O, list(range(1, 10)))


def test_cpp_oracle_mid_size_accepts():
    """2^10 squaring chain: the C++ oracle's proof is accepted by the restated reference verifier."""
    from algoplonk_b200 import frontend as fe
    cv = po.BN254
    cs, values = fe.squaring_chain("BN254", 10)
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    assert fe.check_gates(tc, L, R, O)
    srs_le = co.srs_from_tau_bytes(0, H.TAU, tc.n + 3)
    circ = co.Circuit(0, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    blob = circ.prove(L, R, O, H.scalars_uniform(cv.r, 9, 77))
    vk = H.vk_from_points(tc, circ.vk_points(), cv.g1, tau=H.TAU)
    assert po.verify_proof(vk, blob, po.marshal_public_inputs(L[: tc.nb_public]))


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("logn,seed", [(3, 1), (5, 2), (6, 3)])
def test_random_dense_circuits_both_oracles_agree(curve, logn, seed):
    """Random-dense circuits (full-width random values in EVERY selector column, irregular copy
    cycles, three public inputs): the big-integer prover and the C++ prover produce the same bytes, and the
    restated reference verifier accepts them and rejects a wrong public input."""
    from algoplonk_b200 import frontend as fe
    cv = po.CURVES[curve]
    cs, values = fe.random_dense_circuit(curve, logn, seed=seed)
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    assert fe.check_gates(tc, L, R, O)
    blinding = H.scalars_uniform(cv.r, 9, seed)
    srs = po.srs_from_tau(cv, H.TAU, tc.n + 3)
    tr = H.oracle_trace(tc)
    vk = po.setup(tr, srs, tau=H.TAU)
    blob = po.marshal_proof(cv, po.prove(tr, vk, srs, L, R, O, blinding, [], []))
    circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (),
                      co.points_le(cv.cid, srs))
    assert circ.prove(L, R, O, blinding) == blob
    circ.free()
    pub = po.marshal_public_inputs(L[: tc.nb_public])
    assert po.verify_proof(vk, blob, pub)
    bad = bytearray(pub)
    bad[31] ^= 1
    assert not po.verify_proof(vk, blob, bytes(bad))


def test_random_dense_circuit_cpp_oracle_mid_size():
    """2^11 rows, past the single-tile sizes: the C++ prover's proof is accepted by the restated verifier."""
    from algoplonk_b200 import frontend as fe
    cv = po.BLS12_381
    cs, values = fe.random_dense_circuit("BLS12_381", 11, seed=4)
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    srs_le = co.srs_from_tau_bytes(cv.cid, H.TAU, tc.n + 3)
    circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    blob = circ.prove(L, R, O, H.scalars_uniform(cv.r, 9, 5))
    vk = H.vk_from_points(tc, circ.vk_points(), cv.g1, tau=H.TAU)
    circ.free()
    assert po.verify_proof(vk, blob, po.marshal_public_inputs(L[: tc.nb_public]))


@pytest.mark.parametrize("curve", CURVES)
def test_no_public_inputs_and_zero_blinding(curve):
    """Edge of the input space: a circuit without public inputs (empty public-input bytes, PI(zeta) = 0) proved
    with all-zero blinding scalars (which puts a proof commitment at infinity): both oracles agree, the restated
    verifier accepts (except where the reference BLS12-381 template itself cannot, see below) and a flipped
    byte of a claimed value is rejected."""
    from algoplonk_b200 import frontend as fe
    cv = po.CURVES[curve]
    B = fe.Builder(curve)
    x = B.secret(3)
    z = B.add(B.mul(x, x), x)
    B.assert_is_equal(z, B.secret(12))
    B.assert_is_different_from_zero(z)
    cs = B.build()
    tc = fe.build_trace(cs)
    assert tc.nb_public == 0 and tc.n == 4
    L, R, O = fe.solve_lro(cs, B.values, tc.n)
    assert fe.check_gates(tc, L, R, O)
    srs = po.srs_from_tau(cv, H.TAU, tc.n + 3)
    tr = H.oracle_trace(tc)
    vk = po.setup(tr, srs, tau=H.TAU)
    for blinding in ([0] * 9, list(range(1, 10))):
        blob = po.marshal_proof(cv, po.prove(tr, vk, srs, L, R, O, blinding, [], []))
        circ = co.Circuit(cv.cid, tc.n, 0, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), co.points_le(cv.cid, srs))
        assert circ.prove(L, R, O, blinding) == blob
        circ.free()
        # Without blinding the quotient is short here and one of the three H commitments is the point at infinity.  The
        # BLS12-381 template hashes an all-zero point with flag 0x80 (templateLogicSigBLS12_381.go:401-407) where
        # gnark's own transcript uses 0x40 (verifier/verifier.go:95-99): the generated AVM verifier then derives
        # other challenges than the prover and rejects -- a quirk of the reference that blinded proofs never
        # meet (footnote), restated faithfully.  BN254 has no flag byte and accepts.
        infinity_in_proof = blinding[0] == 0
        assert po.verify_proof(vk, blob, b"") == (not (infinity_in_proof and curve == "BLS12_381"))
        bad = bytearray(blob)
        bad[6 * 2 * cv.fp_bytes + 5] ^= 1          # inside l(zeta), the first claimed value (Appendix B)
        assert not po.verify_proof(vk, bytes(bad), b"")


def test_hash_to_field_is_rfc9380_expand_message_xmd():
    """The BSB22 challenge hash (templateLogicSigBN254.go:386-397, gnark's fr.Hash with DST "BSB22-Plonk") as
    restated in the oracle is RFC 9380 expand_message_xmd(SHA-256) with 48 output bytes, reduced mod r.  The generic
    expander written here is pinned on the RFC's own known answers (Appendix K.1), so the restatement is anchored on
    a published vector, not only on itself."""
    import hashlib

    def xmd(msg: bytes, dst: bytes, n: int) -> bytes:
        ell = (n + 31) // 32
        dst_prime = dst + bytes([len(dst)])
        b0 = hashlib.sha256(bytes(64) + msg + n.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
        blocks = [hashlib.sha256(b0 + b"\x01" + dst_prime).digest()]
        for i in range(2, ell + 1):
            blocks.append(hashlib.sha256(bytes(x ^ y for x, y in zip(b0, blocks[-1])) + bytes([i]) + dst_prime).digest())
        return b"".join(blocks)[:n]

    quux = b"QUUX-V01-CS02-with-expander-SHA256-128"
    assert xmd(b"", quux, 32).hex() == "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"
    assert xmd(b"abc", quux, 32).hex() == "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"
    assert xmd(b"", quux, 128).hex().startswith("af84c27ccfd45d41914fdff5df25293e221afc53d8ad2ac06d5e3e29485dadbe")
    for cv in (po.BN254, po.BLS12_381):
        for msg in (bytes(2 * cv.fp_bytes), po.g1_raw_bytes(cv, cv.g1), b"\x40" + bytes(95)):
            assert po.hash_fr(cv, msg) == int.from_bytes(xmd(msg, b"BSB22-Plonk", 48), "big") % cv.r


@pytest.mark.parametrize("name", ["PerpetualPowersOfTauBN254", "DuskBLS12_381", "EethereumKzgCeremonyBLS12_381"])
def test_cpp_decompress_equals_the_pinned_python_decoder(name):
    """ora_g1_decompress (the checker of b2p_srs_load_compressed at config sizes) against plonk_oracle.g1_decompress,
    which test_srs_known_answers pins on setup/trusted_setup_test.go's vectors; bad streams are reported by index."""
    ent = H.srs_kat()[name]
    cv = po.CURVES[ent["curve"]]
    raw = bytes.fromhex(ent["first"]) + bytes.fromhex(ent["index_32767"])
    want = [po.g1_decompress(cv, raw[i:i + cv.fp_bytes]) for i in range(0, len(raw), cv.fp_bytes)]
    assert co.points_from_le(cv.cid, co.g1_decompress_bytes(cv.cid, raw)) == want
    inf = po.g1_compress(cv, None)
    assert co.points_from_le(cv.cid, co.g1_decompress_bytes(cv.cid, inf + raw[:cv.fp_bytes])) == [None, want[0]]
    bad = bytearray(raw[: 3 * cv.fp_bytes])
    bad[2 * cv.fp_bytes] &= 0x1F                       # flag bits cleared: not a compressed point
    with pytest.raises(ValueError, match="point 2"):
        co.g1_decompress_bytes(cv.cid, bytes(bad))
    # an x that is not on the curve
    for x in range(2, 50):
        if po.fp_sqrt(cv, (x ** 3 + cv.b) % cv.p) is None:
            break
    off = bytearray(x.to_bytes(cv.fp_bytes, "big"))
    off[0] |= 0x80
    with pytest.raises(ValueError, match="point 1"):
        co.g1_decompress_bytes(cv.cid, raw[:cv.fp_bytes] + bytes(off))


def test_config2_golden_on_the_real_ppot_bytes():
    """tests/golden/ppot_bn254_first_131075.bin + config2_ppot_2p17.json (tools/gen_ppot_slice.py): the fixture's
    first points are the ones srs_kat.json copied from the reference file, the C++ oracle reproduces the committed
    proof, and the library's plonk.Verify accepts it with the G2 points of the setup's own vk.bin (host code, no GPU)."""
    import hashlib
    import json
    import os
    from algoplonk_b200 import api, frontend as fe
    with open(os.path.join(H.GOLDEN, "config2_ppot_2p17.json")) as f:
        gold = json.load(f)
    with open(os.path.join(H.GOLDEN, "ppot_bn254_first_131075.bin"), "rb") as f:
        pk_bin = f.read()
    assert hashlib.sha256(pk_bin).hexdigest() == gold["srs_sha256"]
    count = gold["srs_points"]
    assert int.from_bytes(pk_bin[:4], "big") == count == (1 << 17) + 3 and len(pk_bin) == 4 + 32 * count
    ent = H.srs_kat()["PerpetualPowersOfTauBN254"]
    assert pk_bin[4:4 + 32 * ent["count"]].hex() == ent["first"]
    assert pk_bin[4 + 32 * 32767:4 + 32 * 32768].hex() == ent["index_32767"]
    pts_le = co.g1_decompress_bytes(0, pk_bin[4:])
    cs, values = fe.squaring_chain("BN254", gold["log2"], x0=gold["x0"])
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    circ = co.Circuit(0, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), pts_le)
    vk_pts = circ.vk_points()
    assert co.points_le(0, vk_pts).hex() == gold["vk_points_le"]
    proof = circ.prove(L, R, O, gold["blinding"])
    circ.free()
    assert proof.hex() == gold["proof"]
    g2, g1 = api.kzg_vk_load("BN254", bytes.fromhex(ent["vk_bin"]))
    pub = bytes.fromhex(gold["public_inputs"])
    api.verify("BN254", tc.n, tc.nb_public, [], api.points_to_mont_bytes("BN254", vk_pts), g1, g2, proof, pub)
    with pytest.raises(ValueError):
        api.verify("BN254", tc.n, tc.nb_public, [], api.points_to_mont_bytes("BN254", vk_pts), g1, g2,
                   proof[:100] + bytes([proof[100] ^ 1]) + proof[101:], pub)


@pytest.mark.parametrize("curve", ["BN254", "BLS12_381"])
def test_merkle_mimc_circuit_of_the_reference_example(curve):
    """examples/merkle/logicsigVerifier/main.go: MiMC (Keccak-derived round constants, x^5, Miyaguchi-Preneel), the
    depth-16 tree with six leaves "leaf<i>", proof for the fourth.  The circuit's public root equals the root main.go
    computes by hand (:70-92); a wrong index or sibling violates a gate; the C++ oracle's proof is accepted by the
    restated AVM verifier and by the library's plonk.Verify, and is rejected for another root."""
    from algoplonk_b200 import api, frontend as fe
    cv = po.CURVES[curve]
    assert fe.keccak256_legacy(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert fe.keccak256_legacy(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    assert fe.keccak256_legacy(bytes(135))[:4] != fe.keccak256_legacy(bytes(136))[:4]      # padding at the rate boundary
    B, root = fe.merkle_circuit(curve)
    Hh = lambda *xs: fe.mimc_hash(curve, xs)
    leaves = [int.from_bytes(b"leaf%d" % i, "big") % cv.r for i in range(6)]
    z = [Hh(0)]
    for _ in range(16):
        z.append(Hh(z[-1], z[-1]))
    path = [leaves[3], Hh(leaves[2]), Hh(Hh(leaves[0]), Hh(leaves[1])), Hh(Hh(Hh(leaves[4]), Hh(leaves[5])), z[1])] \
        + [z[i - 1] for i in range(4, 17)]
    want = Hh(path[2], Hh(path[1], Hh(path[0])))
    for i in range(3, 17):
        want = Hh(want, path[i])
    assert root == want
    cs = B.build()
    tc = fe.build_trace(cs)
    assert tc.nb_public == 1 and tc.n == 1 << 14
    L, R, O = fe.solve_lro(cs, B.values, tc.n)
    assert fe.check_gates(tc, L, R, O)
    B2, _ = fe.merkle_circuit(curve, index=2)            # another leaf: another witness, the same root
    assert B2.values[0] == root
    bad = list(B.values)
    bad[3] = (bad[3] + 1) % cv.r                         # a sibling of the path
    assert not fe.check_gates(tc, *fe.solve_lro(cs, bad, tc.n))
    srs_le = co.srs_from_tau_bytes(cv.cid, H.TAU, tc.n + 3)
    circ = co.Circuit(cv.cid, tc.n, 1, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    blob = circ.prove(L, R, O, list(range(1, 10)))
    vk_pts = circ.vk_points()
    circ.free()
    vk = H.vk_from_points(tc, vk_pts, cv.g1, tau=H.TAU)
    pub = po.marshal_public_inputs([root])
    assert po.verify_proof(vk, blob, pub)
    assert not po.verify_proof(vk, blob, po.marshal_public_inputs([root + 1]))
    g1 = api.points_to_mont_bytes(curve, [cv.g1])
    api.verify(curve, tc.n, 1, [], api.points_to_mont_bytes(curve, vk_pts), g1, api.g2_unsafe(curve, H.TAU), blob, pub)


@pytest.mark.parametrize("curve,name", [("BN254", "PerpetualPowersOfTauBN254"), ("BLS12_381", "DuskBLS12_381")])
def test_oracle_g2_group_law(curve, name):
    """oracle/pairing.py's affine G2 law (the checker of b2p_msm_g2): closed under the twist equation, the ceremony's
    G2 points have order r, and [tau]_2 from the library's independent host code (Jacobian, 64-bit limbs) agrees."""
    from algoplonk_b200 import api
    from oracle import pairing as opair
    cv = po.CURVES[curve]
    g2 = H.real_srs_g2(name)
    for Q in g2:
        assert opair.g2_on_curve(cv, Q)
        assert opair.g2_mul(cv, Q, cv.r) is None
        assert opair.g2_mul(cv, Q, cv.r + 1) == Q
        assert opair.g2_add(cv, Q, opair.g2_neg(cv, Q)) is None
    a, b = 0x1234567, 0xABCDEF0123
    P, Q = g2
    s = opair.g2_add(cv, opair.g2_mul(cv, P, a), opair.g2_mul(cv, Q, b))
    assert opair.g2_on_curve(cv, s) and s == opair.g2_msm_naive(cv, [Q, P], [b, a])
    assert opair.g2_add(cv, opair.g2_add(cv, P, Q), Q) == opair.g2_add(cv, P, opair.g2_add(cv, Q, Q))
    for tau in (2, 12345, cv.r - 1):
        assert api.g2_from_mont_bytes(curve, api.g2_unsafe(curve, tau))[1] == opair.g2_mul(cv, P, tau)


@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_oracle_solver_against_the_eager_front_end(curve):
    """oracle/solver.py (gnark's solving rule, the checker of b2p_solver_solve) reproduces the witness the front end
    computed while it built the circuit -- two independent routes to the same assignment -- and reports broken inputs."""
    from algoplonk_b200 import frontend as fe
    from oracle import solver as osolver
    cv = po.CURVES[curve]
    B = fe.basic_circuit(curve)
    M, _ = fe.merkle_circuit(curve, depth=2)
    for cs, values in ((B.build(), B.values), (M.build(), M.values), fe.squaring_chain(curve, 8, x0=3),
                       fe.random_dense_circuit(curve, 8, seed=1), fe.wide_mimc_circuit(curve, 9, 3)):
        inputs = [values[v] for v in cs.input_vars]
        val, levels = osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, inputs)
        assert val == [v % cv.r for v in values]
        tc = fe.build_trace(cs)
        assert fe.check_gates(tc, *fe.solve_lro(cs, val, tc.n))
        assert [fe.solver_wires(cs, tc.n)[k][cs.nb_public + j] for k in range(3) for j in (0,)] == list(cs.constraints[0][5:8])
    cs = B.build()
    with pytest.raises(osolver.Unsatisfied):
        osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, [3, 4, 6])
    assert max(osolver.solve(cv.r, 1, *(lambda c, v: (c.nb_variables, c.constraints, c.input_vars, [v[0], v[1]]))(*fe.squaring_chain(curve, 6)))[1]) == 62


@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_oracle_solver_with_hints(curve):
    """Hints in the oracle's solver (the checker of b2p_solver_create_hinted): the NBits hint behind ToBinary, a hint
    whose input is another hint's output, and the BSB22 commitment hint with its unchecked rows -- against the witness
    the eager front end computed."""
    from algoplonk_b200 import api, frontend as fe
    from oracle import solver as osolver
    cv = po.CURVES[curve]
    M, _ = fe.merkle_circuit(curve, depth=2, bits_from_hint=True)
    B = fe.Builder(curve)
    x = B.public(0b1011001)
    bits = B.to_binary(B.mul(x, B.secret(3)), 12)
    low = B.to_binary(B.hint(fe.HINT_NBITS, [bits[2]], [B.values[bits[2]] & 1])[0], 1)
    B.assert_is_equal(low[0], bits[2])
    for Bd in (M, B):
        cs, values = Bd.build(), Bd.values
        hs = [(h.id, h.in_vars, h.out_vars) for h in cs.hints]
        inputs = [values[v] for v in cs.input_vars]
        val, _ = osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, inputs, hs, api.std_hint_fn())
        assert val == [v % cv.r for v in values]
        with pytest.raises((ValueError, TypeError)):
            osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, inputs)     # hints not given
    pts = po.srs_from_tau(cv, H.TAU, 16)
    commit = lambda col: po.msm_naive(cv, pts, po.intt(cv, col, po.domain_generator(cv, len(col))))
    cs, values, pi2s, coms = H.build_bsb22(curve, 2, commit)
    assert sorted(cs.unchecked_rows) == sorted(r for c in cs.commitments for r in c.committed_rows + [c.commitment_row])
    hs = [(h.id, h.in_vars, h.out_vars) for h in cs.hints]
    extra = lambda hid, vals, n_out: [po.hash_fr(cv, po.fs_point(cv, coms[hid - fe.HINT_BSB22]))]
    inputs = [values[v] for v in cs.input_vars]
    val, _ = osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, inputs, hs,
                           api.std_hint_fn(extra), cs.unchecked_rows)
    assert val == [v % cv.r for v in values]
    with pytest.raises(osolver.Unsatisfied):      # without the exemption the committed rows read  -v + 0 = 0
        osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, inputs, hs, api.std_hint_fn(extra))
