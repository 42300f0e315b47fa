"""plonk.Verify and its pairing check in the library's host arithmetic (csrc/verify_host.hpp, pairing_host.hpp;
b2p_verify / b2p_pairing_check / b2p_g2_generate_unsafe) -- the step right after plonk.Prove in
(*CompiledCircuit).Verify (/root/reference/algoplonk.go:93).  No GPU on this path, so all of it runs here.

Pinned on data this repo did not produce: the ceremony files' own points (e([tau]_1,[1]_2) = e([1]_1,[tau]_2) with
every operand from setup/<name>/pk.bin / vk.bin), whose G2[0] must be the generator the library hard-codes; and on
the golden proofs, which the oracle's line-by-line restatement of the reference's AVM verifier accepts.
"""
import random
from math import gcd

import pytest

import helpers as H
from algoplonk_b200 import _lib, api
from oracle import pairing as opair
from oracle import plonk_oracle as po

CURVES = ("BN254", "BLS12_381")
REAL = {"BN254": "PerpetualPowersOfTauBN254", "BLS12_381": "DuskBLS12_381"}


def test_final_exponentiation_chains_are_integer_identities():
    """The two x-chains pairing_host.hpp evaluates equal (p^4 - p^2 + 1)/r (BN254) / 3 times it (BLS12-381, 3 prime to r)."""
    x = 4965661367192848881
    p = 36 * x**4 + 36 * x**3 + 24 * x**2 + 6 * x + 1
    r = 36 * x**4 + 36 * x**3 + 18 * x**2 + 6 * x + 1
    assert (p, r) == (po.CURVES["BN254"].p, po.CURVES["BN254"].r)
    d, rem = divmod(p**4 - p**2 + 1, r)
    assert rem == 0
    e = (-2 - 18 * x - 30 * x**2 - 36 * x**3) + (1 - 12 * x - 18 * x**2 - 36 * x**3) * p + (1 + 6 * x**2) * p**2 + p**3
    assert e == d                                                           # the BN hard part, exactly
    assert (6 * x * x - p) % r == 0 and (6 * x * x).bit_length() == 127        # the Miller loop length t - 1
    x = -0xD201000000010000
    p = (x - 1) ** 2 * (x**4 - x**2 + 1) // 3 + x
    r = x**4 - x**2 + 1
    assert (p, r) == (po.CURVES["BLS12_381"].p, po.CURVES["BLS12_381"].r)
    d, rem = divmod(p**4 - p**2 + 1, r)
    assert rem == 0
    assert (x - 1) ** 2 * (x + p) * (x**2 + p**2 - 1) + 3 == 3 * d and gcd(3, r) == 1
    assert (x - p) % r == 0


def _g1(curve, k):
    cv = po.CURVES[curve]
    return po.g1_mul(cv, cv.g1, k % cv.r) if k % cv.r else None


@pytest.mark.parametrize("curve", CURVES)
def test_g2_generator_is_the_ceremonys_and_on_the_twist(curve):
    cv = po.CURVES[curve]
    gen, tau_g2 = api.g2_from_mont_bytes(curve, api.g2_unsafe(curve, 12345))
    assert opair.g2_on_curve(cv, gen) and opair.g2_on_curve(cv, tau_g2)
    assert gen == H.real_srs_g2(REAL[curve])[0]          # vk.bin of the reference's setup: G2[0] = [1]_2
    # tau = 1 gives the generator twice, tau = 0 infinity
    assert api.g2_from_mont_bytes(curve, api.g2_unsafe(curve, 1)) == [gen, gen]
    assert api.g2_from_mont_bytes(curve, api.g2_unsafe(curve, 0)) == [gen, None]


@pytest.mark.parametrize("curve", CURVES)
def test_pairing_is_bilinear_and_non_degenerate(curve):
    cv = po.CURVES[curve]
    rng = random.Random(11)
    for _ in range(2):
        a, b = rng.randrange(1, cv.r), rng.randrange(1, cv.r)
        g2 = api.g2_unsafe(curve, b)                      # [1]_2, [b]_2
        q_one, q_b = g2[: len(g2) // 2], g2[len(g2) // 2:]
        P = lambda k: api.points_to_mont_bytes(curve, [_g1(curve, k)])
        # e(a G1, b G2) e(-ab G1, G2) == 1
        assert api.pairing_check(curve, P(a) + P(-a * b), q_b + q_one)
        # e(G1, b G2) e(-b G1, G2) == 1  (the KZG setup equation)
        assert api.pairing_check(curve, P(1) + P(-b), q_b + q_one)
        # three pairs: e(a G1, G2) e(b G1, G2) e(-(a+b) G1, G2) == 1
        assert api.pairing_check(curve, P(a) + P(b) + P(-(a + b)), q_one * 3)
        # not degenerate / not always true
        assert not api.pairing_check(curve, P(1), q_one)
        assert not api.pairing_check(curve, P(a) + P(-a * b + 1), q_b + q_one)
        assert not api.pairing_check(curve, P(a) + P(-a), q_b + q_one)
    # infinity on either side contributes 1; the empty product is 1
    inf1, inf2 = api.points_to_mont_bytes(curve, [None]), api.g2_to_mont_bytes(curve, [None])
    assert api.pairing_check(curve, inf1, q_one) and api.pairing_check(curve, P(5), inf2)
    assert api.pairing_check(curve, b"", b"")
    # points off their curve are an argument error, not "false"
    bad = api.points_to_mont_bytes(curve, [(1, 1)])
    with pytest.raises(_lib.B200PlonkError):
        api.pairing_check(curve, bad, q_one)
    with pytest.raises(_lib.B200PlonkError):
        api.pairing_check(curve, P(1), api.g2_to_mont_bytes(curve, [((1, 2), (3, 4))]))


@pytest.mark.parametrize("name", ["PerpetualPowersOfTauBN254", "DuskBLS12_381", "EethereumKzgCeremonyBLS12_381"])
def test_pairing_on_the_reference_setups_own_points(name):
    """e([tau]_1, [1]_2) == e([1]_1, [tau]_2), every operand read from the ceremony files
    (setup/trusted_setup_test.go checks its setups with the same equation)."""
    if name not in H.srs_kat():
        pytest.skip("setup not in the committed fixture")
    ent = H.srs_kat()[name]
    curve = ent["curve"]
    cv = po.CURVES[curve]
    pts = H.real_srs_points(name)
    g2 = api.g2_to_mont_bytes(curve, H.real_srs_g2(name))
    half = len(g2) // 2
    g1s = api.points_to_mont_bytes(curve, [pts[1], po.g1_neg(cv, pts[0])])
    assert api.pairing_check(curve, g1s, g2[:half] + g2[half:])
    # consecutive powers too: e([tau^3]_1, [1]_2) == e([tau^2]_1, [tau]_2); and a wrong pairing of them fails
    assert api.pairing_check(curve, api.points_to_mont_bytes(curve, [pts[3], po.g1_neg(cv, pts[2])]), g2)
    assert not api.pairing_check(curve, api.points_to_mont_bytes(curve, [pts[3], po.g1_neg(cv, pts[1])]), g2)


@pytest.mark.parametrize("curve", CURVES)
def test_library_pairing_agrees_with_the_oracles_pairing(curve):
    """Same yes/no as oracle/pairing.py (an unrelated construction: Fp12 as plain polynomials, E(Fp12), plain
    final exponentiation) on a true and a false product."""
    cv = po.CURVES[curve]
    a = 0xC0FFEE
    g2 = api.g2_unsafe(curve, a)
    Q = api.g2_from_mont_bytes(curve, g2)
    for k, want in ((-a, True), (-a + 1, False)):
        pairs = [(cv.g1, Q[1]), (_g1(curve, k), Q[0])]
        assert opair.pairing_product_is_one(cv, pairs) is want
        assert api.pairing_check(curve, api.points_to_mont_bytes(curve, [p for p, _ in pairs]),
                                 api.g2_to_mont_bytes(curve, [q for _, q in pairs])) is want


# ---- plonk.Verify ----------------------------------------------------------------------------------------------
def _verify_args(case):
    """(curve, n, nb_public, commitment indexes, vk points raw, g1 raw, g2 raw) of a golden case."""
    c = H.build_case(case)
    cv, tc, curve = c["cv"], c["tc"], case["curve"]
    nb = 2 * cv.fp_bytes
    vkb = bytes.fromhex(case["vk"])
    vk_pts = []
    for i in range(0, len(vkb), nb):
        chunk = bytearray(vkb[i:i + nb])
        chunk[0] &= 0x1F if curve == "BLS12_381" else 0x3F          # gnark's infinity flag
        vk_pts.append(po.g1_from_raw_bytes(cv, bytes(chunk)))
    if c["tau"] is not None:
        g1, g2 = cv.g1, api.g2_unsafe(curve, c["tau"])
    else:
        g1, g2 = c["srs"][0], api.g2_to_mont_bytes(curve, H.real_srs_g2(case["srs"]))
    return (curve, tc.n, tc.nb_public, list(tc.commitment_constraint_indexes),
            api.points_to_mont_bytes(curve, vk_pts), api.points_to_mont_bytes(curve, [g1]), g2), c, vk_pts


@pytest.mark.parametrize("case", H.golden_proofs(), ids=H.case_id)
def test_golden_proofs_are_accepted_and_tampered_ones_rejected(case):
    args, c, vk_pts = _verify_args(case)
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    api.verify(*args, proof, pub)                                     # accepted: no exception
    cv = c["cv"]
    pb = 2 * cv.fp_bytes
    k = case.get("k", 0)
    # one flipped bit in every field of the proof (testutils/verifier_integration_test.go:188-228 tampers the same way)
    fields, off = [], 0
    for size in [pb] * 6 + [32] * 5 + [pb, 32, pb, pb] + [32] * k + [pb] * k:
        fields.append((off, size))
        off += size
    assert off == len(proof)
    for off, size in fields:
        bad = bytearray(proof)
        bad[off + size - 1] ^= 1
        with pytest.raises(ValueError, match="error verifying proof"):
            api.verify(*args, bytes(bad), pub)
    # wrong public input, wrong lengths
    if pub:
        bad = bytearray(pub)
        bad[-1] ^= 1
        with pytest.raises(ValueError, match="pairing"):
            api.verify(*args, proof, bytes(bad))
        with pytest.raises(ValueError, match="public inputs have the wrong length"):
            api.verify(*args, proof, pub[:-32])
    with pytest.raises(ValueError, match="wrong length"):
        api.verify(*args, proof + b"\0", pub)
    # an evaluation that is not reduced mod r, a coordinate that is not reduced mod p
    bad = bytearray(proof)
    bad[6 * pb:6 * pb + 32] = (int.from_bytes(proof[6 * pb:6 * pb + 32], "big") + cv.r).to_bytes(32, "big")
    with pytest.raises(ValueError, match="not reduced"):
        api.verify(*args, bytes(bad), pub)
    bad = bytearray(proof)
    bad[: cv.fp_bytes] = b"\xff" * cv.fp_bytes
    with pytest.raises(ValueError, match="not a point"):
        api.verify(*args, bytes(bad), pub)
    # infinity encodings: all zero on both curves; on BLS12-381 also RawBytes()' 0x40 flag (helper.go:35) -- a point
    # at infinity is a well-formed point (the proof then fails at the pairing), any other flag byte is not
    for enc_ok, first in ((True, 0x00), (curve_is_bls := case["curve"] == "BLS12_381", 0x40), (False, 0x80)):
        bad = bytearray(proof)
        bad[:pb] = bytes([first]) + bytes(pb - 1)
        with pytest.raises(ValueError, match="pairing" if enc_ok else "not a point"):
            api.verify(*args, bytes(bad), pub)
    # BLS12-381: a point on the curve outside the r-torsion subgroup is refused (gnark's decoder checks it; BN254's
    # G1 has cofactor 1)
    if curve_is_bls:
        x = 5
        while po.fp_sqrt(cv, (x ** 3 + cv.b) % cv.p) is None:
            x += 1
        P = (x, po.fp_sqrt(cv, (x ** 3 + cv.b) % cv.p))
        assert po.g1_add(cv, po.g1_mul(cv, P, cv.r - 1), P) is not None     # on the curve, [r] P != infinity
        bad = bytearray(proof)
        bad[:pb] = po.g1_raw_bytes(cv, P)
        with pytest.raises(ValueError, match="not a point"):
            api.verify(*args, bytes(bad), pub)
    # another key: a different Ql commitment, or the wrong G2 pair
    other = list(vk_pts)
    other[3] = po.g1_add(cv, other[3], cv.g1)
    with pytest.raises(ValueError, match="pairing"):
        api.verify(*args[:4], api.points_to_mont_bytes(case["curve"], other), *args[5:], proof, pub)
    with pytest.raises(ValueError, match="pairing"):
        api.verify(*args[:6], api.g2_unsafe(case["curve"], 777), proof, pub)


@pytest.mark.parametrize("curve", CURVES)
def test_verdicts_match_the_restated_reference_verifier(curve):
    """Random single-byte corruptions anywhere in proof or public inputs: b2p_verify and the oracle's restatement
    of the generated AVM verifier give the same verdict (exceptions of the oracle = reject)."""
    case = next(c for c in H.golden_proofs() if c["curve"] == curve and c["name"].startswith("bsb22") and c["srs"] == "tau")
    args, c, vk_pts = _verify_args(case)
    vk = H.vk_from_points(c["tc"], vk_pts, c["cv"].g1, tau=c["tau"])
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    rng = random.Random(5)
    outcomes = set()
    for trial in range(12):
        p2, q2 = bytearray(proof), bytearray(pub)
        if trial:                                                     # trial 0: untouched
            tgt = p2 if trial % 3 else q2
            tgt[rng.randrange(len(tgt))] ^= 1 << rng.randrange(8)
        try:
            want = bool(po.verify_proof(vk, bytes(p2), bytes(q2)))
        except Exception:
            want = False
        try:
            api.verify(*args, bytes(p2), bytes(q2))
            got = True
        except ValueError:
            got = False
        assert got is want, trial
        outcomes.add(got)
    assert outcomes == {True, False}


def test_verify_rejects_impossible_domains():
    case = H.golden_proofs()[0]
    args, _, _ = _verify_args(case)
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    for n, why in ((12, "power of two"), (0, "power of two"), (1, "power of two"), (1 << 40, "2-adicity"),
                   (2 * args[1], "pairing")):
        with pytest.raises(ValueError, match=why):
            api.verify(args[0], n, *args[2:], proof, pub)


def test_verify_argument_errors():
    lib = _lib.load()
    assert lib.b2p_verify(7, 8, 0, 0, None, None, None, None, None, 0, None, 0) == _lib.ERR_ARG
    assert lib.b2p_verify(0, 8, 0, 0, None, None, None, None, None, 0, None, 0) == _lib.ERR_ARG
    assert b"null" in lib.b2p_last_error()


# ---- many proofs, one pairing check ---------------------------------------------------------------------------
@pytest.mark.parametrize("curve", CURVES)
def test_batch_verification_folds_many_proofs_into_one_pairing(curve):
    """b2p_verify_batch: proofs of one circuit with different witnesses / blinding (made by the C++ oracle here --
    the GPU makes the same bytes) are accepted together; one bad proof anywhere in the batch is caught, either
    by its own pre-pairing checks (index reported) or by the folded pairing check ("batch")."""
    from algoplonk_b200 import frontend as fe
    from oracle import cpu_oracle as co
    cv = po.CURVES[curve]
    proofs, pubs, args = [], [], None
    for i in range(5):
        cs, values = fe.squaring_chain(curve, 6, x0=3 + i)
        tc = fe.build_trace(cs)
        L, R, O = fe.solve_lro(cs, values, tc.n)
        if args is None:
            srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, tc.n + 3)
            circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
            args = (curve, tc.n, tc.nb_public, [], api.points_to_mont_bytes(curve, circ.vk_points()),
                    api.points_to_mont_bytes(curve, [cv.g1]), api.g2_unsafe(curve, api.TEST_TAU))
        proofs.append(circ.prove(L, R, O, H.scalars_uniform(cv.r, 9, 100 + i)))
        pubs.append(b"".join((v % cv.r).to_bytes(32, "big") for v in L[: tc.nb_public]))   # helper.go:91-110
    circ.free()
    assert len(set(proofs)) == 5
    for p, q in zip(proofs, pubs):
        api.verify(*args, p, q)
    api.verify_batch(*args, proofs, pubs)
    api.verify_batch(*args, [], [])
    api.verify_batch(*args, proofs[:1], pubs[:1])
    # a proof point moved to another point of the curve passes every pre-pairing check: only the fold sees it
    pb = 2 * cv.fp_bytes
    other = po.g1_raw_bytes(cv, po.g1_mul(cv, cv.g1, 99))
    for victim in (0, 3, 4):
        bad = list(proofs)
        bad[victim] = other + proofs[victim][pb:]
        with pytest.raises(ValueError, match="batch: pairing"):
            api.verify_batch(*args, bad, pubs)
        with pytest.raises(ValueError, match="pairing"):
            api.verify(*args, bad[victim], pubs[victim])
    # two invalid proofs whose errors would cancel under equal weights still fail: proof 1 and 2 swapped
    with pytest.raises(ValueError, match="batch"):
        api.verify_batch(*args, [proofs[0], proofs[2], proofs[1]], pubs[:3])
    # a malformed proof is reported by index
    bad = list(proofs)
    bad[2] = b"\xff" * cv.fp_bytes + proofs[2][cv.fp_bytes:]
    with pytest.raises(ValueError, match="error verifying proof 2: "):
        api.verify_batch(*args, bad, pubs)
    # public inputs belong to their proof
    if pubs[0] != pubs[1]:
        with pytest.raises(ValueError, match="batch"):
            api.verify_batch(*args, proofs, [pubs[1], pubs[0]] + pubs[2:])


# ---- setup/<name>/vk.bin ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["PerpetualPowersOfTauBN254", "DuskBLS12_381", "EethereumKzgCeremonyBLS12_381"])
def test_vk_bin_of_the_reference_setups_decodes_like_the_oracle(name):
    """b2p_kzg_vk_load (srs.Vk.ReadFrom, setup/setup.go:174,190) on the committed bytes of the reference's vk.bin:
    same points as oracle/pairing.py (itself pinned on the Dusk known answers of setup/trusted_setup_test.go:93-95),
    G1 = the curve generator; damaged files are refused with the reason."""
    ent = H.srs_kat()[name]
    curve = ent["curve"]
    cv = po.CURVES[curve]
    vk_bin = bytes.fromhex(ent["vk_bin"])
    g2_raw, g1_raw = api.kzg_vk_load(curve, vk_bin)
    assert tuple(api.g2_from_mont_bytes(curve, g2_raw)) == tuple(H.real_srs_g2(name))
    assert api.points_from_mont_bytes(curve, g1_raw) == [cv.g1]
    nb = cv.fp_bytes
    # the other sign of y: flip the "largest" flag bit of G2[1] -> the negated point
    flipped = bytearray(vk_bin)
    flipped[2 * nb] ^= 0x40 if curve == "BN254" else 0x20
    q = api.g2_from_mont_bytes(curve, api.kzg_vk_load(curve, bytes(flipped))[0])[1]
    want = H.real_srs_g2(name)[1]
    assert q == (want[0], ((-want[1][0]) % cv.p, (-want[1][1]) % cv.p))
    for damage, why in ((lambda b: b[:-1], "wrong length"),
                        (lambda b: bytes([b[0] & (0x3F if curve == "BN254" else 0x1F)]) + b[1:], "invalid flag"),
                        (lambda b: bytes([b[0] | (0x3F if curve == "BN254" else 0x1F)]) + b[1:], "not reduced")):
        with pytest.raises(_lib.B200PlonkError, match=why):
            api.kzg_vk_load(curve, damage(vk_bin))
    # a perturbed x is either off the twist or -- the twist has a cofactor -- on it but outside the r-torsion subgroup
    # (gnark's decoder checks both); it is never accepted
    seen = set()
    for delta in range(1, 24):
        bad = bytearray(vk_bin)
        bad[2 * nb - 1] = (bad[2 * nb - 1] + delta) & 0xFF
        with pytest.raises(_lib.B200PlonkError, match="not on the twist|r-torsion subgroup") as ei:
            api.kzg_vk_load(curve, bytes(bad))
        seen.add("twist" if "not on the twist" in str(ei.value) else "subgroup")
    assert seen == {"twist", "subgroup"}
    # infinity encodings
    inf = bytes([0x40 if curve == "BN254" else 0xC0]) + bytes(2 * nb - 1)
    g2i, _ = api.kzg_vk_load(curve, inf + inf + vk_bin[4 * nb:])
    assert api.g2_from_mont_bytes(curve, g2i) == [None, None]


@pytest.mark.parametrize("case", [c for c in H.golden_proofs() if c["name"] in ("basic", "bsb22_k1")], ids=H.case_id)
def test_compiled_circuit_verifyproof_glue_without_a_gpu(case):
    """api.CompiledCircuit.VerifyProof with the key material it would otherwise fetch from the device already in
    place: the Python glue between the circuit object and b2p_verify (argument order, commitment indexes, G2 of
    the SRS object) -- the part of cc.Verify the CPU suite can reach."""
    args, c, vk_pts = _verify_args(case)
    curve = case["curve"]

    class _Srs:
        g2 = args[6]
    cc = api.CompiledCircuit.__new__(api.CompiledCircuit)
    cc.Ccs, cc.trace, cc.srs, cc.handle, cc.Curve = c["cs"], c["tc"], _Srs(), None, curve
    cc._vk_points, cc._vk_raw, cc._g1_raw = None, args[4], args[5]
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    cc.VerifyProof(proof, pub)
    cc.VerifyProofs([proof] * 3, [pub] * 3)
    bad = bytearray(proof)
    bad[-1] ^= 1
    with pytest.raises(ValueError, match="error verifying proof"):
        cc.VerifyProof(bytes(bad), pub)
    with pytest.raises(ValueError, match="error verifying proof"):
        cc.VerifyProofs([proof, bytes(bad)], [pub] * 2)
    cc.srs.g2 = None
    with pytest.raises(ValueError, match="G2 points are unknown"):
        cc.VerifyProof(proof, pub)


def test_verifier_is_reentrant_across_host_threads():
    """b2p_verify from 6 host threads at once (ctypes drops the GIL), both curves interleaved, good and tampered
    proofs: the shared state -- the per-G2 line cache and the Frobenius constants -- is built once under a lock."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = []
    for case in H.golden_proofs():
        if case["name"] != "basic":
            continue
        args, _, _ = _verify_args(case)
        proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
        bad = bytearray(proof)
        bad[100] ^= 4
        # a G2 pair nobody has prepared yet, so that the first calls race on the cache
        fresh = api.g2_unsafe(case["curve"], 424242)
        jobs += [(args, proof, pub, True), (args, bytes(bad), pub, False),
                 (args[:6] + (fresh,), proof, pub, False)] * 6

    def run(job):
        args, proof, pub, want = job
        try:
            api.verify(*args, proof, pub)
            return want is True
        except ValueError:
            return want is False
    with ThreadPoolExecutor(6) as ex:
        assert all(ex.map(run, jobs))


def test_pairing_internal_identities(tmp_path):
    """tests/csrc/pairing_selfcheck.cpp: cyclotomic squaring == squaring on the cyclotomic subgroup, != off it."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "selfcheck"
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "csrc", "pairing_selfcheck.cpp"), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["bad=0", "bad=0"], out.stdout
