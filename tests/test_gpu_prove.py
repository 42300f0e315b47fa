"""GPU parity: b2p_prove (plonk.Prove, /root/reference/algoplonk.go:89) -- proofs byte-identical to the
oracles given the same blinding scalars, accepted by the restated reference verifier, on the reference's
own circuits (examples/basic, bsb22_test.go), on real-SRS slices and at the BASELINE config sizes."""
import pytest

import helpers as H
from algoplonk_b200 import _lib, api, frontend as fe
from oracle import cpu_oracle as co
from oracle import plonk_oracle as po

pytestmark = pytest.mark.gpu
SETUP = {"BN254": api.SetupName.TestOnlyBN254, "BLS12_381": api.SetupName.TestOnlyBLS12381}


def _compile(c, case):
    curve = case["curve"]
    if case["srs"] == "tau":
        return api.Compile(c["cs"], curve, SETUP[curve])
    real = {"PerpetualPowersOfTauBN254": api.SetupName.PerpetualPowersOfTauBN254,
            "DuskBLS12_381": api.SetupName.DuskBLS12381}[case["srs"]]
    g2, _ = api.kzg_vk_load(curve, bytes.fromhex(H.srs_kat()[case["srs"]]["vk_bin"]))   # the setup's vk.bin: cc.Verify runs b2p_verify
    return api.Compile(c["cs"], curve, real, srs=api.SRS.from_points(curve, c["srs"], g2=g2))


@pytest.mark.parametrize("case", H.golden_proofs(), ids=H.case_id)
def test_golden_proofs_byte_identical(gpu, case):
    c = H.build_case(case)
    cc = _compile(c, case)
    vk_pts = cc.vk_commitments()
    assert b"".join(po.g1_raw_bytes(c["cv"], P, gnark_infinity_flag=True) for P in vk_pts).hex() == case["vk"]
    vp = cc.Verify(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"])
    blob = api.MarshalProof(vp.Proof)
    assert blob.hex() == case["proof"]
    assert api.MarshalPublicInputs(case["curve"], vp.Witness).hex() == case["public_inputs"]
    if c["tau"] is not None:
        vk = H.vk_from_points(c["tc"], vk_pts, c["cv"].g1, tau=c["tau"])
    else:       # real ceremony SRS: the verifier's pairing check against the G2 points of the setup's vk.bin
        vk = H.vk_from_points(c["tc"], vk_pts, c["srs"][0], tau=None, g2=H.real_srs_g2(case["srs"]))
    assert po.verify_proof(vk, blob, bytes.fromhex(case["public_inputs"]))
    cc.free()


@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
@pytest.mark.parametrize("k", (1, 2))
def test_bsb22_commitment_hint_on_gpu(gpu, curve, k):
    """The BSB22 solver hint commits through b2p_msm_g1 on the Lagrange basis (what the Go shim does)."""
    cv = po.CURVES[curve]
    n_dry = fe.bsb22_circuit(curve, k, lambda a, b, c: 1).build().domain_size
    srs = api.SRS.unsafe(curve, n_dry + 3, H.TAU)
    cs, values, pi2s, coms = H.build_bsb22(curve, k, lambda col: srs.msm(col, basis=_lib.BASIS_LAGRANGE))
    case = next(c for c in H.golden_proofs() if c["curve"] == curve and c["name"] == f"bsb22_k{k}")
    assert [po.g1_raw_bytes(cv, P).hex() for P in coms] == case["bsb22"]
    cc = api.Compile(cs, curve, SETUP[curve], srs=srs)
    L, R, O = fe.solve_lro(cs, values, cc.trace.n)
    blob = api.MarshalProof(cc.Prove(L, R, O, case["blinding"], pi2s, coms))
    assert blob.hex() == case["proof"]
    cc.free()
    srs.free()


@pytest.mark.parametrize("curve,logn", [("BN254", 10), ("BN254", 14), ("BLS12_381", 12)])
def test_mid_size_byte_identical_to_cpp_oracle(gpu, curve, logn):
    cv = po.CURVES[curve]
    cs, values = fe.squaring_chain(curve, logn, x0=5)
    cc = api.Compile(cs, curve, SETUP[curve])
    tc = cc.trace
    L, R, O = fe.solve_lro(cs, values, tc.n)
    blinding = H.scalars_uniform(cv.r, 9, logn)
    blob = api.MarshalProof(cc.Prove(L, R, O, blinding))
    srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, tc.n + 3)
    circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    assert cc.vk_commitments() == circ.vk_points()
    assert blob == circ.prove(L, R, O, blinding)
    vk = H.vk_from_points(tc, cc.vk_commitments(), cv.g1, tau=api.TEST_TAU)
    pub = api.MarshalPublicInputs(curve, L[: tc.nb_public])
    assert po.verify_proof(vk, blob, pub)
    # blinding really enters: different scalars, different proof, still accepted
    blob2 = api.MarshalProof(cc.Prove(L, R, O, H.scalars_uniform(cv.r, 9, 999)))
    assert blob2 != blob and po.verify_proof(vk, blob2, pub)
    # proving twice with the same inputs is deterministic
    assert api.MarshalProof(cc.Prove(L, R, O, blinding)) == blob
    circ.free()
    cc.free()


def _many_public_circuit(curve, nb_public):
    """sum of nb_public public inputs == a secret, padded with a few multiplications."""
    B = fe.Builder(curve)
    pubs = [B.public(3 * i + 2) for i in range(nb_public)]
    total = B.secret(sum(3 * i + 2 for i in range(nb_public)))
    acc = pubs[0]
    for v in pubs[1:]:
        acc = B.add(acc, v)
    B.assert_is_equal(acc, total)
    sq = B.mul(total, total)
    B.assert_is_equal(B.mul(sq, pubs[1]), B.secret(B.values[sq] * B.values[pubs[1]]))
    return B


@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
@pytest.mark.parametrize("nb_public", (8, 12))
def test_many_public_inputs(gpu, curve, nb_public):
    """8 public inputs: their Lagrange terms are added by the quotient kernel; 12: the general completeQk
    path (qk completed in Lagrange form, iNTT + coset NTT).  Both byte-identical to the C++ oracle."""
    cv = po.CURVES[curve]
    B = _many_public_circuit(curve, nb_public)
    cs = B.build()
    cc = api.Compile(cs, curve, SETUP[curve])
    tc = cc.trace
    assert tc.nb_public == nb_public
    L, R, O = fe.solve_lro(cs, B.values, tc.n)
    assert fe.check_gates(tc, L, R, O)
    blinding = H.scalars_uniform(cv.r, 9, nb_public)
    blob = api.MarshalProof(cc.Prove(L, R, O, blinding))
    srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, tc.n + 3)
    circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    assert blob == circ.prove(L, R, O, blinding)
    vk = H.vk_from_points(tc, cc.vk_commitments(), cv.g1, tau=api.TEST_TAU)
    assert po.verify_proof(vk, blob, api.MarshalPublicInputs(curve, L[: tc.nb_public]))
    circ.free()
    cc.free()


@pytest.mark.parametrize("name", ("basic", "bsb22_k2"))
def test_direct_public_input_terms_equal_completed_qk(gpu, name, monkeypatch):
    """B2P_NO_PI_DIRECT=1 forces the general path on circuits that normally take the direct one."""
    case = next(c for c in H.golden_proofs() if c["curve"] == "BN254" and c["name"] == name and c["srs"] == "tau")
    c = H.build_case(case)
    cc = _compile(c, case)
    monkeypatch.setenv("B2P_NO_PI_DIRECT", "1")
    general = api.MarshalProof(cc.Prove(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"]))
    monkeypatch.delenv("B2P_NO_PI_DIRECT")
    direct = api.MarshalProof(cc.Prove(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"]))
    assert general.hex() == case["proof"] == direct.hex()
    cc.free()


@pytest.mark.parametrize("curve,logn", [("BN254", 17), ("BN254", 20), ("BLS12_381", 17)])
def test_config_sizes_accepted_by_reference_verifier(gpu, curve, logn):
    """BASELINE configs[1], [2] (2^17 and 2^20 BN254) and a 2^17 BLS12-381 run: size-independent
    check -- the restated AVM verifier accepts, and rejects a tampered public input.  The 2^17 BN254
    case is additionally compared byte for byte with the C++ oracle."""
    cv = po.CURVES[curve]
    cs, values = fe.squaring_chain(curve, logn, x0=7)
    cc = api.Compile(cs, curve, SETUP[curve])
    tc = cc.trace
    assert tc.n == 1 << logn
    L, R, O = fe.solve_lro(cs, values, tc.n)
    blinding = H.scalars_uniform(cv.r, 9, logn)
    blob = api.MarshalProof(cc.Prove(L, R, O, blinding))
    vk = H.vk_from_points(tc, cc.vk_commitments(), cv.g1, tau=api.TEST_TAU)
    pub = api.MarshalPublicInputs(curve, L[: tc.nb_public])
    assert po.verify_proof(vk, blob, pub)
    bad = bytearray(pub)
    bad[31] ^= 1
    assert not po.verify_proof(vk, blob, bytes(bad))
    if (curve, logn) == ("BN254", 17):
        srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, tc.n + 3)
        circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
        assert blob == circ.prove(L, R, O, blinding)
        circ.free()
    cc.free()


@pytest.mark.parametrize("curve,logn", [("BN254", 20), ("BLS12_381", 20), ("BLS12_381", 21)])
def test_benchmark_sizes_byte_identical_to_cpp_oracle(gpu, curve, logn):
    """The sizes bench.py reports on -- 2^20 BN254 (BASELINE configs[2], the headline), 2^20 BLS12-381 and 2^21
    BLS12-381 (configs[4]'s size) -- compared BYTE FOR BYTE with the C++ CPU oracle, verifying-key commitments
    included, with full-width random blinding scalars; then accepted by the restated AVM verifier."""
    cv = po.CURVES[curve]
    cs, values = fe.squaring_chain(curve, logn, x0=13)
    cc = api.Compile(cs, curve, SETUP[curve])
    tc = cc.trace
    assert tc.n == 1 << logn
    L, R, O = fe.solve_lro(cs, values, tc.n)
    blinding = H.scalars_uniform(cv.r, 9, 1000 + logn)
    blob = api.MarshalProof(cc.Prove(L, R, O, blinding))
    vk_pts = cc.vk_commitments()
    cc.free()
    cc.srs.free()
    srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, tc.n + 3)
    circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    assert vk_pts == circ.vk_points()
    want = circ.prove(L, R, O, blinding)
    circ.free()
    assert blob == want
    vk = H.vk_from_points(tc, vk_pts, cv.g1, tau=api.TEST_TAU)
    assert po.verify_proof(vk, blob, api.MarshalPublicInputs(curve, L[: tc.nb_public]))


def test_config2_real_ppot_srs_at_2p17(gpu):
    """BASELINE configs[1]: the 2^17-row circuit on the REAL PerpetualPowersOfTau bytes (the first 2^17 + 3
    compressed points of setup/PerpetualPowersOfTauBN254/pk.bin, committed as a fixture by tools/gen_ppot_slice.py):
    b2p_srs_load_compressed (setup/setup.go:196-228) -> Compile -> cc.Verify, whose plonk.Verify runs the pairing
    check against the G2 points of the setup's own vk.bin; every decompressed point, the verifying key and the
    proof bytes equal the C++ oracle's (live and the committed golden); the restated AVM verifier accepts with
    the ceremony's G2 points."""
    import hashlib
    import json
    import os
    with open(os.path.join(H.GOLDEN, "config2_ppot_2p17.json")) as f:
        gold = json.load(f)
    with open(os.path.join(H.GOLDEN, "ppot_bn254_first_131075.bin"), "rb") as f:
        pk_bin = f.read()
    assert hashlib.sha256(pk_bin).hexdigest() == gold["srs_sha256"]
    count = gold["srs_points"]
    name = "PerpetualPowersOfTauBN254"
    srs = api.SRS.from_pk_bin("BN254", pk_bin, count, vk_bin=bytes.fromhex(H.srs_kat()[name]["vk_bin"]))
    assert srs.size == count
    pts_le = co.g1_decompress_bytes(0, pk_bin[4:])
    assert co.points_le(0, srs.points(0, count)) == pts_le
    cs, values = fe.squaring_chain("BN254", gold["log2"], x0=gold["x0"])
    cc = api.Compile(cs, "BN254", api.SetupName.PerpetualPowersOfTauBN254, srs=srs)
    tc = cc.trace
    L, R, O = fe.solve_lro(cs, values, tc.n)
    vp = cc.Verify(L, R, O, gold["blinding"])                # raises if b2p_verify (real pairing) rejects
    blob = api.MarshalProof(vp.Proof)
    assert co.points_le(0, cc.vk_commitments()) == bytes.fromhex(gold["vk_points_le"])
    assert blob.hex() == gold["proof"]
    circ = co.Circuit(0, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), pts_le)
    assert blob == circ.prove(L, R, O, gold["blinding"])
    circ.free()
    pub = api.MarshalPublicInputs("BN254", vp.Witness)
    assert pub.hex() == gold["public_inputs"]
    vk = H.vk_from_points(tc, cc.vk_commitments(), H.real_srs_points(name)[0], tau=None, g2=H.real_srs_g2(name))
    assert po.verify_proof(vk, blob, pub)
    bad = bytearray(blob)
    bad[700] ^= 1
    with pytest.raises(ValueError):
        cc.VerifyProof(bytes(bad), pub)
    cc.free()
    srs.free()


@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_merkle_mimc_circuit_of_the_reference_example(gpu, curve):
    """The circuit every integration test of the reference proves (examples/merkle/logicsigVerifier/main.go:45-61,
    testutils/verifier_integration_test.go:175-230): MiMC Merkle proof, depth 16, public root.  BN254 on the REAL
    PerpetualPowersOfTau points, as main.go:113 compiles it (cc.Verify then checks the pairing against the setup's
    vk.bin); BLS12-381 on the TestOnly setup.  Proof and verifying key byte for byte the C++ oracle's, accepted by
    the restated AVM verifier, rejected for another root."""
    import os
    cv = po.CURVES[curve]
    B, root = fe.merkle_circuit(curve)
    cs = B.build()
    if curve == "BN254":
        with open(os.path.join(H.GOLDEN, "ppot_bn254_first_131075.bin"), "rb") as f:
            pk_bin = f.read()
        name = "PerpetualPowersOfTauBN254"
        count = (1 << 14) + 3
        srs = api.SRS.from_pk_bin(curve, pk_bin, count, vk_bin=bytes.fromhex(H.srs_kat()[name]["vk_bin"]))
        cc = api.Compile(cs, curve, api.SetupName.PerpetualPowersOfTauBN254, srs=srs)
        srs_le = co.g1_decompress_bytes(0, pk_bin[4:4 + 32 * count])
        g1, tau, g2 = H.real_srs_points(name)[0], None, H.real_srs_g2(name)
    else:
        cc = api.Compile(cs, curve, SETUP[curve])
        srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, cc.trace.n + 3)
        g1, tau, g2 = cv.g1, api.TEST_TAU, None
    tc = cc.trace
    assert tc.n == 1 << 14 and tc.nb_public == 1
    L, R, O = fe.solve_lro(cs, B.values, tc.n)
    blinding = H.scalars_uniform(cv.r, 9, 16)
    vp = cc.Verify(L, R, O, blinding)                         # plonk.Prove + plonk.Verify (b2p_verify)
    assert vp.Witness == [root]
    blob = api.MarshalProof(vp.Proof)
    circ = co.Circuit(cv.cid, tc.n, 1, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    assert cc.vk_commitments() == circ.vk_points()
    assert blob == circ.prove(L, R, O, blinding)
    circ.free()
    vk = H.vk_from_points(tc, cc.vk_commitments(), g1, tau=tau, g2=g2)
    pub = api.MarshalPublicInputs(curve, [root])
    assert po.verify_proof(vk, blob, pub)
    assert not po.verify_proof(vk, blob, api.MarshalPublicInputs(curve, [root + 1]))
    with pytest.raises(ValueError):
        cc.VerifyProof(blob, api.MarshalPublicInputs(curve, [root + 1]))
    cc.free()
    cc.srs.free()


def test_two_proofs_in_flight_on_two_handles(gpu):
    """The library is re-entrant across handles and calls may come from any host thread (fresh threads start
    on CUDA device 0; every entry point switches to its handle's device): two keys, two threads, the same
    bytes as a lone proof."""
    import threading
    curve, cv = "BN254", po.CURVES["BN254"]
    cs, values = fe.squaring_chain(curve, 12, x0=3)
    ccs = [api.Compile(cs, curve, SETUP[curve]) for _ in range(2)]
    L, R, O = fe.solve_lro(cs, values, ccs[0].trace.n)
    blindings = [H.scalars_uniform(cv.r, 9, s) for s in (1, 2)]
    want = [api.MarshalProof(ccs[i].Prove(L, R, O, blindings[i])) for i in range(2)]
    got = [[None] * 4 for _ in range(2)]
    errs = []

    def work(i):
        try:
            for rep in range(4):
                got[i][rep] = api.MarshalProof(ccs[i].Prove(L, R, O, blindings[i]))
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for i in range(2):
        assert all(g == want[i] for g in got[i])
    assert want[0] != want[1]
    for cc in ccs:
        cc.free()


def test_concurrent_proofs_are_deterministic_at_benchmark_size(gpu):
    """3 proving keys, 3 host threads, 2^20 rows: every proof equals the lone reference proof, with resident
    and with host inputs.  (Bucket contents are summed in a different order in every run -- the ranks come
    from atomics -- so a value-dependent arithmetic slip shows up here as a rare mismatch; this caught a
    dropped carry that hit one point addition in 10^9.)"""
    import ctypes as C
    import threading
    curve, lanes, reps = "BN254", 3, 8
    cs, values = fe.squaring_chain(curve, 20, x0=11)
    ccs = [api.Compile(cs, curve, SETUP[curve]) for _ in range(lanes)]
    n = ccs[0].trace.n
    L, R, O = fe.solve_lro(cs, values, n)
    bufs = [C.create_string_buffer(api.fr_to_mont_bytes(curve, col)) for col in (L, R, O)]
    bl = api.fr_to_mont_bytes(curve, list(range(11, 20)))
    ref = ccs[0].prove_raw(bufs[0], bufs[1], bufs[2], bl).raw
    bad, errs = [], []

    def work(i):
        try:
            for k in range(reps):
                if ccs[i].prove_raw(bufs[0], bufs[1], bufs[2], bl).raw != ref:
                    bad.append((i, k))
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(lanes)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    assert not bad, f"proofs differ from the reference proof: {sorted(bad)}"
    for cc in ccs:
        cc.free()


def test_prove_from_pinned_host_buffers(gpu):
    """b2p_host_alloc / b2p_host_free: the page-locked staging the Go shim copies the solver's columns into."""
    import ctypes as C
    lib = _lib.load()
    curve, cv = "BN254", po.CURVES["BN254"]
    cs, values = fe.squaring_chain(curve, 11, x0=9)
    cc = api.Compile(cs, curve, SETUP[curve])
    n = cc.trace.n
    L, R, O = fe.solve_lro(cs, values, n)
    blinding = api.fr_to_mont_bytes(curve, H.scalars_uniform(cv.r, 9, 5))
    want = cc.prove_raw(*(api.fr_to_mont_bytes(curve, c) for c in (L, R, O)), blinding).raw
    ptrs = []
    for col in (L, R, O):
        p = C.c_void_p()
        _lib.check(lib.b2p_host_alloc(32 * n, C.byref(p)))
        C.memmove(p, api.fr_to_mont_bytes(curve, col), 32 * n)
        ptrs.append(p)
    out = C.create_string_buffer(lib.b2p_proof_raw_size(api.CURVE_ID[curve], 0))
    _lib.check(lib.b2p_prove(cc.handle, ptrs[0], ptrs[1], ptrs[2], None, None, C.create_string_buffer(blinding), out))
    assert out.raw == want
    for p in ptrs:
        lib.b2p_host_free(p)
    lib.b2p_host_free(None)
    cc.free()


@pytest.mark.parametrize("case", [c for c in H.golden_proofs() if c["name"] in ("basic", "bsb22_k1", "squaring_2p6")],
                         ids=H.case_id)
def test_circuit_load_with_the_callers_vk_transcript(gpu, case):
    """The Go shim hands b2p_circuit_load the verifying key's commitments as gnark marshals them (pk.Vk.*,
    G1Affine.Marshal(), infinity = 0x40 flag on BLS12-381) instead of letting the library commit: same proofs,
    and [Lin] is then built from the points decoded out of those bytes."""
    c = H.build_case(case)
    curve = case["curve"]
    vkb = bytes.fromhex(case["vk"])
    if case["srs"] == "tau":
        cc = api.Compile(c["cs"], curve, SETUP[curve], vk_transcript=vkb)
    else:
        real = {"PerpetualPowersOfTauBN254": api.SetupName.PerpetualPowersOfTauBN254,
                "DuskBLS12_381": api.SetupName.DuskBLS12381}[case["srs"]]
        cc = api.Compile(c["cs"], curve, real, srs=api.SRS.from_points(curve, c["srs"]), vk_transcript=vkb)
    blob = api.MarshalProof(cc.Prove(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"]))
    assert blob.hex() == case["proof"]
    with pytest.raises(_lib.B200PlonkError, match="wrong length"):
        api.Compile(c["cs"], curve, SETUP[curve] if case["srs"] == "tau" else real,
                    srs=None if case["srs"] == "tau" else api.SRS.from_points(curve, c["srs"]), vk_transcript=vkb[:-1])
    cc.free()


def test_prove_errors(gpu):
    B = fe.basic_circuit("BN254")
    cs = B.build()
    small = api.SRS.unsafe("BN254", 8, H.TAU)           # needs n + 3 = 11
    with pytest.raises(ValueError, match="too small"):  # setup/setup.go:219-223
        api.Compile(cs, "BN254", api.SetupName.TestOnlyBN254, srs=small)
    lib = gpu.load()
    import ctypes as C
    h = C.c_void_p()
    col = C.create_string_buffer(32 * 8)
    perm = (C.c_int64 * 24)(*range(24))
    assert lib.b2p_circuit_load(small.handle, 8, 0, col, col, col, col, col, perm, 0, None, None, None, 0,
                                C.byref(h)) == -1
    assert b"n+3" in lib.b2p_last_error()
    bls = api.SRS.unsafe("BLS12_381", 16, H.TAU)
    perm[0] = 99                                         # out-of-range permutation entry
    assert lib.b2p_circuit_load(bls.handle, 8, 0, col, col, col, col, col, perm, 0, None, None, None, 0,
                                C.byref(h)) == -1
    small.free()
    bls.free()
