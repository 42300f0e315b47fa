"""GPU parity: b2p_verify_batch_dev (plonk.Verify for a batch, /root/reference/algoplonk.go:93, with the point
combinations on the GPU) gives the verdicts of the host verifier -- itself pinned on the restated reference verifier --
on golden proofs, batches of distinct proofs, tampered proofs, and BLS12-381 points outside the r-torsion subgroup."""
import pytest

import helpers as H
import test_verify_host as tvh
from algoplonk_b200 import _lib, api, frontend as fe
from oracle import cpu_oracle as co
from oracle import plonk_oracle as po

pytestmark = pytest.mark.gpu
CURVES = ("BN254", "BLS12_381")


@pytest.mark.parametrize("case", H.golden_proofs(), ids=H.case_id)
def test_golden_proofs_on_the_device_batch_verifier(gpu, case):
    """Every golden proof (k = 0, 1, 2; known-tau and real ceremony setups), alone and eight times over; one flipped
    bit in each field of the proof is rejected the way the host verifier rejects it."""
    args, c, _ = tvh._verify_args(case)
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    api.verify_batch(*args, [proof], [pub], device=True)
    api.verify_batch(*args, [proof] * 8, [pub] * 8, device=True)
    api.verify_batch(*args, [], [], device=True)
    cv = c["cv"]
    step = cv.fp_bytes
    for off in range(step - 1, len(proof), 4 * step):
        bad = bytearray(proof)
        bad[off] ^= 1
        batch = [proof] * 3 + [bytes(bad)] + [proof] * 5
        with pytest.raises(ValueError) as dev:
            api.verify_batch(*args, batch, [pub] * 9, device=True)
        with pytest.raises(ValueError) as host:
            api.verify_batch(*args, batch, [pub] * 9)
        # same index / "batch" and, up to the subgroup wording, the same reason
        assert str(dev.value).split(":")[0] == str(host.value).split(":")[0], off
    if pub:
        bad = bytearray(pub)
        bad[-1] ^= 1
        with pytest.raises(ValueError):
            api.verify_batch(*args, [proof, proof], [pub, bytes(bad)], device=True)


@pytest.mark.parametrize("curve", CURVES)
def test_device_batch_of_distinct_proofs(gpu, curve):
    cv = po.CURVES[curve]
    proofs, pubs, args = [], [], None
    for i in range(12):
        cs, values = fe.squaring_chain(curve, 6, x0=3 + i)
        tc = fe.build_trace(cs)
        L, R, O = fe.solve_lro(cs, values, tc.n)
        if args is None:
            srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, tc.n + 3)
            circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
            args = (curve, tc.n, tc.nb_public, [], api.points_to_mont_bytes(curve, circ.vk_points()),
                    api.points_to_mont_bytes(curve, [cv.g1]), api.g2_unsafe(curve, api.TEST_TAU))
        proofs.append(circ.prove(L, R, O, H.scalars_uniform(cv.r, 9, 100 + i)))
        pubs.append(b"".join((v % cv.r).to_bytes(32, "big") for v in L[: tc.nb_public]))
    circ.free()
    api.verify_batch(*args, proofs, pubs, device=True)
    api.verify_batch(*args, proofs, pubs)
    pb = 2 * cv.fp_bytes
    other = po.g1_raw_bytes(cv, po.g1_mul(cv, cv.g1, 99))
    for victim in (0, 7, 11):
        bad = list(proofs)
        bad[victim] = other + proofs[victim][pb:]
        with pytest.raises(ValueError, match="batch: pairing"):
            api.verify_batch(*args, bad, pubs, device=True)
    with pytest.raises(ValueError, match="batch"):
        api.verify_batch(*args, [proofs[0], proofs[2], proofs[1]], pubs[:3], device=True)
    bad = list(proofs)
    bad[9] = b"\xff" * cv.fp_bytes + proofs[9][cv.fp_bytes:]
    with pytest.raises(ValueError, match="error verifying proof 9: "):
        api.verify_batch(*args, bad, pubs, device=True)
    if pubs[0] != pubs[1]:
        with pytest.raises(ValueError, match="batch"):
            api.verify_batch(*args, proofs, [pubs[1], pubs[0]] + pubs[2:], device=True)


def test_device_batch_rejects_points_outside_the_subgroup(gpu):
    """BLS12-381's G1 has a cofactor: a point on the curve but outside the r-torsion subgroup must be refused (gnark's
    decoders test it); the device batch runs that test as [r] P on the GPU."""
    curve = "BLS12_381"
    cv = po.CURVES[curve]
    case = next(c for c in H.golden_proofs() if c["curve"] == curve and c["name"] == "basic")
    args, _, _ = tvh._verify_args(case)
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    x = 5
    while True:                                         # a curve point of full order h * r (overwhelmingly likely)
        y2 = (x * x * x + cv.b) % cv.p
        y = pow(y2, (cv.p + 1) // 4, cv.p)
        # [r] P as (r - 1) P + P: the oracle's scalar multiplication reduces its scalar mod r
        if y * y % cv.p == y2 and po.g1_add(cv, po.g1_mul(cv, (x, y), cv.r - 1), (x, y)) is not None:
            break
        x += 1
    rogue = x.to_bytes(cv.fp_bytes, "big") + y.to_bytes(cv.fp_bytes, "big")
    bad = rogue + proof[2 * cv.fp_bytes:]
    with pytest.raises(ValueError, match="error verifying proof 1: .*subgroup"):
        api.verify_batch(*args, [proof, bad, proof], [pub] * 3, device=True)
    with pytest.raises(ValueError, match="error verifying proof 1: "):
        api.verify_batch(*args, [proof, bad, proof], [pub] * 3)


def test_device_batch_argument_errors(gpu):
    lib = _lib.load()
    assert lib.b2p_verify_batch_dev(0, 8, 1, 0, None, None, None, None, None, 0, None, 0, 0, None) == _lib.ERR_ARG
    assert lib.b2p_verify_batch_dev(9, 8, 1, 0, None, b"x", b"x", b"x", None, 0, None, 0, 0, None) == _lib.ERR_ARG
