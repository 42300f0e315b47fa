"""The kernels added in the second half of round 2 -- witness solver (wide levels, the single-block narrow path,
division rows), segment combinations of the device batch verifier (incl. BLS12-381 r-torsion segments), G2 MSM,
key snapshot round trip -- as one small workload for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from algoplonk_b200 import _lib, api, frontend as fe
import helpers as H
import test_verify_host as tvh
from oracle import pairing as opair, plonk_oracle as po
_lib.init(0)
for curve, setup in (("BN254", api.SetupName.TestOnlyBN254), ("BLS12_381", api.SetupName.TestOnlyBLS12381)):
    cv = po.CURVES[curve]
    # solver: wide levels (multi-block launches), narrow runs (one block), a chain, a division row
    for cs, values in (fe.wide_mimc_circuit(curve, 300, 3), fe.wide_mimc_circuit(curve, 20, 5), fe.squaring_chain(curve, 7)):
        s = api.Solver(cs)
        want = fe.solve_lro(cs, values, s.n)
        for where in (_lib.SOLVE_DEVICE, _lib.SOLVE_HOST):
            assert s.solve([values[v] for v in cs.input_vars], where) == want
        s.free()
    # hints: the launch plan cut at hint levels, single values down and up (device), plain calls (host)
    Bh = fe.Builder(curve)
    xh = Bh.public(0b101101)
    bits = Bh.to_binary(Bh.mul(xh, Bh.secret(3)), 10)
    Bh.to_binary(Bh.hint(fe.HINT_NBITS, [bits[1]], [Bh.values[bits[1]] & 1])[0], 1)
    csh = Bh.build()
    s = api.Solver(csh, hint_fn=api.std_hint_fn())
    for where in (_lib.SOLVE_DEVICE, _lib.SOLVE_HOST):
        assert s.solve([Bh.values[v] for v in csh.input_vars], where) == fe.solve_lro(csh, Bh.values, s.n)
    s.free()
    B = fe.Builder(curve)
    x = B.public(5)
    B.assert_is_different_from_zero(x)
    s = api.Solver(B.build())
    assert s.solve([5], _lib.SOLVE_DEVICE)[1][1] == pow(5, -1, cv.r)
    s.free()
    # inputs -> verified proof with L R O staying on the device, then the key snapshot
    cs, values = fe.squaring_chain(curve, 8)
    cc = api.Compile(cs, curve, setup)
    s = api.Solver(cs, cc.trace)
    vp = api.VerifyFromInputs(cc, s, [values[v] for v in cs.input_vars], list(range(1, 10)), _lib.SOLVE_DEVICE)
    blob, pub = api.MarshalProof(vp.Proof), api.MarshalPublicInputs(curve, vp.Witness)
    with tempfile.TemporaryDirectory() as d:
        api.SerializeCompiledCircuit(cc, os.path.join(d, "k.b2pk"))
        cc2 = api.DeserializeCompiledCircuit(os.path.join(d, "k.b2pk"), cs, cc.srs)
        assert api.MarshalProof(cc2.Prove(*fe.solve_lro(cs, values, cc.trace.n), list(range(1, 10)))) == blob
        cc2.free()
    # device batch verification: accepted batch, a tampered proof
    cc.VerifyProofs([blob] * 9, [pub] * 9, device=True)
    bad = bytearray(blob); bad[40] ^= 1
    try:
        cc.VerifyProofs([blob] * 4 + [bytes(bad)] + [blob] * 4, [pub] * 9, device=True)
        raise SystemExit("tampered proof accepted")
    except ValueError:
        pass
    s.free(); cc.free()
    # G2 MSM
    gen = api.g2_from_mont_bytes(curve, api.g2_unsafe(curve, 1))[0]
    pts = [opair.g2_mul(cv, gen, k) for k in range(1, 40)]
    sc = H.scalars_uniform(cv.r, len(pts), 3)
    got = api.g2_from_mont_bytes(curve, api.msm_g2_raw(curve, api.g2_to_mont_bytes(curve, pts), sc))[0]
    assert got == opair.g2_mul(cv, gen, sum(k * s_ for k, s_ in zip(range(1, 40), sc)) % cv.r)
print("sanitize workload (round-2 rows) ok")
