"""Small proofs on both curves (k = 0 and a BSB22 commitment) plus an MSM and an NTT: the workload run under
compute-sanitizer (memcheck / racecheck / initcheck) -- no torch, a few seconds natively."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from algoplonk_b200 import _lib, api, frontend as fe
import helpers as H
_lib.init(0)
for curve, setup in (("BN254", api.SetupName.TestOnlyBN254), ("BLS12_381", api.SetupName.TestOnlyBLS12381)):
    cs, values = fe.squaring_chain(curve, int(os.environ.get("LOG2", 9)))
    cc = api.Compile(cs, curve, setup)
    L, R, O = fe.solve_lro(cs, values, cc.trace.n)
    p1 = api.MarshalProof(cc.Prove(L, R, O, list(range(1, 10))))
    assert p1 == api.MarshalProof(cc.Prove(L, R, O, list(range(1, 10))))
    cc.free()
    case = next(c for c in H.golden_proofs() if c["curve"] == curve and c["name"] == "bsb22_k1" and c["srs"] == "tau")
    c = H.build_case(case)
    cc = api.Compile(c["cs"], curve, setup)
    vp = cc.Verify(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"])
    assert api.MarshalProof(vp.Proof).hex() == case["proof"]
    cc.free()
    srs = api.SRS.unsafe(curve, 300)
    srs.msm(list(range(300)))
    srs.msm(H.scalars_witness_like(api.R_MOD[curve], 300, 1))
    srs.free()
    api.ntt(curve, list(range(1 << 9)))
    api.ntt(curve, list(range(1 << 12)), inverse=True, coset=True)
print("sanitize workload ok")
