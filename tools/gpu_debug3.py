#!/usr/bin/env python3
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import plonk_oracle as po
from algoplonk_b200 import api, frontend as fe, _lib
import helpers as H
_lib.init()
case = next(c for c in H.golden_proofs() if c["curve"] == "BLS12_381" and c["name"] == "bsb22_k1")
c = H.build_case(case); cv = c["cv"]; curve = "BLS12_381"
tr = H.oracle_trace(c["tc"]); vk = po.setup(tr, c["srs"], tau=c["tau"])
pf_o, dbg = po.prove(tr, vk, c["srs"], c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"], return_debug=True)
srs = api.SRS.unsafe(curve, c["tc"].n + 3, H.TAU)
print("params", srs.msm_params())
lin = dbg["lin"]; folded = dbg["folded"]
print("lin len", len(lin), "commit lin ok:", srs.msm(lin) == dbg["com_lin"])
q = po.poly_div_linear(cv, folded, dbg["zeta"])
got = srs.msm(q)
print("Wz msm ok:", got == pf_o.batched_H)
if got != pf_o.batched_H:
    # bisect: which single term is wrong?
    pts = c["srs"]
    for i, s in enumerate(q):
        v = [0] * len(q); v[i] = s
        g = srs.msm(v); e = po.g1_mul(cv, pts[i], s)
        print("  term", i, "ok" if g == e else "BAD", hex(s))
for name, vec in (("lc", dbg["lc"]), ("zc", dbg["zc"])):
    print(name, srs.msm(vec) == po.msm_naive(cv, c["srs"], vec))
