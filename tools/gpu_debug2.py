#!/usr/bin/env python3
"""Focused debug: BLS12-381 bsb22 k=1/2 batched opening mismatch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import plonk_oracle as po
from algoplonk_b200 import api, frontend as fe, _lib
import helpers as H
_lib.init()
for case in H.golden_proofs():
    if case["srs"] != "tau": continue
    c = H.build_case(case)
    cv = c["cv"]; curve = case["curve"]
    cc = api.Compile(c["cs"], curve, api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381)
    tr = H.oracle_trace(c["tc"]); vk = po.setup(tr, c["srs"], tau=c["tau"])
    pf_o, dbg = po.prove(tr, vk, c["srs"], c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"], return_debug=True)
    for rep in range(3):
        pf = cc.Prove(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"])
        blob = api.MarshalProof(pf)
        npts = 9 * 2 * cv.fp_bytes
        frs = api.fr_from_mont_bytes(curve, pf.raw[npts:])
        pts = api.points_from_mont_bytes(curve, pf.raw[:npts])
        print(H.case_id(case), "rep", rep, "identical", blob.hex() == case["proof"], "lin_z ok", frs[0] == pf_o.claimed[0],
              "claimed ok", frs[1:-1] == pf_o.claimed[1:], "Wz ok", pts[7] == pf_o.batched_H, flush=True)
    # what would W_z be if the GPU had used k claimed values fewer / more in the fold hash?
    cc.free()
