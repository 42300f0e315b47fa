#!/usr/bin/env python3
"""Staged GPU-vs-oracle comparison with verbose output (development aid)."""
import random
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import plonk_oracle as po
from algoplonk_b200 import api, frontend as fe, _lib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as H

random.seed(5)
_lib.init()
print("version", _lib.load().b2p_version())
for curve in ("BN254", "BLS12_381"):
    cv = po.CURVES[curve]
    # --- NTT
    for logn in (1, 3, 6, 11, 13):
        n = 1 << logn
        a = [random.randrange(cv.r) for _ in range(n)]
        w = po.domain_generator(cv, n)
        got = api.ntt(curve, a)
        exp = po.ntt(cv, a, w)
        bad = sum(1 for x, y in zip(got, exp) if x != y)
        inv = api.ntt(curve, got, inverse=True)
        cos = api.ntt(curve, a, coset=True)
        expc = po.coset_ntt(cv, a, w, cv.coset_shift)
        cinv = api.ntt(curve, cos, inverse=True, coset=True)
        print(curve, "ntt", n, "fwd bad", bad, "roundtrip", inv == a, "coset bad",
              sum(1 for x, y in zip(cos, expc) if x != y), "coset roundtrip", cinv == a, flush=True)
    # --- SRS + MSM
    for n in (4, 11, 300):
        t0 = time.time()
        srs = api.SRS.unsafe(curve, n, H.TAU)
        pts = srs.points(0, n)
        exp_pts = po.srs_from_tau(cv, H.TAU, n)
        print(curve, "srs", n, "match", pts == exp_pts, srs.msm_params(), "%.2fs" % (time.time() - t0), flush=True)
        for trial, sc in enumerate([[random.randrange(cv.r) for _ in range(n)], [0] * n, [1] * n,
                                    [cv.r - 1] * n, [random.choice([0, 1, 2, cv.r - 1, random.randrange(cv.r)]) for _ in range(n)]]):
            got = srs.msm(sc)
            exp = po.msm_naive(cv, exp_pts, sc)
            print("   msm trial", trial, "ok" if got == exp else "MISMATCH", flush=True)
        # loaded SRS (host points) path
        srs2 = api.SRS.from_points(curve, exp_pts)
        sc = [random.randrange(cv.r) for _ in range(n)]
        print("   loaded-srs msm", srs2.msm(sc) == po.msm_naive(cv, exp_pts, sc), flush=True)
        srs.free(); srs2.free()
    # --- prove
    for k in (0, 1, 2):
        if k == 0:
            B = fe.basic_circuit(curve)
            cs, values, pi2s, coms = B.build(), B.values, [], []
        else:
            n_dry = fe.bsb22_circuit(curve, k, lambda a, b, c: 1).build().domain_size
            srs_o = po.srs_from_tau(cv, H.TAU, n_dry + 3)
            trd = type("T", (), {"curve": cv, "n": n_dry})
            cs, values, pi2s, coms = H.build_bsb22(curve, k, lambda col: po.bsb22_commit(trd, srs_o, col))
        cc = api.Compile(cs, curve, api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381)
        tc = cc.trace
        L, R, O = fe.solve_lro(cs, values, tc.n)
        srs_o = po.srs_from_tau(cv, H.TAU, tc.n + 3)
        tr = H.oracle_trace(tc)
        vk_o = po.setup(tr, srs_o, tau=H.TAU)
        vk_pts = cc.vk_commitments()
        exp_vk = vk_o.S + [vk_o.Ql, vk_o.Qr, vk_o.Qm, vk_o.Qo, vk_o.Qk] + vk_o.Qcp
        print(curve, "k", k, "n", tc.n, "vk match", vk_pts == exp_vk, flush=True)
        blinding = list(range(1, 10))
        pf = cc.Prove(L, R, O, blinding, pi2s, coms)
        blob = api.MarshalProof(pf)
        pf_o, dbg = po.prove(tr, vk_o, srs_o, L, R, O, blinding, pi2s, coms, return_debug=True)
        blob_o = po.marshal_proof(cv, pf_o)
        pub = api.MarshalPublicInputs(curve, L[: tc.nb_public])
        print("   proof bytes identical:", blob == blob_o, "oracle verify(gpu proof):",
              po.verify_proof(vk_o, blob, pub), flush=True)
        if blob != blob_o:
            PB = 2 * cv.fp_bytes
            names = ["L", "R", "O", "H0", "H1", "H2"]
            off = 0
            for nm in names:
                print("     ", nm, blob[off:off + PB] == blob_o[off:off + PB]); off += PB
            for nm in ["l", "r", "o", "s1", "s2"]:
                print("     ", nm, blob[off:off + 32] == blob_o[off:off + 32]); off += 32
            print("      Z", blob[off:off + PB] == blob_o[off:off + PB]); off += PB
            print("      z(wz)", blob[off:off + 32] == blob_o[off:off + 32]); off += 32
            print("      Wz", blob[off:off + PB] == blob_o[off:off + PB]); off += PB
            print("      Wzw", blob[off:off + PB] == blob_o[off:off + PB]); off += PB
        print("   stats", cc.stats(), flush=True)
print("launches", _lib.load().b2p_launch_count())
