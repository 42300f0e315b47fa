#!/usr/bin/env python3
"""Generates tests/golden/*.json.  Run in the dev container (needs /root/reference for
the embedded SRS files); the outputs are committed so the tests never read
/root/reference at run time.

  srs_kat.json   compressed G1 points copied out of the reference's embedded SRS blobs
                 (/root/reference/setup/*/pk.bin, format: setup/setup.go:196-228): the
                 first points of each file (+ index 32767, the one
                 setup/trusted_setup_test.go:132,256 pins), and the slices used as real-SRS
                 MSM bases by the parity tests (PPoT-BN254: 259 points, Dusk: 67 points),
                 and each setup's vk.bin (160 / 240 bytes: the two G2 points of the pairing check).
  proofs.json    proofs produced by the big-integer oracle (oracle/plonk_oracle.py) on the
                 reference's own circuits (examples/basic, bsb22_test.go) and on small
                 squaring chains over the real SRS slices, with fixed blinding scalars.
                 They freeze the oracle's output: C++ oracle and CUDA path must reproduce
                 them byte for byte.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import plonk_oracle as po  # noqa: E402
from algoplonk_b200 import frontend as fe  # noqa: E402
import helpers as H  # noqa: E402

REF = "/root/reference/setup"
FILES = {
    "PerpetualPowersOfTauBN254": ("BN254", 259),
    "DuskBLS12_381": ("BLS12_381", 67),
    "EethereumKzgCeremonyBLS12_381": ("BLS12_381", 8),
}
OUT = os.path.join(ROOT, "tests", "golden")


def gen_srs():
    out = {}
    for name, (curve, count) in FILES.items():
        cv = po.CURVES[curve]
        with open(os.path.join(REF, name, "pk.bin"), "rb") as f:
            header = f.read(4)
            first = f.read(count * cv.fp_bytes)
            f.seek(4 + 32767 * cv.fp_bytes)
            p32767 = f.read(cv.fp_bytes)
        with open(os.path.join(REF, name, "vk.bin"), "rb") as f:
            vk_bin = f.read()       # 2 compressed G2 + 1 compressed G1 (setup/setup.go:216-225)
        out[name] = {
            "curve": curve,
            "declared_count": int.from_bytes(header, "big"),
            "first": first.hex(),
            "count": count,
            "index_32767": p32767.hex(),
            "vk_bin": vk_bin.hex(),
        }
    return out


def squaring_inputs(curve, log2_rows):
    cs, values = fe.squaring_chain(curve, log2_rows, x0=3)
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    return tc, L, R, O


def gen_proofs(srs_kat):
    cases = []
    blinding = list(range(1, 10))
    for curve in ("BN254", "BLS12_381"):
        cv = po.CURVES[curve]
        # examples/basic (examples/basic/logicsigVerifier/main.go:30-52), TestOnly-style known-tau SRS
        for k in (0, 1, 2):
            if k == 0:
                B = fe.basic_circuit(curve)
                cs, values, pi2s, coms = B.build(), B.values, [], []
                name = "basic"
            else:
                n_dry = fe.bsb22_circuit(curve, k, lambda a, b, c: 1).build().domain_size
                srs_o = po.srs_from_tau(cv, H.TAU, n_dry + 3)
                trd = type("T", (), {"curve": cv, "n": n_dry})
                cs, values, pi2s, coms = H.build_bsb22(curve, k, lambda col: po.bsb22_commit(trd, srs_o, col))
                name = f"bsb22_k{k}"
            tc = fe.build_trace(cs)
            L, R, O = fe.solve_lro(cs, values, tc.n)
            srs = po.srs_from_tau(cv, H.TAU, tc.n + 3)
            tr = H.oracle_trace(tc)
            vk = po.setup(tr, srs, tau=H.TAU)
            pf = po.prove(tr, vk, srs, L, R, O, blinding, pi2s, coms)
            blob = po.marshal_proof(cv, pf)
            pub = po.marshal_public_inputs(L[: tc.nb_public])
            assert po.verify_proof(vk, blob, pub)
            cases.append({
                "name": name, "curve": curve, "k": k, "srs": "tau", "n": tc.n, "blinding": blinding,
                "pi2": [[hex(v) for v in col] for col in pi2s],
                "bsb22": [po.g1_raw_bytes(cv, P).hex() for P in coms],
                "vk": po.vk_transcript_bytes(vk).hex(), "proof": blob.hex(), "public_inputs": pub.hex(),
            })
    # squaring chains over slices of the real SRS files
    for fname, log2_rows in (("PerpetualPowersOfTauBN254", 8), ("DuskBLS12_381", 6)):
        ent = srs_kat[fname]
        cv = po.CURVES[ent["curve"]]
        raw = bytes.fromhex(ent["first"])
        tc, L, R, O = squaring_inputs(ent["curve"], log2_rows)
        srs = [po.g1_decompress(cv, raw[i * cv.fp_bytes:(i + 1) * cv.fp_bytes]) for i in range(tc.n + 3)]
        tr = H.oracle_trace(tc)
        vk = po.setup(tr, srs)
        pf = po.prove(tr, vk, srs, L, R, O, blinding, [], [])
        cases.append({
            "name": f"squaring_2p{log2_rows}", "curve": ent["curve"], "k": 0, "srs": fname, "n": tc.n,
            "blinding": blinding, "pi2": [], "bsb22": [], "vk": po.vk_transcript_bytes(vk).hex(),
            "proof": po.marshal_proof(cv, pf).hex(),
            "public_inputs": po.marshal_public_inputs(L[: tc.nb_public]).hex(),
        })
    # BASELINE config 4's substitute: bsb22Circuit (bsb22_test.go:18-39) with ONE commitment on the
    # real Dusk BLS12-381 setup, and its BN254 / PPoT counterpart; the generated verifier's pairing check is
    # evaluated with the setups' own G2 points (oracle/pairing.py)
    from oracle import pairing
    for fname in ("PerpetualPowersOfTauBN254", "DuskBLS12_381"):
        ent = srs_kat[fname]
        curve = ent["curve"]
        cv = po.CURVES[curve]
        raw = bytes.fromhex(ent["first"])
        pts = [po.g1_decompress(cv, raw[i * cv.fp_bytes:(i + 1) * cv.fp_bytes]) for i in range(ent["count"])]
        n_dry = fe.bsb22_circuit(curve, 1, lambda a, b, c: 1).build().domain_size
        trd = type("T", (), {"curve": cv, "n": n_dry})
        cs, values, pi2s, coms = H.build_bsb22(curve, 1, lambda col: po.bsb22_commit(trd, pts[: n_dry + 3], col))
        tc = fe.build_trace(cs)
        L, R, O = fe.solve_lro(cs, values, tc.n)
        srs = pts[: tc.n + 3]
        g2, _ = pairing.parse_vk_bin(cv, bytes.fromhex(ent["vk_bin"]))
        tr = H.oracle_trace(tc)
        vk = po.setup(tr, srs, g2=g2)
        pf = po.prove(tr, vk, srs, L, R, O, blinding, pi2s, coms)
        blob = po.marshal_proof(cv, pf)
        pub = po.marshal_public_inputs(L[: tc.nb_public])
        assert po.verify_proof(vk, blob, pub)
        cases.append({
            "name": "bsb22_k1", "curve": curve, "k": 1, "srs": fname, "n": tc.n, "blinding": blinding,
            "pi2": [[hex(v) for v in col] for col in pi2s],
            "bsb22": [po.g1_raw_bytes(cv, P).hex() for P in coms],
            "vk": po.vk_transcript_bytes(vk).hex(), "proof": blob.hex(), "public_inputs": pub.hex(),
        })
    return cases


def main():
    os.makedirs(OUT, exist_ok=True)
    srs = gen_srs()
    with open(os.path.join(OUT, "srs_kat.json"), "w") as f:
        json.dump(srs, f, indent=1)
    proofs = gen_proofs(srs)
    with open(os.path.join(OUT, "proofs.json"), "w") as f:
        json.dump(proofs, f, indent=1)
    print("wrote", len(srs), "SRS entries and", len(proofs), "proofs")


if __name__ == "__main__":
    main()
