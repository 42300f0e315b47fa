"""Times b2p_ntt-style transforms through the prover's own domain code: the per-proof NTT phase (phases_ms.ntt of a
profiled proof) with and without the compact low-pass twiddle table.  One JSON line per setting.
    B2P_NTT_COMPACT_TW=0 python tools/ntt_time.py [log2]   /   python tools/ntt_time.py [log2]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from algoplonk_b200 import api, frontend as fe       # noqa: E402

if __name__ == "__main__":
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    curve = sys.argv[2] if len(sys.argv) > 2 else "BN254"
    cs, values = fe.squaring_chain(curve, lg)
    cc = api.Compile(cs, curve, api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381)
    L, R, O = (api.fr_to_mont_bytes(curve, c) for c in fe.solve_lro(cs, values, cc.trace.n))
    bl = api.fr_to_mont_bytes(curve, list(range(1, 10)))
    for _ in range(3):
        first = cc.prove_raw(L, R, O, bl).raw
    cc.set_profiling(True)
    best = None
    for _ in range(8):
        assert cc.prove_raw(L, R, O, bl).raw == first
        st = cc.stats()
        if best is None or st["ntt_ms"] < best["ntt_ms"]:
            best = st
    print(json.dumps({"curve": curve, "log2": lg, "compact_low_twiddles": os.environ.get("B2P_NTT_COMPACT_TW", "1") != "0",
                      "ntt_ms_per_proof": round(best["ntt_ms"], 3), "msm_ms": round(best["msm_ms"], 3),
                      "total_ms": round(best["total_ms"], 3), "proof_sha": __import__("hashlib").sha256(first).hexdigest()[:16]}))
