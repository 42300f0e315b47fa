"""One-shot G2 MSM timing (b2p_msm_g2: table build + Pippenger over Fp2, host buffers in, one G2Affine out), checked
against the closed form sum s_i k_i G2.  python tools/msm_g2_bench.py [log2_points] [curve]"""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from algoplonk_b200 import api                       # noqa: E402
from oracle import pairing as opair                  # noqa: E402  (checker only)
from oracle import plonk_oracle as po                # noqa: E402

if __name__ == "__main__":
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    curve = sys.argv[2] if len(sys.argv) > 2 else "BN254"
    cv = po.CURVES[curve]
    n = 1 << lg
    gen = api.g2_from_mont_bytes(curve, api.g2_unsafe(curve, 1))[0]
    rng = random.Random(1)
    k = rng.randrange(cv.r)
    P = opair.g2_mul(cv, gen, k)
    ks, pts = [], []
    for _ in range(n):
        ks.append(k)
        pts.append(P)
        k = (k + 1) % cv.r
        P = opair.g2_add(cv, P, gen)
    sc = [rng.randrange(cv.r) for _ in range(n)]
    raw = api.g2_to_mont_bytes(curve, pts)
    api.msm_g2_raw(curve, raw[: 4 * 4 * cv.fp_bytes], sc[:4])          # warm-up: context, module load
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        out = api.msm_g2_raw(curve, raw, sc)
        best = min(best, time.perf_counter() - t0)
    ok = api.g2_from_mont_bytes(curve, out)[0] == opair.g2_mul(cv, gen, sum(s * q for s, q in zip(sc, ks)) % cv.r)
    print(json.dumps({"what": "b2p_msm_g2 one shot (scalar conversion in Python included)", "curve": curve, "points": n,
                      "ms": round(best * 1e3, 2), "matches_closed_form": ok}))
