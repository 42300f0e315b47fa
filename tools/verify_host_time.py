"""Times the host verifier (b2p_verify, b2p_verify_batch, b2p_pairing_check) on the golden BasicCircuit proofs.
CPU only: this is host code, so the build container's cores are a legitimate place to measure it (one JSON line per
curve; the CPU model and core count are in the line).  Median of 5 repetitions.

    python tools/verify_host_time.py > profiles/verify_host_r1.jsonl
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
import test_verify_host as T  # noqa: E402
from algoplonk_b200 import api  # noqa: E402
from oracle import plonk_oracle as po  # noqa: E402


def med(fn, reps=5, inner=10):
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for _ in range(inner):
            fn()
        out.append((time.perf_counter() - t0) / inner * 1e3)
    return statistics.median(out)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            return next(ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name"))
    except (OSError, StopIteration):
        return "unknown"


for case in H.golden_proofs():
    if case["name"] != "basic":
        continue
    args, _, _ = T._verify_args(case)
    curve = case["curve"]
    cv = po.CURVES[curve]
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    api.verify(*args, proof, pub)                       # builds the per-G2 line cache
    g1s = api.points_to_mont_bytes(curve, [cv.g1, po.g1_neg(cv, cv.g1)])
    g2 = args[6]
    same_g2 = g2[: len(g2) // 2] * 2
    line = {"curve": curve, "where": "build container host, 1 thread", "cpu": cpu_model(), "cores": os.cpu_count(),
            "verify_ms": med(lambda: api.verify(*args, proof, pub)),
            "pairing_check_2_pairs_ms": med(lambda: api.pairing_check(curve, g1s, same_g2)),
            "verify_batch_of_32_ms_per_proof": med(lambda: api.verify_batch(*args, [proof] * 32, [pub] * 32), inner=1) / 32}
    print(json.dumps(line), flush=True)
