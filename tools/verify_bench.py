"""Batch verification: host threads (b2p_verify_batch) against the device path (b2p_verify_batch_dev) on batches of one
circuit's proofs.  python tools/verify_bench.py [curve] [batch sizes ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H                                   # noqa: E402
import test_verify_host as tvh                        # noqa: E402
from algoplonk_b200 import api                        # noqa: E402

if __name__ == "__main__":
    curve = sys.argv[1] if len(sys.argv) > 1 else "BN254"
    sizes = [int(a) for a in sys.argv[2:]] or [16, 128, 1024]
    case = next(c for c in H.golden_proofs() if c["curve"] == curve and c["name"] == "basic")
    args, _, _ = tvh._verify_args(case)
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    api.verify_batch(*args, [proof] * 4, [pub] * 4, device=True)              # context, module load
    for b in sizes:
        out = {"curve": curve, "batch": b, "host_threads": os.environ.get("B2P_VERIFY_THREADS", "default min(cores, 8)"),
               "cores": len(os.sched_getaffinity(0))}
        for label, dev in (("host", False), ("device", True)):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                api.verify_batch(*args, [proof] * b, [pub] * b, device=dev)
                best = min(best, time.perf_counter() - t0)
            out[label + "_us_per_proof"] = round(best / b * 1e6, 2)
        print(json.dumps(out), flush=True)
