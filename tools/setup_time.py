"""Setup-side timings on one GPU: what ap.Compile pays gnark for on the CPU, measured here.
  ToLagrangeG1   kzg.ToLagrangeG1 over n SRS points (setup/setup.go:124,138)            b2p_srs_to_lagrange
  SRS load       loadTrustedSetupBytes + srs.Pk.ReadFrom of a compressed pk.bin (setup.go:165-228)   b2p_srs_load_compressed
  circuit load   trace upload, 8+k transforms to canonical / coset form, plonk.Setup's commitments (setup.go:149)
One JSON line.   python tools/setup_time.py BN254 20
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from algoplonk_b200 import _lib, api, frontend as fe  # noqa: E402


def main():
    curve, lg = (sys.argv[1] if len(sys.argv) > 1 else "BN254"), int(sys.argv[2]) if len(sys.argv) > 2 else 20
    _lib.init(0)
    n = 1 << lg
    out = {"curve": curve, "log2": lg}
    t0 = time.perf_counter()
    srs = api.SRS.unsafe(curve, n + 3)
    out["srs_generate_unsafe_s"] = time.perf_counter() - t0
    srs.to_lagrange_raw(64)                                        # module load, warm-up
    t0 = time.perf_counter()
    srs.to_lagrange_raw(n)
    out["to_lagrange_g1_s"] = time.perf_counter() - t0
    cs, values = fe.squaring_chain(curve, lg)
    setup = api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381
    t0 = time.perf_counter()
    tc = fe.build_trace(cs)
    out["host_trace_build_s (python front end, not gnark)"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cc = api.Compile(cs, curve, setup, srs=srs)
    out["compile_total_s (python trace build + circuit_load + VK commitments)"] = time.perf_counter() - t0
    if curve == "BN254" and lg >= 14:
        golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                              "ppot_bn254_first_131075.bin")
        with open(golden, "rb") as f:
            pk = f.read()
        t0 = time.perf_counter()
        s2 = api.SRS.from_pk_bin("BN254", pk, 131075)
        out["srs_load_compressed_131075_points_s"] = time.perf_counter() - t0
        s2.free()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
