"""Times b2p_solver_solve_dev on the two shapes that matter: a wide, shallow circuit (lanes of
MiMC-style rounds) and a dependency chain (the benchmark's squaring chain), host path against device path, and
checks both against the witness the front end computed.  Prints one JSON line per case.
    python tools/solver_bench.py [log2_rows ...]      (default 16 18 20)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from algoplonk_b200 import _lib, api, frontend as fe      # noqa: E402


def case(name, cs, values, reps=5):
    tc = fe.build_trace(cs)
    t0 = time.perf_counter()
    s = api.Solver(cs, tc)
    create_s = time.perf_counter() - t0
    inputs = [values[v] for v in cs.input_vars]
    want = fe.solve_lro(cs, values, tc.n)
    out = {"case": name, "curve": cs.curve, "rows": tc.n, "create_s": round(create_s, 3)}
    for where, label in ((_lib.SOLVE_HOST, "host"), (_lib.SOLVE_DEVICE, "device")):
        assert s.solve(inputs, where) == want, (name, label)
        best = 1e30
        for _ in range(reps):
            s.solve_dev(inputs, where)
            best = min(best, s.info()["last_us"])
        out[label + "_ms"] = round(best / 1000.0, 3)
    s.solve_dev(inputs, _lib.SOLVE_AUTO)
    info = s.info()
    out.update({k: info[k] for k in ("levels", "widest_level", "solved_rows", "launches", "est_host_us", "est_device_us")})
    out["auto_picks"] = "device" if info["last_where"] == _lib.SOLVE_DEVICE else "host"
    out["note"] = "host_ms includes the upload of L, R, O (3 x 32 x rows bytes) that b2p_prove_dev needs; device_ms leaves them in HBM"
    s.free()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    logs = [int(a) for a in sys.argv[1:]] or [16, 18, 20]
    for lg in logs:
        rounds = 64 if lg >= 18 else 16
        lanes = ((1 << lg) - 2) // (4 * rounds)
        case(f"wide_mimc_{lanes}x{rounds}", *fe.wide_mimc_circuit("BN254", lanes, rounds))
    case("squaring_chain_2p16", *fe.squaring_chain("BN254", 16))
    case("merkle_depth16", *(lambda B: (B.build(), B.values))(fe.merkle_circuit("BN254")[0]))
