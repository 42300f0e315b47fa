"""Stress: F lanes x K proofs concurrently, every raw proof must equal the lone reference proof."""
import ctypes as C, sys, threading, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from algoplonk_b200 import _lib, api
_lib.init(0); lib = _lib.load()
curve, log2, F, K = os.environ.get("CURVE", "BN254"), int(os.environ.get("LOG2", 20)), int(os.environ.get("F", 3)), int(os.environ.get("K", 6))
cs, tc, L, R, O = bench.build_workload(curve, log2)
setup = api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381
ccs = [api.Compile(cs, curve, setup) for _ in range(F)]
dev = torch.device("cuda", 0)
pin = lambda d: torch.frombuffer(bytearray(d), dtype=torch.uint8).pin_memory()
hL, hR, hO = (pin(api.fr_to_mont_bytes(curve, c)) for c in (L, R, O))
dL, dR, dO = (t.to(dev) for t in (hL, hR, hO))
bl = C.create_string_buffer(api.fr_to_mont_bytes(curve, list(range(1, 10))))
size = lib.b2p_proof_raw_size(api.CURVE_ID[curve], 0)
def prove(i, host):
    out = C.create_string_buffer(size)
    if host:
        _lib.check(lib.b2p_prove(ccs[i].handle, hL.data_ptr(), hR.data_ptr(), hO.data_ptr(), None, None, bl, out))
    else:
        _lib.check(lib.b2p_prove_dev(ccs[i].handle, dL.data_ptr(), dR.data_ptr(), dO.data_ptr(), None, None, bl, out))
    return bytes(out.raw)
ref = prove(0, False)
for mode in ("dev", "host"):
    bad = []
    def work(i):
        for k in range(K):
            got = prove(i, mode == "host")
            if got != ref:
                # raw proof: 9 points of 64 B (LRO Z H0 H1 H2 Wz Wzw), then 32 B scalars
                pb = 2 * api.FP_BYTES[curve]
                pts = [j for j in range(9) if got[pb * j:pb * j + pb] != ref[pb * j:pb * j + pb]]
                frs = [j for j in range((len(ref) - 9 * pb) // 32) if got[9 * pb + 32 * j:9 * pb + 32 * j + 32] != ref[9 * pb + 32 * j:9 * pb + 32 * j + 32]]
                bad.append((i, k, "pts", pts, "frs", frs))
    ths = [threading.Thread(target=work, args=(i,)) for i in range(F)]
    [t.start() for t in ths]; [t.join() for t in ths]
    print(mode, "lanes", F, "proofs", F * K, "mismatches", sorted(bad))
# sequential control
print("sequential mismatches", [(i, k) for i in range(F) for k in range(2) if prove(i, True) != ref])
