"""Ad-hoc probe: throughput of F proofs in flight, resident vs host inputs, in different orders."""
import ctypes as C, sys, threading, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from algoplonk_b200 import _lib, api
_lib.init(0); lib = _lib.load()
curve, log2, F = "BN254", int(os.environ.get("LOG2", 20)), 2
cs, tc, L, R, O = bench.build_workload(curve, log2)
ccs = [api.Compile(cs, curve, api.SetupName.TestOnlyBN254) for _ in range(F)]
dev = torch.device("cuda", 0)
pin = lambda d: torch.frombuffer(bytearray(d), dtype=torch.uint8).pin_memory()
hL, hR, hO = (pin(api.fr_to_mont_bytes(curve, c)) for c in (L, R, O))
dLs = [[t.to(dev).clone() for t in (hL, hR, hO)] for _ in range(F)]
bl = C.create_string_buffer(api.fr_to_mont_bytes(curve, list(range(1, 10))))
outs = [C.create_string_buffer(lib.b2p_proof_raw_size(0, 0)) for _ in range(F)]
def pdev(i):
    a, b, c = dLs[i]
    _lib.check(lib.b2p_prove_dev(ccs[i].handle, a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, bl, outs[i]))
def phost(i):
    _lib.check(lib.b2p_prove(ccs[i].handle, hL.data_ptr(), hR.data_ptr(), hO.data_ptr(), None, None, bl, outs[i]))
def run(fn, per_lane, lanes, delay=0.0):
    torch.cuda.synchronize()
    walls = [0.0] * lanes
    def work(i):
        if i and delay: time.sleep(delay)
        t0 = time.perf_counter()
        for _ in range(per_lane): fn(i)
        walls[i] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(lanes)]
    [t.start() for t in ths]; [t.join() for t in ths]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return f"{per_lane * lanes / dt:6.2f} proofs/s  total {dt*1e3:7.1f} ms  lanes " + " ".join(f"{w*1e3:7.1f}" for w in walls)
for i in range(F):
    for _ in range(2): pdev(i); phost(i)
print("dev  x1      ", run(pdev, 4, 1))
print("host x1      ", run(phost, 4, 1))
print("dev  x2      ", run(pdev, 3, 2))
print("host x2      ", run(phost, 3, 2))
print("dev  x2 again", run(pdev, 3, 2))
print("dev  x2 +15ms", run(pdev, 3, 2, 0.015))
print("host x2 again", run(phost, 3, 2))
# --- variants that mimic bench.py -------------------------------------------------
sh = [t.to(dev) for t in (hL, hR, hO)]
def pdev_shared(i):
    _lib.check(lib.b2p_prove_dev(ccs[i].handle, sh[0].data_ptr(), sh[1].data_ptr(), sh[2].data_ptr(), None, None, bl, outs[i]))
print("dev shared x2", run(pdev_shared, 3, 2))
ccs[0].set_profiling(True)
print("dev x1 prof  ", run(pdev, 4, 1))
ccs[0].set_profiling(False)
print("dev x2 after ", run(pdev, 3, 2))
print("dev shared x2", run(pdev_shared, 3, 2))
streams = [torch.cuda.ExternalStream(lib.b2p_circuit_stream(cc.handle), device=dev) for cc in ccs]
def run_ev(fn, per_lane, lanes):
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
    def work(i):
        for _ in range(per_lane): fn(i)
        ends[i].record(streams[i])
    for i in range(lanes): starts[i].record(streams[i])
    ths = [threading.Thread(target=work, args=(i,)) for i in range(lanes)]
    [t.start() for t in ths]; [t.join() for t in ths]
    for e in ends: e.synchronize()
    ms = max(s.elapsed_time(e) for s in starts for e in ends)
    return f"{per_lane*lanes/ms*1e3:6.2f} proofs/s (events) {ms:7.1f} ms"
print("dev x2 events", run_ev(pdev, 3, 2))
print("dev sh events", run_ev(pdev_shared, 3, 2))
print("host x2 event", run_ev(phost, 3, 2))
