"""Per-rank launch times of the domain-sharded NTT (csrc/ntt_shard.cuh) on ONE GPU: rank 0 of a world of G does
exactly the work it does on a G-GPU box; the chunks it would read from / write to its peers over NVLink live in
local HBM here, so the numbers are the compute side of the exchange kernels (the NVLink side needs `--gpus G`).
Prints one JSON line per (curve, n, world).  CUDA events on the launching stream, 3 warm-up + 20 timed launches.

    python tools/ntt_shard_time.py [--logn 21] [--worlds 1,2,4,8]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from algoplonk_b200 import sharded_ntt as sn  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=21)
    ap.add_argument("--worlds", default="1,2,4,8")
    args = ap.parse_args()
    n = 1 << args.logn
    for curve in ("BN254", "BLS12_381"):
        for world in [int(w) for w in args.worlds.split(",")]:
            st = sn.CudaSteps(curve, n, world, 0)
            ln, chunk = n // world, n // world // world
            g = torch.Generator(device="cpu").manual_seed(1)
            coeffs = torch.randint(0, 1 << 60, (ln, 4), generator=g, dtype=torch.int64).cuda()
            # stand-ins for the peers' buffers: world distinct buffers, so the access pattern is the real one
            peers = [torch.empty((ln, 4), dtype=torch.int64, device="cuda") for _ in range(world)]
            for p in peers:
                p.copy_(coeffs)
            out = torch.empty((ln, 4), dtype=torch.int64, device="cuda")
            chunks = [p[:chunk] for p in peers]
            res = {
                "curve": curve, "logn": args.logn, "world": world, "local_elements": ln,
                "forward_local_coset_ms": timed(lambda: st.forward_local(coeffs, ln, True, peers[0])),
                "forward_combine_ms": timed(lambda: st.forward_combine(chunks, out)),
                "inverse_split_ms": timed(lambda: st.inverse_split(out, chunks)),
                "inverse_local_coset_ms": timed(lambda: st.inverse_local(peers[0], True)),
            }
            res["exchange_bytes_per_rank"] = 32 * ln * (world - 1) // world
            res["combine_algorithmic_GBps"] = 2 * 32 * ln / res["forward_combine_ms"] / 1e6
            res["split_algorithmic_GBps"] = 2 * 32 * ln / res["inverse_split_ms"] / 1e6
            print(json.dumps(res), flush=True)
            st.free()


if __name__ == "__main__":
    main()
