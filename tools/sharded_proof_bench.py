"""One proof over the GPUs torchrun gives it: latency of b2p_prove with its 9 MSMs sharded over the point set
(algoplonk_b200/sharded_prover.py) against the same proof on rank 0 alone.  One JSON line per case.  NOT RUN ON A GPU
YET (written after round 1's GPU budget was spent); the first thing to measure in the next round:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 \
        tools/sharded_proof_bench.py BN254:20 BLS12_381:20
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from algoplonk_b200 import _lib, api, sharded_prover as sp  # noqa: E402


def main():
    cases = sys.argv[1:] or ["BN254:20"]
    steps = int(os.environ.get("B2P_STEPS", "5"))
    rank, local_rank, world = bench.dist_env()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    _lib.init(local_rank)
    for case in cases:
        curve, log2 = case.split(":")
        cs, tc, L, R, O = bench.build_workload(curve, int(log2))
        setup = api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381
        blinding = list(range(1, 10))
        prover = sp.ShardedProver(cs, curve, setup)
        if rank != 0:
            prover.serve()
            prover.close()
            continue
        plain = None
        try:
            plain = api.Compile(cs, curve, setup)
            cols = [api.fr_to_mont_bytes(curve, c) for c in (L, R, O)]
            bl = api.fr_to_mont_bytes(curve, blinding)

            def timed(cc):
                cc.prove_raw(*cols, bl)                       # warm-up
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(steps):
                    proof = cc.prove_raw(*cols, bl)           # blocking: returns with the proof on the host
                return (time.perf_counter() - t0) / steps * 1e3, proof.raw
            ms_plain, want = timed(plain)
            ms_sharded, got = timed(prover.cc)
            line = {"curve": curve, "log2_constraints": int(log2), "n_gpus": world, "steps": steps,
                    "ms_per_proof_one_gpu": ms_plain, "ms_per_proof_msm_sharded": ms_sharded,
                    "speedup": ms_plain / ms_sharded, "byte_identical": got == want,
                    "commits_per_proof": prover.committer.commits // (steps + 1),
                    "timing": "host wall clock around blocking b2p_prove calls on rank 0 (a proof ends with its D2H)"}
        except Exception as e:  # noqa: BLE001 -- reported; the other ranks must still be released
            hook_err = getattr(getattr(prover, "_hook", None), "error", None)
            line = {"curve": curve, "log2_constraints": int(log2), "n_gpus": world,
                    "error": f"{type(e).__name__}: {e}"[:300], "hook_error": repr(hook_err)[:300] if hook_err else None}
        finally:
            print(json.dumps(line), flush=True)
            if plain is not None:
                plain.free()
            prover.close()                                    # sends STOP: ranks > 0 leave serve()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
