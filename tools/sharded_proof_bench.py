"""One proof over the GPUs torchrun gives it: latency of b2p_prove with its 9 commitments sharded over the point set
(native path: algoplonk_b200/shard_group.py, csrc/shard_group.cuh) against the same proof on rank 0 alone.  One JSON
line per case; the proof bytes must equal the single-GPU proof.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 \
        tools/sharded_proof_bench.py BN254:20 BN254:20:c=18 BLS12_381:21

case = curve:log2_rows[:c=<window bits of the per-rank SRS blocks>][:ntt]   (B2P_STEPS proofs timed, default 5)
":ntt" also spreads the proof's five size-4n transforms over the ranks (world a power of two).
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from algoplonk_b200 import _lib, api, shard_group as sg  # noqa: E402


def main():
    cases = sys.argv[1:] or ["BN254:20"]
    steps = int(os.environ.get("B2P_STEPS", "5"))
    rank, local_rank, world = bench.dist_env()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    _lib.init(local_rank)
    lib = _lib.load()
    for case in cases:
        parts = case.split(":")
        curve, log2 = parts[0], int(parts[1])
        opts = parts[2:]
        shard_ntt = "ntt" in opts
        shard_c = next((int(o.split("=")[1]) for o in opts if o.startswith("c=")), 0)
        cs, tc, L, R, O = bench.build_workload(curve, log2)
        setup = api.SetupName.TestOnlyBN254 if curve == "BN254" else api.SetupName.TestOnlyBLS12381
        if shard_c:
            os.environ["B2P_MSM_C"] = str(shard_c)
        grp = sg.ShardGroup(curve, tc.n + 3, ntt_rows=tc.n if shard_ntt else 0)    # collective: every rank's SRS block
        os.environ.pop("B2P_MSM_C", None)
        line = {"curve": curve, "log2_constraints": log2, "n_gpus": world, "steps": steps, "transforms_sharded": shard_ntt}
        cc = None
        if rank == 0:
            try:
                cc = api.Compile(cs, curve, setup)        # full SRS on rank 0: the setup's commitments are local
                grp.attach(cc)
            except Exception as e:  # noqa: BLE001
                line["error"] = f"setup: {type(e).__name__}: {e}"[:400]
        grp.connect()                                     # collective: IPC handles, mapping
        if rank != 0:
            grp.serve()
            grp.free()
            continue
        try:
            if "error" in line:
                raise RuntimeError(line["error"])
            lib.b2p_shard_group_attach(grp.handle, None, None)      # the single-GPU reference proofs come first
            cols = [torch.frombuffer(bytearray(api.fr_to_mont_bytes(curve, c)), dtype=torch.uint8).pin_memory()
                    for c in (L, R, O)]
            bl = C.create_string_buffer(api.fr_to_mont_bytes(curve, list(range(1, 10))))
            out = C.create_string_buffer(lib.b2p_proof_raw_size(api.CURVE_ID[curve], 0))
            stream = torch.cuda.ExternalStream(lib.b2p_circuit_stream(cc.handle), device=device)

            def timed(announce):
                def one():
                    if announce:
                        grp.announce(tc.n)
                    _lib.check(lib.b2p_prove(cc.handle, cols[0].data_ptr(), cols[1].data_ptr(), cols[2].data_ptr(),
                                             None, None, bl, out))
                for _ in range(2):
                    one()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                t0 = time.perf_counter()
                for _ in range(steps):
                    one()
                wall = (time.perf_counter() - t0) / steps * 1e3
                e1.record(stream)
                e1.synchronize()
                return e0.elapsed_time(e1) / steps, wall, bytes(out.raw)

            ms_plain, wall_plain, want = timed(False)
            _lib.check(lib.b2p_shard_group_attach(grp.handle, cc.srs.handle, cc.handle))
            cc.set_profiling(True)
            ms_sh, wall_sh, got = timed(True)
            stats = cc.stats()
            cc.set_profiling(False)
            c_bits, windows, _ = api.SRS(curve, grp.shard.handle).msm_params()
            line.update({"ms_per_proof_one_gpu": ms_plain, "ms_per_proof_sharded": ms_sh, "speedup": ms_plain / ms_sh,
                         "host_wall_ms_one_gpu": wall_plain, "host_wall_ms_sharded": wall_sh,
                         "byte_identical": got == want, "points_per_gpu": grp.shard.count, "shard_c": c_bits,
                         "shard_windows": windows,
                         "phases_ms_sharded": {"msm": stats["msm_ms"], "ntt": stats["ntt_ms"],
                                               "quotient": stats["quotient_ms"]},
                         "timing": "CUDA events on rank 0's proving stream around blocking b2p_prove calls with pinned "
                                   "host columns (H2D and D2H inside)"})
        except Exception as e:  # noqa: BLE001 -- reported; the other ranks must still be released
            line["error"] = f"{type(e).__name__}: {e}"[:400]
        finally:
            print(json.dumps(line), flush=True)
            grp.stop()                                   # the other ranks leave serve()
            grp.free()
            if cc is not None:
                srs = cc.srs
                cc.free()
                srs.free()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
