"""The domain-sharded NTT leg of bench.py on its own (no proving keys, no circuit): both exchange modes over the
GPUs torchrun gives it.  One JSON line per (curve, log2 of the circuit size); the transform has 4 * 2^log2 points.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        tools/ntt_shard_bench.py BN254:20 BLS12_381:21
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from algoplonk_b200 import _lib  # noqa: E402


def main():
    cases = sys.argv[1:] or ["BN254:20"]
    rank, local_rank, world = bench.dist_env()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    _lib.init(local_rank)
    for case in cases:
        curve, log2 = case.split(":")
        args = argparse.Namespace(curve=curve, log2=int(log2))
        line = bench.measure_sharded_ntt(args, rank, world, device)
        if rank == 0:
            print(json.dumps({"curve": curve, "log2_constraints": int(log2), "n_gpus": world, **line}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
