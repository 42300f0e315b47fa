"""One proof over several GPUs: the commit hook + the three collectives of a sharded commitment
(algoplonk_b200/sharded_prover.py, b2p_srs_set_commit_hook).

CPU (gloo): the committer's protocol -- header, scalar broadcast, per-rank slices of ragged vectors, all_gather of one
point per rank, host add, STOP -- with the local sums supplied by the big-int oracle (no GPU compute on this path).
The GPU half (a whole proof through the hook, world 1, byte-identical to the plain prover) is in
tests/test_gpu_dense_and_cyclic.py.
"""
import json
import os
import subprocess
import sys

import pytest

from algoplonk_b200 import sharded_prover as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slices_partition_every_vector_length():
    for total in (11, 64, 67):
        for world in (1, 2, 3, 8):
            for n in (0, 1, total - 3, total - 1, total):
                parts = [sp.slice_for(total, r, world, n) for r in range(world)]
                assert sum(c for _, c in parts) == n
                pos = 0
                for off, cnt in parts:
                    assert cnt >= 0 and (cnt == 0 or off == pos)
                    pos += cnt


WORKER = r"""
import json, os, random, sys
sys.path.insert(0, {root!r})
import torch
import torch.distributed as dist
from algoplonk_b200 import api, sharded, sharded_prover as sp
from oracle import plonk_oracle as po
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
curve = {curve!r}
cv = po.CURVES[curve]
total = 35                                   # SRS points: n + 3 with n = 32
rng = random.Random(23)                      # same SRS on every rank
tau = rng.randrange(cv.r)
srs = [po.g1_mul(cv, cv.g1, pow(tau, j, cv.r)) for j in range(total)]
calls = []

def local_msm(scalars, offset, count):       # the oracle stands in for b2p_msm_g1_dev on this rank's shard
    first, cnt = sharded.shard_range(total, rank, world)
    assert offset >= first or count == 0
    assert offset + count <= first + cnt or count == 0
    vals = api.fr_from_mont_bytes(curve, bytes(scalars[32 * offset:32 * (offset + count)].numpy().tobytes()))
    acc = None
    for j, s in enumerate(vals):
        acc = po.g1_add(cv, acc, po.g1_mul(cv, srs[offset + j], s))
    calls.append(count)
    return api.points_to_mont_bytes(curve, [acc])

c = sp.ShardedCommitter(curve, total, group=None, device="cpu", local_msm=local_msm)
if rank == 0:
    rng2 = random.Random(5)
    results = []
    for n in (34, 35, 32, 1, 0, 20):         # the prover's lengths n+2, n+3, n; and vectors that leave ranks idle
        scalars = [rng2.randrange(cv.r) for _ in range(n)]
        t = torch.frombuffer(bytearray(api.fr_to_mont_bytes(curve, scalars) or b"\0"), dtype=torch.uint8)[: 32 * n]
        got = api.points_from_mont_bytes(curve, c.commit(t, n))[0]
        want = None
        for P, s in zip(srs, scalars):
            want = po.g1_add(cv, want, po.g1_mul(cv, P, s))
        assert got == want, (n, got, want)
        results.append(n)
    try:
        c.commit(torch.zeros(32 * 36, dtype=torch.uint8), 36)
        raise SystemExit("oversized vector accepted")
    except ValueError:
        pass
    c.stop()
    served = c.commits
else:
    served = c.serve()
assert served == 6, served
dist.barrier()
if rank == 0:
    print(json.dumps({{"ok": True, "world": world, "commits": served, "rank0_counts": calls}}))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("curve,port,world", [("BN254", 29571, 2), ("BLS12_381", 29572, 2), ("BN254", 29573, 3)])
def test_sharded_commit_protocol_gloo(tmp_path, curve, port, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, curve=curve))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and line["world"] == world and line["commits"] == 6
    first = {2: [18, 18, 18, 1, 0, 18], 3: [12, 12, 12, 1, 0, 12]}[world]     # rank 0 owns points [0, ceil(35/world))
    assert line["rank0_counts"] == first


def test_world_one_needs_no_process_group_and_checks_its_arguments():
    import torch
    from algoplonk_b200 import api
    c = sp.ShardedCommitter("BN254", 8, device="cpu", local_msm=lambda s, off, cnt: bytes([cnt]) + bytes(63))
    assert c.commit(torch.zeros(32 * 5, dtype=torch.uint8), 5)[0] == 5
    c.stop()
    with pytest.raises(ValueError):
        c.commit(torch.zeros(32 * 9, dtype=torch.uint8), 9)
    with pytest.raises(ValueError):
        sp.ShardedCommitter("BN254", 8)
    with pytest.raises(RuntimeError):
        c.serve()
    assert api.FP_BYTES["BN254"] == 32
