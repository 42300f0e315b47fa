"""MSM sharded over the point set (algoplonk_b200/sharded.py).

CPU (gloo, world_size 2): the partition, the one collective (all_gather of one point per rank) and the
host-side G-point add of the library, checked against the big-int oracle.  No GPU compute on this path:
the per-rank partial sums come from the oracle here, from the CUDA MSM in the `-m gpu` test below.
"""
import json
import os
import random
import subprocess
import sys

import pytest

from algoplonk_b200 import api, sharded
from oracle import plonk_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CURVES = ("BN254", "BLS12_381")


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharded.shard_range(8, 2, 2)


def test_shard_indices_both_layouts():
    for total in (0, 1, 7, 64, 1003):
        for world in (1, 2, 4, 8):
            for layout in ("blocks", "cyclic"):
                idx = [sharded.shard_indices(total, r, world, layout) for r in range(world)]
                assert sorted(i for rng_ in idx for i in rng_) == list(range(total))
                assert max(len(r) for r in idx) - min(len(r) for r in idx) <= 1
            # cyclic = the coefficient distribution of the domain-sharded NTT
            assert list(sharded.shard_indices(total, world - 1, world, "cyclic")) == list(range(world - 1, total, world))
    with pytest.raises(ValueError):
        sharded.shard_indices(8, 0, 2, "striped")
    with pytest.raises(ValueError):
        sharded.shard_indices(8, 2, 2, "cyclic")


@pytest.mark.parametrize("curve", CURVES)
def test_g1_sum_matches_oracle(curve):
    cv = po.CURVES[curve]
    rng = random.Random(5)
    pts = [po.g1_mul(cv, cv.g1, rng.randrange(1, cv.r)) for _ in range(6)]
    cases = [[], [pts[0]], pts, [pts[0], None, pts[1]], [pts[2], po.g1_neg(cv, pts[2])], [pts[3], pts[3]]]
    for case in cases:
        want = None
        for P in case:
            want = po.g1_add(cv, want, P)
        got = api.points_from_mont_bytes(curve, sharded.g1_sum(curve, api.points_to_mont_bytes(curve, case)))[0]
        assert got == want


WORKER = r"""
import json, os, random, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from algoplonk_b200 import api, sharded
from oracle import plonk_oracle as po
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
curve = {curve!r}
cv = po.CURVES[curve]
n = 37                                    # ragged: 19 + 18
rng = random.Random(11)                   # same inputs on every rank
tau = rng.randrange(cv.r)
srs = [po.g1_mul(cv, cv.g1, pow(tau, j, cv.r)) for j in range(n)]
scalars = [rng.randrange(cv.r) for _ in range(n)]
first, count = sharded.shard_range(n, rank, world)
layout = {layout!r}
mine = sharded.shard_indices(n, rank, world, layout)
# this rank's partial sum (oracle here; the CUDA MSM in the GPU test) ...
local = None
for i in mine:
    local = po.g1_add(cv, local, po.g1_mul(cv, srs[i], scalars[i]))
# ... then the product's collective + local add
gathered = sharded.all_gather_points(curve, api.points_to_mont_bytes(curve, [local]))
assert len(gathered) == world * 2 * api.FP_BYTES[curve]
total = api.points_from_mont_bytes(curve, sharded.g1_sum(curve, gathered))[0]
want = None
for P, s in zip(srs, scalars):
    want = po.g1_add(cv, want, po.g1_mul(cv, P, s))
assert total == want, (rank, total, want)
dist.barrier()
if rank == 0:
    print(json.dumps({{"ok": True, "world": world, "first_count": [first, count]}}))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("curve,port,layout", [("BN254", 29551, "blocks"), ("BLS12_381", 29552, "blocks"),
                                               ("BN254", 29553, "cyclic")])
def test_sharded_msm_collective_gloo_world2(tmp_path, curve, port, layout):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, curve=curve, layout=layout))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    assert json.loads(line) == {"ok": True, "world": 2, "first_count": [0, 19]}


@pytest.mark.gpu
@pytest.mark.parametrize("curve", CURVES)
def test_sharded_msm_on_one_gpu_equals_whole_msm(gpu, curve):
    """All shards on one device (world simulated in-process): sum of per-shard CUDA MSMs == whole CUDA MSM
    == oracle, ragged split, including a shard generated with b2p_srs_generate_unsafe_range."""
    cv = po.CURVES[curve]
    n, world = 1000, 3
    rng = random.Random(3)
    scalars = [rng.randrange(cv.r) for _ in range(n)]
    whole = api.SRS.unsafe(curve, n)
    want = whole.msm(scalars)
    parts = []
    for r in range(world):
        sh = sharded.ShardedSRS.unsafe(curve, n, r, world)
        assert (sh.first, sh.count) == sharded.shard_range(n, r, world)
        # the shard holds exactly the whole SRS's points of its range
        pts = api.SRS(curve, sh.handle).points(0, 2)
        assert pts == whole.points(sh.first, 2)
        parts.append(sh.local_msm_raw(api.fr_to_mont_bytes(curve, scalars[sh.first:sh.first + sh.count])))
        sh.free()
    got = api.points_from_mont_bytes(curve, sharded.g1_sum(curve, b"".join(parts)))[0]
    assert got == want
    # and against the oracle on the same points
    srs_pts = whole.points(0, n)
    acc = None
    for P, s in zip(srs_pts[:50], scalars[:50]):
        acc = po.g1_add(cv, acc, po.g1_mul(cv, P, s))
    assert whole.msm(scalars[:50]) == acc
    whole.free()


