"""CPU test (gloo, world_size 2) of the host plumbing of the native multi-GPU path (algoplonk_b200/shard_group.py): the
ONE control message per proof / stand-alone commitment -- (op, n, repetitions) broadcast by rank 0 -- and the serve loop
that answers it with b2p_shard_group_serve_proof / _serve_msm.  The library calls are recorded by a stand-in (the C
entry points need a GPU and are covered by tests/test_gpu_shard_group.py); nothing here computes anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from algoplonk_b200 import _lib, shard_group as sg
dist.init_process_group("gloo")
rank = dist.get_rank()

class FakeLib:
    def __init__(self): self.calls = []
    def b2p_shard_group_serve_proof(self, h, n): self.calls.append(("proof", h, n)); return 0
    def b2p_shard_group_serve_msm(self, h, n): self.calls.append(("msm", h, n)); return 0
    def b2p_last_error(self): return b""
fake = FakeLib()
_lib.load = lambda: fake

g = object.__new__(sg.ShardGroup)            # the constructor needs a GPU: wire the plumbing's fields by hand
g.curve, g.total, g.group, g.ntt_rows = "BN254", 1 << 12, None, 0
g.dist, g.rank, g.world, g.device = dist, rank, dist.get_world_size(), torch.device("cpu")
g.handle, g._attached = 4242, None
if rank == 0:
    g.announce(4096)                         # one proof
    g.announce(4096, 3)                      # three proofs announced at once
    g.announce_msm(1000, 2)
    g.stop()
    try:
        g.serve()
        raise SystemExit("rank 0 must not serve")
    except RuntimeError:
        pass
    print(json.dumps({{"rank0_calls": fake.calls}}))
else:
    served = g.serve()
    assert served == 6, served
    assert fake.calls == [("proof", 4242, 4096)] * 4 + [("msm", 4242, 1000)] * 2, fake.calls
dist.barrier()
dist.destroy_process_group()
"""


def test_control_messages_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    assert json.loads(line) == {"rank0_calls": []}          # rank 0 only announces; it never calls the serve entry points
