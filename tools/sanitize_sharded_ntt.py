"""Workload for compute-sanitizer (memcheck / racecheck / initcheck) over the domain-sharded NTT kernels
(csrc/ntt_shard.cuh): all ranks of worlds of 2, 4 and 8 simulated on one device, both curves, plain and coset,
ragged zero padding, plus the class in both exchange modes at world 1.

    compute-sanitizer --tool memcheck python tools/sanitize_sharded_ntt.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from algoplonk_b200 import _lib, sharded_ntt as sn  # noqa: E402
from test_sharded_ntt import _random_mont, _simulate  # noqa: E402

_lib.init(0)
for curve in ("BN254", "BLS12_381"):
    for world, logn in ((2, 9), (4, 12), (8, 13)):
        n = 1 << logn
        t = _random_mont(curve, n, logn, "cuda")
        for coset in (False, True):
            outs, back = _simulate(curve, n, world, t, coset)
            assert torch.equal(back, t)
        _simulate(curve, n, world, t, False, short=n // 2 + 3)
    for mode in ("staged", "p2p"):
        nt = sn.ShardedNtt(curve, 1 << 10, rank=0, world=1, mode=mode)
        t = _random_mont(curve, 1 << 10, 3, "cuda")
        assert torch.equal(nt.inverse(nt.forward(t, coset=True), coset=True), t)
        nt.free()
print("sanitize sharded-NTT workload ok")
