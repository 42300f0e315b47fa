"""Domain-sharded NTT (algoplonk_b200/sharded_ntt.py, csrc/ntt_shard.cuh; SURVEY 8e-3, BASELINE config 5).

CPU: (1) the per-column bodies of the combine / split kernels (csrc/ntt_shard_math.cuh) run on the host for all
ranks of a world of 1, 2, 4, 8 and checked against the big-int oracle; (2) the exchange logic of ShardedNtt on
gloo with world_size 2 and 4, the four local steps supplied by the oracle (no GPU compute on this path).
GPU (-m gpu): the kernels through the C ABI, all ranks of a world simulated on one device (every rank's chunks
point into the other ranks' buffers, exactly what mapped peer memory looks like), bit-exact against b2p_ntt;
and two processes sharing one GPU through CUDA IPC (mode="p2p").
"""
import ctypes as C
import json
import os
import random
import subprocess
import sys

import pytest

import helpers as H
from algoplonk_b200 import _lib, api, sharded_ntt as sn
from oracle import plonk_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CURVES = ("BN254", "BLS12_381")


# ---- the distribution --------------------------------------------------------------------------------------
def test_distribution_helpers():
    for n, world in ((4, 2), (16, 4), (64, 8), (32, 1)):
        seen_c, seen_e = [], []
        for r in range(world):
            seen_c += list(sn.local_indices(n, r, world))
            ks = sn.local_eval_exponents(n, r, world)
            assert len(ks) == n // world
            # "index mod world" distribution of the evaluations: k mod world = brev(rank)
            assert {k % world for k in ks} == {sn.bit_reverse(r, world.bit_length() - 1)}
            seen_e += ks
        assert sorted(seen_c) == sorted(seen_e) == list(range(n))
    for bad in ((6, 2), (8, 3), (8, 4), (0, 1)):
        with pytest.raises(ValueError):
            sn._check_shape(*bad)


def _expected_blocks(cv, coeffs, world, coset):
    """evaluations in the sharded layout: block r = [A(w^k) for k in local_eval_exponents(r)]"""
    n = len(coeffs)
    w = po.domain_generator(cv, n)
    ev = po.coset_ntt(cv, coeffs, w, cv.coset_shift) if coset else po.ntt(cv, coeffs, w)
    return [[ev[k] for k in sn.local_eval_exponents(n, r, world)] for r in range(world)]


# ---- (1) the kernels' per-column code on the host ------------------------------------------------------------
@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("ns") / "ntt_shard_shim.so"
    src = os.path.join(ROOT, "tests", "csrc", "ntt_shard_shim.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DB2P_HOST_EMULATE_DEVICE_MUL", "-x", "c++", src,
                    "-o", str(out)], check=True)
    return C.CDLL(str(out))


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("logg", [0, 1, 2, 3])
def test_shard_math_on_host_matches_oracle(shim, curve, logg):
    cv = po.CURVES[curve]
    world = 1 << logg
    for logn in sorted({2 * logg, 2 * logg + 1, 7}):
        n = 1 << logn
        a = H.scalars_uniform(cv.r, n, 31 * logn + logg)
        for coset in (False, True):
            src = C.create_string_buffer(api.fr_to_mont_bytes(curve, a), 32 * n)
            dst = C.create_string_buffer(32 * n)
            assert shim.hs_transform(cv.cid, logn, logg, 0, src, dst, int(coset)) == 0
            got = api.fr_from_mont_bytes(curve, dst.raw)
            want = _expected_blocks(cv, a, world, coset)
            assert got == [v for blk in want for v in blk], (logn, coset)
            back = C.create_string_buffer(32 * n)
            assert shim.hs_transform(cv.cid, logn, logg, 1, dst, back, int(coset)) == 0
            assert api.fr_from_mont_bytes(curve, back.raw) == a, (logn, coset)


# ---- (2) exchange logic on gloo, local steps from the oracle --------------------------------------------------
WORKER = r"""
import json, os, random, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch
import torch.distributed as dist
from algoplonk_b200 import api, sharded_ntt as sn
from oracle import plonk_oracle as po
from test_sharded_ntt import OracleSteps, to_tensor, from_tensor, _expected_blocks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
checked = 0
for curve in ("BN254", "BLS12_381"):
    cv = po.CURVES[curve]
    for n in (world * world, 64):
        rng = random.Random(n)                         # same polynomial on every rank
        a = [rng.randrange(cv.r) for _ in range(n)]
        nt = sn.ShardedNtt(curve, n, mode="staged", steps=OracleSteps(curve, n, world, rank))
        mine = [a[i] for i in sn.local_indices(n, rank, world)]
        for coset in (False, True):
            ev = nt.forward(to_tensor(curve, mine), coset=coset)
            assert from_tensor(curve, ev) == _expected_blocks(cv, a, world, coset)[rank], (curve, n, coset)
            assert from_tensor(curve, nt.inverse(ev, coset=coset)) == mine
            checked += 1
        # ragged: fewer coefficients than domain points (zero padded), different counts per rank
        short = a[: n // 2 + 1]
        mine_s = [short[i] for i in range(rank, len(short), world)]
        ev = nt.forward(to_tensor(curve, mine_s))
        assert from_tensor(curve, ev) == _expected_blocks(cv, short + [0] * (n - len(short)), world, False)[rank]
dist.barrier()
if rank == 0:
    print(json.dumps({{"ok": True, "world": world, "checked": checked}}))
dist.destroy_process_group()
"""


def to_tensor(curve, values, device="cpu"):
    import torch
    raw = bytearray(api.fr_to_mont_bytes(curve, values))
    if not raw:
        return torch.empty((0, 4), dtype=torch.int64, device=device)
    return torch.frombuffer(raw, dtype=torch.int64).reshape(-1, 4).clone().to(device)


def from_tensor(curve, t):
    return api.fr_from_mont_bytes(curve, t.cpu().contiguous().numpy().tobytes())


class OracleSteps:
    """The four local steps of one rank restated with Python integers, straight from the definition
    A(w^k) = sum_r w^(r k) C_r(k mod n/G): no butterflies shared with the product.  Test-only."""

    def __init__(self, curve, n, world, rank):
        self.curve, self.cv, self.n, self.world, self.rank = curve, po.CURVES[curve], n, world, rank
        self.ln, self.chunk = n // world, n // world // world
        self.g, self.ll = world.bit_length() - 1, (n // world).bit_length() - 1
        self.w = po.domain_generator(self.cv, n)

    def _k1(self, q_low):
        return sn.bit_reverse(self.rank * self.chunk + q_low, self.ll)

    def forward_local(self, coeffs, local_len, coset, x):
        r, cv = self.cv.r, self.cv
        a = from_tensor(self.curve, coeffs)[:local_len] + [0] * (self.ln - local_len)
        if coset:
            a = [v * pow(cv.coset_shift, j * self.world + self.rank, r) % r for j, v in enumerate(a)]
        c = po.ntt(cv, a, pow(self.w, self.world, r)) if self.ln > 1 else a
        x.copy_(to_tensor(self.curve, [c[sn.bit_reverse(q, self.ll)] for q in range(self.ln)]))

    def forward_combine(self, chunks, out):
        r = self.cv.r
        cols = [from_tensor(self.curve, c) for c in chunks]
        res = [0] * self.ln
        for q in range(self.chunk):
            for t in range(self.world):
                k = self._k1(q) + self.ln * sn.bit_reverse(t, self.g)
                res[(q << self.g) | t] = sum(cols[s][q] * pow(self.w, s * k, r) for s in range(self.world)) % r
        out.copy_(to_tensor(self.curve, res))

    def inverse_split(self, evals, chunks):
        r = self.cv.r
        e = from_tensor(self.curve, evals)
        wi = pow(self.w, -1, r)
        for d in range(self.world):
            col = []
            for q in range(self.chunk):
                ks = [self._k1(q) + self.ln * sn.bit_reverse(t, self.g) for t in range(self.world)]
                col.append(sum(e[(q << self.g) | t] * pow(wi, d * ks[t], r) for t in range(self.world)) % r)
            chunks[d].copy_(to_tensor(self.curve, col))

    def inverse_local(self, x, coset, out=None):
        r, cv = self.cv.r, self.cv
        xs = from_tensor(self.curve, x)
        c = [xs[sn.bit_reverse(k, self.ll)] for k in range(self.ln)]
        a = po.intt(cv, c, pow(self.w, self.world, r)) if self.ln > 1 else c     # includes 1/(n/G)
        ginv = pow(self.world, -1, r)
        a = [v * ginv % r for v in a]
        if coset:
            si = pow(cv.coset_shift, -1, r)
            a = [v * pow(si, j * self.world + self.rank, r) % r for j, v in enumerate(a)]
        (x if out is None else out).copy_(to_tensor(self.curve, a))


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_logic_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res == {"ok": True, "world": world, "checked": 8}


def test_world_one_needs_no_process_group():
    """world = 1 through the same class (oracle steps): the exchange degenerates to a copy."""
    cv = po.CURVES["BN254"]
    a = H.scalars_uniform(cv.r, 16, 3)
    nt = sn.ShardedNtt("BN254", 16, rank=0, world=1, steps=OracleSteps("BN254", 16, 1, 0))
    ev = nt.forward(to_tensor("BN254", a))
    assert from_tensor("BN254", ev) == _expected_blocks(cv, a, 1, False)[0]
    assert from_tensor("BN254", nt.inverse(ev)) == a
    with pytest.raises(ValueError):
        nt.forward(to_tensor("BN254", a + [1]))
    with pytest.raises(ValueError):
        nt.inverse(to_tensor("BN254", a[:8]))
    with pytest.raises(ValueError):
        sn.ShardedNtt("BN254", 16, rank=0, world=3, steps=object())


def test_argument_errors_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.b2p_ntt_shard_create(9, 16, 2, 0, C.byref(h)) == -1
    assert b"unsupported curve" in lib.b2p_last_error()
    assert lib.b2p_ntt_shard_forward_local(None, None, 0, 0, None, None) == -1
    assert lib.b2p_ntt_shard_local_size(None) == 0 and lib.b2p_ntt_shard_chunk_size(None) == 0
    assert lib.b2p_peer_open(None, None) == -1


# ---- GPU -----------------------------------------------------------------------------------------------------
def _random_mont(curve, count, seed, device):
    """Random canonical Montgomery representatives as a (count, 4) int64 tensor: the top limb is drawn below the
    modulus' top limb, so every value is < r; no big-int conversion at 2^20 elements."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = torch.randint(-(1 << 63), (1 << 63) - 1, (count, 4), generator=g, dtype=torch.int64)
    top = po.CURVES[curve].r >> 192
    t[:, 3] = torch.randint(0, top, (count,), generator=g, dtype=torch.int64)
    return t.to(device)


def _single_gpu_ntt_bytes(curve, t_natural, inverse=False, coset=False):
    """b2p_ntt on the same Montgomery bytes (natural order in and out)."""
    raw = bytearray(t_natural.cpu().contiguous().numpy().tobytes())
    buf = (C.c_char * len(raw)).from_buffer(raw)
    flags = (_lib.NTT_INVERSE if inverse else 0) | (_lib.NTT_COSET if coset else 0)
    _lib.check(_lib.load().b2p_ntt(api.CURVE_ID[curve], buf, len(raw) // 32, flags))
    import torch
    return torch.frombuffer(raw, dtype=torch.int64).reshape(-1, 4).clone()


def _simulate(curve, n, world, coeffs_nat, coset, short=None):
    """All ranks of a world on one GPU through the C ABI: rank d's chunk pointers point into the other ranks'
    exchange buffers (what mapped peer memory looks like).  Returns (blocks of evaluations, coefficients back)."""
    import torch
    dev = coeffs_nat.device
    steps = [sn.CudaSteps(curve, n, world, r) for r in range(world)]
    ln, chunk = n // world, n // world // world
    try:
        xs = [torch.empty((ln, 4), dtype=torch.int64, device=dev) for _ in range(world)]
        for r in range(world):
            mine = coeffs_nat[r::world].contiguous()
            if short is not None:
                mine = mine[: len(range(r, short, world))].contiguous()
            steps[r].forward_local(mine, mine.shape[0], coset, xs[r])
        outs = []
        for d in range(world):
            out = torch.empty((ln, 4), dtype=torch.int64, device=dev)
            steps[d].forward_combine([xs[s][d * chunk:(d + 1) * chunk] for s in range(world)], out)
            outs.append(out)
        # inverse: rank s pushes chunk r into rank r's buffer at slot s
        ys = [torch.full((ln, 4), -1, dtype=torch.int64, device=dev) for _ in range(world)]
        for s in range(world):
            steps[s].inverse_split(outs[s], [ys[r][s * chunk:(s + 1) * chunk] for r in range(world)])
        back = torch.empty((n, 4), dtype=torch.int64, device=dev)
        for r in range(world):
            res = torch.empty((ln, 4), dtype=torch.int64, device=dev)
            steps[r].inverse_local(ys[r], coset, res)
            assert torch.equal(res, ys[r])
            back[r::world] = res
        torch.cuda.synchronize()
        return outs, back
    finally:
        for s in steps:
            s.free()


@pytest.mark.gpu
@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_gpu_simulated_ranks_small_vs_bigint_oracle(gpu, curve, world):
    import torch
    cv = po.CURVES[curve]
    g = world.bit_length() - 1
    for logn in sorted({2 * g, 2 * g + 1, 6, 9}):
        n = 1 << logn
        a = H.scalars_uniform(cv.r, n, 7 * logn + world)
        t = to_tensor(curve, a, "cuda")
        for coset in (False, True):
            outs, back = _simulate(curve, n, world, t, coset)
            want = _expected_blocks(cv, a, world, coset)
            for r in range(world):
                assert from_tensor(curve, outs[r]) == want[r], (logn, coset, r)
            assert torch.equal(back, t)
        if n >= 8:
            short = n // 2 + 3                 # zero padding, ragged local lengths
            outs, _ = _simulate(curve, n, world, t, False, short=short)
            want = _expected_blocks(cv, a[:short] + [0] * (n - short), world, False)
            assert [from_tensor(curve, o) for o in outs] == want


@pytest.mark.gpu
@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("world,logn", [(2, 14), (8, 17), (4, 20), (8, 21)])
def test_gpu_simulated_ranks_match_single_gpu_ntt(gpu, curve, world, logn):
    """Config sizes (BASELINE config 5 is 2^21 on 8 GPUs): the sharded transform equals b2p_ntt bit for bit."""
    import torch
    n = 1 << logn
    t = _random_mont(curve, n, logn, "cuda")
    bits = logn
    idx = torch.tensor([sn.bit_reverse(p, bits) for p in range(n)], dtype=torch.int64) if logn <= 17 else None
    for coset in ((False, True) if logn <= 17 else (True,)):
        outs, back = _simulate(curve, n, world, t, coset)
        assert torch.equal(back, t)
        ref = _single_gpu_ntt_bytes(curve, t, coset=coset)            # natural order
        got = torch.cat(outs).cpu()                                  # bit-reversed order
        if idx is None:
            # bit reversal of 2^20+ indices without a Python loop: reverse the bits of arange
            p = torch.arange(n, dtype=torch.int64)
            idx = torch.zeros(n, dtype=torch.int64)
            for b in range(bits):
                idx |= ((p >> b) & 1) << (bits - 1 - b)
        assert torch.equal(got, ref[idx])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["staged", "p2p"])
def test_gpu_class_world_one(gpu, mode):
    import torch
    curve, n = "BLS12_381", 1 << 12
    t = _random_mont(curve, n, 5, "cuda")
    nt = sn.ShardedNtt(curve, n, rank=0, world=1, mode=mode)
    try:
        for coset in (False, True):
            ev = nt.forward(t, coset=coset)
            ref = _single_gpu_ntt_bytes(curve, t, coset=coset)
            idx = torch.tensor([sn.bit_reverse(p, 12) for p in range(n)], dtype=torch.int64)
            assert torch.equal(ev.cpu(), ref[idx])
            assert torch.equal(nt.inverse(ev, coset=coset), t)
    finally:
        nt.free()


IPC_WORKER = r"""
import json, os, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch
import torch.distributed as dist
from algoplonk_b200 import sharded_ntt as sn
from test_sharded_ntt import _random_mont, _single_gpu_ntt_bytes
torch.cuda.set_device(0)                    # both ranks share the one GPU: CUDA IPC between two processes
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
curve, logn = "BN254", 13
n = 1 << logn
t = _random_mont(curve, n, 77, "cuda")       # same polynomial on every rank
nt = sn.ShardedNtt(curve, n, mode="p2p")
idx = torch.tensor([sn.bit_reverse(p, logn) for p in range(n)], dtype=torch.int64)
ln = n // world
for it in range(3):                          # three transforms each way: both buffer parities are reused
    for coset in (False, True):
        mine = t[rank::world].contiguous()
        ev = nt.forward(mine, coset=coset)
        ref = _single_gpu_ntt_bytes(curve, t, coset=coset)[idx]
        assert torch.equal(ev.cpu(), ref[rank * ln:(rank + 1) * ln]), ("forward", it, coset)
        assert torch.equal(nt.inverse(ev, coset=coset), mine), ("inverse", it, coset)
nt.free()
dist.barrier()
if rank == 0:
    print(json.dumps({{"ok": True, "world": world}}))
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_gpu_p2p_two_processes_over_cuda_ipc(gpu, tmp_path):
    script = tmp_path / "ipc_worker.py"
    script.write_text(IPC_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    assert json.loads(out.stdout.strip().splitlines()[-1]) == {"ok": True, "world": 2}
