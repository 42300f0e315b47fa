"""One proof spread over the GPUs of a box: the prover's 9 MSMs sharded over the point set.

At 2^20 BN254 the MSMs are 26 of the 31 ms of a proof, and an MSM is a sum: rank g of G owns the
SRS points [first_g, first_g + count_g) (sharded.ShardedSRS, layout "blocks") and, per commitment, adds up its slice.
Rank 0 runs the prover (b2p_prove) with a *commit hook* on its SRS handle (b2p_srs_set_commit_hook): every
kzg.Commit inside the proof is handed to `ShardedCommitter.commit`, which

    1. broadcasts a 2-word header (op, n) and then the n device-resident scalars to every rank  (NCCL over NVLink;
       32 n bytes, 34 MB at 2^20 -- the one bulk exchange, rank 0 -> all),
    2. every rank runs its local Pippenger on its slice, on its own GPU             (b2p_msm_g1_dev on the shard),
    3. all_gathers one point per rank (64 B / 96 B) and adds the G partial sums      (b2p_g1_sum).

The other ranks sit in `serve()` and take part in the same three collectives until rank 0 sends STOP.  NTTs, the
grand product, the quotient and the openings stay on rank 0 (domain sharding of those is sharded_ntt.py; a prover
schedule over both is not built).  This is the *latency* mode of the multi-GPU story (one proof finishes sooner);
throughput is better served by replicas (`bench.py --gpus N`), DESIGN.md section 7.

Replaces nothing in the reference (gnark's prover is single-process): it is kzg.Commit (setup/setup.go:11,13 import
site; called 9 times by plonk.Prove, algoplonk.go:89) in multi-GPU form.  torch.distributed is plumbing; the
arithmetic is libb200plonk's.  CPU tests (gloo, world 2) drive the same three collectives with the local sums
supplied by the oracle; the GPU path needs one process per GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional

from . import _lib
from . import api
from . import sharded

OP_COMMIT, OP_STOP = 1, 2


class _DevicePtr:
    """The scalars a commit hook is handed: a raw device pointer with the two tensor methods the single-rank path
    uses (data_ptr, len in bytes)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = int(ptr or 0), nbytes

    def data_ptr(self) -> int:
        return self.ptr

    def __len__(self) -> int:
        return self.nbytes


def slice_for(total_points: int, rank: int, world: int, n_scalars: int):
    """(offset, count) of the scalars rank `rank` multiplies: its block of the SRS, cut at the vector's length."""
    first, count = sharded.shard_range(total_points, rank, world)
    lo = min(first, n_scalars)
    hi = min(first + count, n_scalars)
    return lo, hi - lo


class ShardedCommitter:
    """The three collectives of one sharded commitment; every rank constructs one with the same arguments.

    local_msm(scalars, offset, count) -> bytes : this rank's partial sum (one G1Affine, memory layout) over the
        `count` scalars starting at element `offset` of `scalars` (a uint8 tensor of 32-byte Montgomery elements, on
        `device`).  Default: b2p_msm_g1_dev on `shard` (needs a GPU); tests pass the oracle's.
    """

    def __init__(self, curve: str, total_points: int, shard: Optional[sharded.ShardedSRS] = None, group=None,
                 device=None, local_msm: Optional[Callable] = None):
        import torch
        import torch.distributed as dist
        self.curve, self.total, self.group = curve, total_points, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.shard = shard
        if shard is not None and (shard.layout != "blocks" or shard.total != total_points or shard.rank != self.rank
                                  or shard.world != self.world or shard.curve != curve):
            raise ValueError("the SRS shard must be this rank's block of the same SRS over the same world")
        self._local = local_msm if local_msm is not None else self._cuda_local_msm
        if local_msm is None and shard is None:
            raise ValueError("either an SRS shard or a local_msm function is needed")
        self.commits = 0

    # -- step 2 on the GPU --------------------------------------------------------------------------------------
    def _cuda_local_msm(self, scalars, offset: int, count: int) -> bytes:
        import torch
        if count == 0:
            return bytes(2 * api.FP_BYTES[self.curve])
        torch.cuda.current_stream(self.device).synchronize()      # the broadcast has landed before the MSM reads it
        return self.shard.local_msm_dev_raw(scalars.data_ptr() + 32 * offset, count)

    # -- the collectives ---------------------------------------------------------------------------------------------
    def _header(self, op: int, n: int):
        import torch
        import torch.distributed as dist
        h = torch.tensor([op, n], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.broadcast(h, src=self._src(), group=self.group)
        return int(h[0]), int(h[1])

    def _src(self) -> int:
        import torch.distributed as dist
        return dist.get_global_rank(self.group, 0) if self.group is not None else 0

    def _exchange(self, scalars, n: int) -> bytes:
        """Steps 1b-3 with `scalars` = the full vector on rank 0, an empty receive buffer elsewhere."""
        import torch.distributed as dist
        if self.world > 1:
            dist.broadcast(scalars, src=self._src(), group=self.group)
        off, cnt = slice_for(self.total, self.rank, self.world, n)
        local = self._local(scalars, off, cnt)
        self.commits += 1
        if self.world == 1:
            return local
        allp = sharded.all_gather_points(self.curve, local, self.group, self.device if self.device.type == "cuda" else None)
        return sharded.g1_sum(self.curve, allp)

    def commit(self, scalars, n: int) -> bytes:
        """Rank 0: commitment to the n scalars in `scalars` (uint8 tensor, 32 n bytes, on `device`)."""
        if self.rank != 0:
            raise RuntimeError("commit() is rank 0's side; the other ranks run serve()")
        if n > self.total:
            raise ValueError("more scalars than SRS points")
        self._header(OP_COMMIT, n)
        return self._exchange(scalars, n)

    def stop(self) -> None:
        if self.rank == 0 and self.world > 1:
            self._header(OP_STOP, 0)

    def serve(self) -> int:
        """Ranks > 0: take part in rank 0's commitments until it stops; returns how many were served."""
        import torch
        if self.rank == 0:
            raise RuntimeError("serve() is for the ranks that do not run the prover")
        buf = torch.empty(32 * self.total, dtype=torch.uint8, device=self.device)
        while True:
            op, n = self._header(0, 0)
            if op == OP_STOP:
                return self.commits
            if op != OP_COMMIT or n > self.total:
                raise RuntimeError(f"bad header from rank 0: op={op} n={n}")
            self._exchange(buf[: 32 * n], n)


class CommitHook:
    """b2p_srs_set_commit_hook on `srs` (an api.SRS): its commitments go through `committer.commit`."""

    def __init__(self, srs, committer: ShardedCommitter, device):
        import torch
        self.srs, self.error = srs, None
        nb = 2 * api.FP_BYTES[srs.curve]
        stage = torch.empty(32 * committer.total, dtype=torch.uint8, device=device) if committer.world > 1 else None

        def hook(_ctx, d_scalars, n, out):
            try:
                if committer.world == 1:                             # nothing to broadcast: the pointer is enough
                    t = _DevicePtr(d_scalars, 32 * n)
                else:                                                # stage into the tensor that is broadcast
                    t = stage[: 32 * n]
                    _lib.check(_lib.load().b2p_device_copy(t.data_ptr(), d_scalars, 32 * n))
                C.memmove(out, committer.commit(t, n), nb)
                return 0
            except BaseException as e:  # noqa: BLE001 -- must not unwind into C; the caller re-raises it
                self.error = e
                return 1
        self._cb = _lib.COMMIT_FN(hook)                              # kept alive as long as the hook is installed
        _lib.check(_lib.load().b2p_srs_set_commit_hook(srs.handle, C.cast(self._cb, C.c_void_p), None))

    def remove(self) -> None:
        if self._cb is not None and self.srs.handle:
            _lib.check(_lib.load().b2p_srs_set_commit_hook(self.srs.handle, None, None))
        self._cb = None


class ShardedProver:
    """api.CompiledCircuit whose commitments run on every GPU of the process group (rank 0 proves).

    Every rank:  sp = ShardedProver(cs, curve, setup)       # collective: each rank generates / loads its SRS block
    rank 0:      proof = sp.Prove(L, R, O, blinding); ...; sp.close()
    ranks > 0:   sp.serve()                                  # returns after rank 0's close()
    """

    def __init__(self, cs, curve: str, setup_name: int, group=None, tau: int = api.TEST_TAU, srs_points: Optional[bytes] = None):
        import torch
        import torch.distributed as dist
        from . import frontend as fe
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        device = torch.device("cuda", torch.cuda.current_device())
        n = fe.build_trace(cs).n if self.rank != 0 else None
        self.cc = None
        if self.rank == 0:
            srs = api.SRS.from_points(curve, srs_points) if srs_points is not None else None
            self.cc = api.Compile(cs, curve, setup_name, srs=srs)     # VK commitments are made before the hook is set
            n = self.cc.trace.n
        total = n + 3
        self.shard = (sharded.ShardedSRS.from_points(curve, srs_points[: total * 2 * api.FP_BYTES[curve]], self.rank, self.world)
                      if srs_points is not None else sharded.ShardedSRS.unsafe(curve, total, self.rank, self.world, tau))
        self.committer = ShardedCommitter(curve, total, self.shard, group, device)
        self._hook = None
        if self.rank == 0:
            self._install_hook(device)

    def _install_hook(self, device) -> None:
        self._hook = CommitHook(self.cc.srs, self.committer, device)

    def Prove(self, L, R, O, blinding, pi2=(), bsb22_points=()):
        if self.rank != 0:
            raise RuntimeError("rank 0 proves; the other ranks run serve()")
        try:
            return self.cc.Prove(L, R, O, blinding, pi2, bsb22_points)
        except _lib.B200PlonkError:
            if self._hook.error is not None:
                raise self._hook.error
            raise

    def serve(self) -> int:
        return self.committer.serve()

    def close(self) -> None:
        if self.rank == 0:
            self.committer.stop()
            if self.cc is not None and self.cc.handle:
                self._hook.remove()
                self.cc.free()
                self.cc.srs.free()
        self.shard.free()
