"""One proof over the GPUs of a box, native path: host plumbing of b2p_shard_group_* (csrc/shard_group.cuh).

The 9 kzg.Commit calls of plonk.Prove (algoplonk.go:89) are sharded over the point set (BASELINE
configs[2]).  Everything on the data path is in the library: the scalars are read by the other ranks'
Pippenger kernels straight out of rank 0's HBM over NVLink, the partial sums come back as peer stores, flags in
peer memory order the ranks.  What is left for the host language, and all this module does:

  * once: every rank creates its block of the SRS and its group handle, the 2 x 64-byte CUDA IPC handles of all
    ranks are all_gathered (torch.distributed: plumbing) and mapped (b2p_shard_group_connect);
  * per proof: ONE small broadcast "a proof of n rows starts" (or STOP); the other ranks answer it with
    b2p_shard_group_serve_proof(n), which queues their share of the proof's fixed commitment sequence and returns
    when it is done.

Round 1's Python commit hook (sharded_prover.py: a 32 n-byte broadcast and two host hops per commitment) stays as the
reference implementation the tests compare against; this is the path the benchmarks time.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _lib
from . import api
from . import sharded

OP_PROVE, OP_STOP, OP_MSM = 1, 2, 3


class ShardGroup:
    """This rank's end of the group.  Collective: every rank constructs it with the same arguments, then rank 0
    attaches its proving key (`attach`), then every rank calls `connect()`.

    ntt_rows = n: the proof's five size-4n transforms are spread over the ranks as well (world a power of two)."""

    def __init__(self, curve: str, total_points: int, group=None, tau: int = api.TEST_TAU,
                 srs_points: Optional[bytes] = None, device=None, ntt_rows: int = 0):
        import torch
        import torch.distributed as dist
        self.curve, self.total, self.group, self.ntt_rows = curve, total_points, group, ntt_rows
        self.dist = dist if dist.is_initialized() else None
        self.rank = dist.get_rank(group) if self.dist else 0
        self.world = dist.get_world_size(group) if self.dist else 1
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        _lib.init(self.device.index)
        lib = _lib.load()
        # this rank's block of the SRS: its own table, planned for its own size
        self.shard = (sharded.ShardedSRS.from_points(curve, srs_points[: total_points * 2 * api.FP_BYTES[curve]],
                                                     self.rank, self.world)
                      if srs_points is not None else
                      sharded.ShardedSRS.unsafe(curve, total_points, self.rank, self.world, tau))
        h = C.c_void_p()
        _lib.check(lib.b2p_shard_group_create(api.CURVE_ID[curve], self.world, self.rank, total_points,
                                              self.shard.handle, ntt_rows, C.byref(h)))
        self.handle = h.value
        self._attached = None

    def connect(self) -> None:
        """Collective, after rank 0's attach(): all_gather of the ranks' CUDA IPC handles (torch.distributed:
        plumbing), every rank maps what it needs of its peers."""
        import torch
        lib = _lib.load()
        if self.world == 1:
            return
        nb = _lib.SHARD_HANDLES * _lib.IPC_HANDLE_BYTES
        handles = C.create_string_buffer(nb)
        _lib.check(lib.b2p_shard_group_export(self.handle, handles))
        mine = torch.frombuffer(bytearray(handles.raw), dtype=torch.uint8).to(self.device)
        allh = torch.empty(self.world * nb, dtype=torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(allh, mine, group=self.group)
        buf = C.create_string_buffer(bytes(allh.cpu().numpy().tobytes()))
        _lib.check(lib.b2p_shard_group_connect(self.handle, buf))
        self.dist.barrier(group=self.group)       # every rank has mapped its peers before the first flag is written

    # ---- rank 0 ---------------------------------------------------------------------------------------------
    def attach(self, cc: "api.CompiledCircuit") -> None:
        """Rank 0, before connect(): route the commitments (and, with ntt_rows, the transforms) of the proving key
        `cc` through the group."""
        _lib.check(_lib.load().b2p_shard_group_attach(self.handle, cc.srs.handle, cc.handle))
        self._attached = cc

    def detach(self) -> None:
        if self._attached is not None and self.handle:
            _lib.check(_lib.load().b2p_shard_group_attach(self.handle, None, None))
        self._attached = None

    def _header(self, op: int, n: int, reps: int = 1):
        import torch
        h = torch.tensor([op, n, reps], dtype=torch.int64, device=self.device)
        if self.world > 1:
            src = self.dist.get_global_rank(self.group, 0) if self.group is not None else 0
            self.dist.broadcast(h, src=src, group=self.group)
        return int(h[0]), int(h[1]), int(h[2])

    def announce(self, n: int, reps: int = 1) -> None:
        """Rank 0, before b2p_prove on the attached key: the other ranks start serving `reps` proofs of n rows."""
        if self.world > 1:
            self._header(OP_PROVE, n, reps)

    def announce_msm(self, n: int, reps: int = 1) -> None:
        """Rank 0, before `reps` calls of msm_dev_raw(..., announce=False) with n scalars."""
        if self.world > 1:
            self._header(OP_MSM, n, reps)

    def stop(self) -> None:
        if self.rank == 0 and self.world > 1:
            self._header(OP_STOP, 0)

    def msm_dev_raw(self, d_scalars_ptr: int, n: int, announce: bool = True) -> bytes:
        """Rank 0: one commitment to n device-resident scalars over all the GPUs (kzg.Commit, multi-GPU form)."""
        if announce:
            self.announce_msm(n)
        out = C.create_string_buffer(2 * api.FP_BYTES[self.curve])
        _lib.check(_lib.load().b2p_shard_group_msm(self.handle, d_scalars_ptr, n, out))
        return out.raw

    # ---- ranks > 0 ----------------------------------------------------------------------------------------------
    def serve(self) -> int:
        """Take part in rank 0's proofs until it stops; returns how many proofs were served."""
        if self.rank == 0:
            raise RuntimeError("rank 0 proves; the other ranks serve")
        served = 0
        lib = _lib.load()
        while True:
            op, n, reps = self._header(0, 0)
            if op == OP_STOP:
                return served
            if op not in (OP_MSM, OP_PROVE):
                raise RuntimeError(f"bad header from rank 0: op={op} n={n}")
            for _ in range(max(1, reps)):
                _lib.check((lib.b2p_shard_group_serve_msm if op == OP_MSM else lib.b2p_shard_group_serve_proof)(self.handle, n))
                served += 1

    def free(self) -> None:
        self.detach()
        if self.handle:
            _lib.load().b2p_shard_group_free(self.handle)
            self.handle = None
        self.shard.free()


class ShardedProver:
    """api.CompiledCircuit whose commitments (and transforms, shard_ntt=True) run on every GPU of the process group
    (rank 0 proves), native path.

    every rank:  sp = ShardedProver(cs, curve, setup)
    rank 0:      proof = sp.prove_raw(L, R, O, blinding) ...; sp.close()
    ranks > 0:   sp.serve(); sp.close()
    """

    def __init__(self, cs, curve: str, setup_name: int, group=None, tau: int = api.TEST_TAU, shard_ntt: bool = False):
        from . import frontend as fe
        n = fe.build_trace(cs).n if hasattr(cs, "constraints") else int(cs)
        self.n = n
        self.grp = ShardGroup(curve, n + 3, group, tau, ntt_rows=n if shard_ntt else 0)
        self.rank, self.world = self.grp.rank, self.grp.world
        self.cc = None
        if self.rank == 0:
            self.cc = api.Compile(cs, curve, setup_name)          # full SRS on rank 0: setup commitments are local
            self.grp.attach(self.cc)
        self.grp.connect()

    def prove_raw(self, L: bytes, R: bytes, O: bytes, blinding: bytes) -> api.Proof:
        if self.rank != 0:
            raise RuntimeError("rank 0 proves; the other ranks serve")
        self.grp.announce(self.n)
        return self.cc.prove_raw(L, R, O, blinding)

    def serve(self) -> int:
        return self.grp.serve()

    def close(self) -> None:
        self.grp.stop()
        self.grp.free()
        if self.cc is not None:
            srs = self.cc.srs
            self.cc.free()
            srs.free()
            self.cc = None
