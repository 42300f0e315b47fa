"""MSM sharded over the point set across the GPUs of one box (BASELINE config 3).

An MSM is a sum, so rank g of G permanently owns a slice of the SRS -- the contiguous block [first_g, first_g +
count_g) (layout "blocks"), or the indices g, g + G, ... (layout "cyclic": the coefficient distribution of the
domain-sharded NTT, sharded_ntt.py, so a polynomial leaves a sharded inverse transform already placed for its
commitment) -- as its own windowed table in its own HBM and, per MSM, the matching slice of the scalar vector.  Every rank runs
the same kernels on its slice and produces ONE point; the G partial sums are exchanged with a single
`all_gather` (64 B / 96 B per rank: NCCL over NVLink on the GPU box, gloo in the CPU tests) and added
locally.  NCCL has no reduction over group elements, hence gather + local add instead of all_reduce.

One process per GPU (torchrun); torch.distributed is plumbing only -- the arithmetic is libb200plonk's.
Replaces nothing in the reference (gnark's MultiExp is single-process; /root/reference has no
collective): it is the multi-GPU form of kzg.Commit (setup/setup.go:11,13 import site).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

from . import _lib
from . import api


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, total): (first, count) of `rank`; sizes differ by at most 1."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard_indices(total: int, rank: int, world: int, layout: str = "blocks") -> range:
    """Indices of [0, total) that `rank` owns: "blocks" = the contiguous range of shard_range, "cyclic" =
    rank, rank + world, ... -- the coefficient distribution of the domain-sharded NTT (sharded_ntt.py)."""
    if layout == "blocks":
        first, count = shard_range(total, rank, world)
        return range(first, first + count)
    if layout == "cyclic":
        if world < 1 or not 0 <= rank < world:
            raise ValueError(f"bad rank {rank} of {world}")
        return range(rank, total, world)
    raise ValueError("layout must be 'blocks' or 'cyclic'")


def g1_sum(curve: str, points_raw: bytes) -> bytes:
    """Sum of affine points in G1Affine memory layout (host arithmetic of the library, no GPU needed)."""
    nb = 2 * api.FP_BYTES[curve]
    if len(points_raw) % nb:
        raise ValueError("point buffer has the wrong length")
    out = C.create_string_buffer(nb)
    buf = C.create_string_buffer(points_raw, len(points_raw)) if points_raw else None
    _lib.check(_lib.load().b2p_g1_sum(api.CURVE_ID[curve], buf, len(points_raw) // nb, out))
    return out.raw


def all_gather_points(curve: str, local_point_raw: bytes, group=None, device=None) -> bytes:
    """The one collective of the sharded MSM: every rank contributes one point, gets all G of them."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.frombuffer(bytearray(local_point_raw), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return bytes(out.cpu().numpy().tobytes())


class ShardedSRS:
    """Rank-local shard of a canonical-basis SRS of `total` points."""

    def __init__(self, curve: str, total: int, rank: int, world: int, handle: int, first: int, count: int,
                 layout: str = "blocks"):
        self.curve, self.total, self.rank, self.world = curve, total, rank, world
        self.handle, self.first, self.count, self.layout = handle, first, count, layout

    @classmethod
    def unsafe(cls, curve: str, total: int, rank: int, world: int, tau: int = api.TEST_TAU,
               layout: str = "blocks") -> "ShardedSRS":
        """[tau^j]_1 for j in this rank's slice (shard_indices), generated on this rank's GPU."""
        _lib.init()
        idx = shard_indices(total, rank, world, layout)
        h = C.c_void_p()
        t = C.create_string_buffer(api.fr_to_mont_bytes(curve, [tau]))
        if layout == "cyclic":
            _lib.check(_lib.load().b2p_srs_generate_unsafe_strided(api.CURVE_ID[curve], t, idx.start, idx.step,
                                                                   len(idx), C.byref(h)))
        else:
            _lib.check(_lib.load().b2p_srs_generate_unsafe_range(api.CURVE_ID[curve], t, idx.start, len(idx),
                                                                 C.byref(h)))
        return cls(curve, total, rank, world, h.value, idx.start, len(idx), layout)

    @classmethod
    def from_points(cls, curve: str, points_raw: bytes, rank: int, world: int, layout: str = "blocks") -> "ShardedSRS":
        """points_raw: the WHOLE SRS in G1Affine layout (e.g. pk.Kzg.G1); only this rank's slice is uploaded."""
        _lib.init()
        nb = 2 * api.FP_BYTES[curve]
        total = len(points_raw) // nb
        idx = shard_indices(total, rank, world, layout)
        h = C.c_void_p()
        mine = b"".join(points_raw[i * nb:(i + 1) * nb] for i in idx) if layout == "cyclic" else \
            points_raw[idx.start * nb:idx.stop * nb]
        buf = C.create_string_buffer(mine, len(idx) * nb)
        _lib.check(_lib.load().b2p_srs_load(api.CURVE_ID[curve], buf, len(idx), None, 0, C.byref(h)))
        return cls(curve, total, rank, world, h.value, idx.start, len(idx), layout)

    def local_msm_raw(self, scalars_mont: bytes) -> bytes:
        """This rank's partial sum over its slice; `scalars_mont`: the slice's scalars (host, Montgomery)."""
        out = C.create_string_buffer(2 * api.FP_BYTES[self.curve])
        n = len(scalars_mont) // 32
        buf = C.create_string_buffer(scalars_mont, len(scalars_mont)) if n else None
        _lib.check(_lib.load().b2p_msm_g1(self.handle, _lib.BASIS_CANONICAL, buf, n, out))
        return out.raw

    def local_msm_dev_raw(self, d_scalars_ptr: int, n: int) -> bytes:
        """Same with the slice's scalars already resident in this rank's HBM (device pointer)."""
        out = C.create_string_buffer(2 * api.FP_BYTES[self.curve])
        _lib.check(_lib.load().b2p_msm_g1_dev(self.handle, _lib.BASIS_CANONICAL, d_scalars_ptr, n, out))
        return out.raw

    def msm_raw(self, scalars_mont_slice: bytes, group=None, device=None) -> bytes:
        """Whole MSM: local partial sum, all_gather of the G points, local G-point add."""
        local = self.local_msm_raw(scalars_mont_slice)
        return g1_sum(self.curve, all_gather_points(self.curve, local, group, device))

    def msm(self, scalars_slice: Sequence[int], group=None, device=None) -> Optional[Tuple[int, int]]:
        raw = self.msm_raw(api.fr_to_mont_bytes(self.curve, scalars_slice), group, device)
        return api.points_from_mont_bytes(self.curve, raw)[0]

    def free(self):
        if self.handle:
            _lib.load().b2p_srs_free(self.handle)
            self.handle = None
