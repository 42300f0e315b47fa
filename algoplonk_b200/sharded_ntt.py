"""NTT with the domain sharded over the GPUs of one box (BASELINE config 5).

The multi-GPU form of gnark-crypto's fft.Domain.FFT / FFTInverse (the reference is single-process
and has no counterpart in /root/reference).  One process per GPU (torchrun); world = 1, 2, 4 or 8.

Distribution -- one exchange per transform, in either direction:
  coefficients : cyclic -- rank r holds a[j*world + r] at local index j
  evaluations  : the bit-reversed (DIF) order cut into world blocks -- rank r holds A(omega^brev(p)) for
                 p in [r*n/world, (r+1)*n/world)
Both are "index mod world" distributions, so coefficient-wise / point-wise kernels and a point-set-sharded MSM
over a cyclically sharded SRS need no communication between transforms.

The arithmetic is libb200plonk's (csrc/ntt_shard.cuh): forward = local DIF passes | exchange | one combine
kernel (twiddle + size-world butterflies in registers); inverse = one split kernel | exchange | local DIT passes.
Two ways to run the exchange:
  mode="p2p"    : every rank maps the other ranks' exchange buffers (CUDA IPC, NVLink peer memory).  The combine
                  kernel LOADS its chunks straight from the peers and the split kernel STORES them straight
                  into the peers: the transposition is the kernel's access pattern, no all_to_all is launched.
                  The only collective is a one-element all_reduce on the launching stream that orders the ranks
                  (stream-ordered with NCCL; exchange buffers are double-buffered so one barrier per
                  transform suffices).
  mode="staged" : `all_to_all_single` (NCCL over NVLink on the GPU box, gloo in the CPU tests) between the two
                  launches, into / out of a staging buffer.

torch.distributed is plumbing only.  Without the CUDA library or a GPU, construction raises: there is no CPU
fallback (the CPU tests inject the oracle for the four local steps to check the exchange logic).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

from . import _lib
from . import api

FR_BYTES = 32
_WORDS = 4          # an Fr is 4 int64 words of a torch tensor


def local_indices(n: int, rank: int, world: int) -> range:
    """Natural-order coefficient indices rank `rank` holds (cyclic distribution)."""
    return range(rank, n, world)


def bit_reverse(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def local_eval_exponents(n: int, rank: int, world: int) -> List[int]:
    """k such that local evaluation p of rank `rank` is A(omega^k): k = brev(rank*n/world + p)."""
    bits = n.bit_length() - 1
    ln = n // world
    return [bit_reverse(rank * ln + p, bits) for p in range(ln)]


def _check_shape(n: int, world: int) -> None:
    if n < 1 or n & (n - 1):
        raise ValueError("NTT length must be a power of two")
    if world not in (1, 2, 4, 8):
        raise ValueError("world size must be 1, 2, 4 or 8")
    if n < world * world:
        raise ValueError("sharded NTT needs n >= world^2")


def _ptr(buf) -> int:
    """Device pointer of a torch CUDA tensor (contiguous int64 words) or a raw pointer passed through."""
    if isinstance(buf, int):
        return buf
    if not buf.is_cuda or not buf.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return buf.data_ptr()


class CudaSteps:
    """The four local steps on this rank's GPU: include/b200plonk.h b2p_ntt_shard_*.  Launches go to torch's
    current stream."""

    def __init__(self, curve: str, n: int, world: int, rank: int):
        _lib.init()                       # raises without a usable GPU
        self.curve, self.n, self.world, self.rank = curve, n, world, rank
        h = C.c_void_p()
        _lib.check(_lib.load().b2p_ntt_shard_create(api.CURVE_ID[curve], n, world, rank, C.byref(h)))
        self.handle = h.value
        self.local_n = _lib.load().b2p_ntt_shard_local_size(self.handle)
        self.chunk = _lib.load().b2p_ntt_shard_chunk_size(self.handle)

    @staticmethod
    def _stream() -> int:
        import torch
        return torch.cuda.current_stream().cuda_stream

    def _chunk_array(self, chunks: Sequence):
        if len(chunks) != self.world:
            raise ValueError("one chunk per rank expected")
        return (C.c_void_p * self.world)(*[_ptr(c) for c in chunks])

    def forward_local(self, coeffs, local_len: int, coset: bool, x) -> None:
        _lib.check(_lib.load().b2p_ntt_shard_forward_local(
            self.handle, _ptr(coeffs) if local_len else None, local_len, _lib.NTT_COSET if coset else 0, _ptr(x),
            self._stream()))

    def forward_combine(self, chunks: Sequence, out) -> None:
        _lib.check(_lib.load().b2p_ntt_shard_forward_combine(self.handle, self._chunk_array(chunks), _ptr(out),
                                                             self._stream()))

    def inverse_split(self, evals, chunks: Sequence) -> None:
        _lib.check(_lib.load().b2p_ntt_shard_inverse_split(self.handle, _ptr(evals), self._chunk_array(chunks),
                                                           self._stream()))

    def inverse_local(self, x, coset: bool, out=None) -> None:
        flags = _lib.NTT_INVERSE | (_lib.NTT_COSET if coset else 0)
        _lib.check(_lib.load().b2p_ntt_shard_inverse_local(self.handle, _ptr(x), flags,
                                                           None if out is None else _ptr(out), self._stream()))

    def free(self) -> None:
        if self.handle:
            _lib.load().b2p_ntt_shard_free(self.handle)
            self.handle = None


class ShardedNtt:
    """One rank of a domain-sharded NTT of size n.  Buffers are torch int64 tensors of shape (count, 4): Fr in
    gnark's in-memory layout (4 little-endian u64 limbs, Montgomery form), the same bytes b2p_ntt takes."""

    def __init__(self, curve: str, n: int, rank: Optional[int] = None, world: Optional[int] = None, group=None,
                 mode: str = "staged", steps=None, device=None):
        import torch.distributed as dist
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        _check_shape(n, world)
        if not 0 <= rank < world:
            raise ValueError(f"bad rank {rank} of {world}")
        if mode not in ("staged", "p2p"):
            raise ValueError("mode must be 'staged' or 'p2p'")
        self.curve, self.n, self.rank, self.world, self.group, self.mode = curve, n, rank, world, group, mode
        self.local_n = n // world
        self.chunk = self.local_n // world
        self.steps = steps if steps is not None else CudaSteps(curve, n, world, rank)
        self.device = device
        if self.device is None:
            import torch
            self.device = torch.device("cuda", torch.cuda.current_device()) if steps is None else torch.device("cpu")
        self._seq = 0
        self._own = 0                 # p2p: this rank's exchange memory, 2 buffers of n/world Fr
        self._peer: List[int] = []    # p2p: every rank's exchange memory as mapped into this process
        self._opened: List[int] = []
        self._flag = None
        if mode == "p2p":
            self._setup_p2p()

    # ---- buffers -------------------------------------------------------------------------------------------
    def empty(self, count: Optional[int] = None):
        import torch
        return torch.empty((self.local_n if count is None else count, _WORDS), dtype=torch.int64, device=self.device)

    def _chunks_of(self, buf) -> list:
        return [buf[r * self.chunk:(r + 1) * self.chunk] for r in range(self.world)]

    def _check_buf(self, t, max_count: int, exact: bool):
        if t.dim() != 2 or t.shape[1] != _WORDS or str(t.dtype) != "torch.int64":
            raise ValueError("expected an int64 tensor of shape (count, 4)")
        if t.shape[0] > max_count or (exact and t.shape[0] != max_count):
            raise ValueError(f"expected {'exactly' if exact else 'at most'} {max_count} local elements, got {t.shape[0]}")
        return t.contiguous()

    # ---- p2p plumbing --------------------------------------------------------------------------------------
    def _setup_p2p(self) -> None:
        import torch
        import torch.distributed as dist
        lib = _lib.load()
        ptr, handle = C.c_void_p(), C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        failure = ""
        if lib.b2p_peer_alloc(2 * self.local_n * FR_BYTES, C.byref(ptr), handle) != 0:
            failure = f"rank {self.rank} cannot allocate exchange memory: " + lib.b2p_last_error().decode()
            if self.world == 1:
                raise RuntimeError(failure)
        self._own = ptr.value or 0
        self._peer = [0] * self.world
        self._peer[self.rank] = self._own
        if self.world > 1:
            handles: list = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=self.group)
            for r, raw in enumerate(handles):
                if r == self.rank or failure:
                    continue
                p = C.c_void_p()
                if lib.b2p_peer_open(C.create_string_buffer(raw, _lib.IPC_HANDLE_BYTES), C.byref(p)) != 0:
                    failure = f"rank {self.rank} cannot map rank {r}'s exchange buffer: " + \
                        lib.b2p_last_error().decode()
                    break
                self._peer[r] = p.value
                self._opened.append(p.value)
            # a rank that failed must not leave the others waiting in the barrier: agree on the outcome first
            failures: list = [None] * self.world
            dist.all_gather_object(failures, failure, group=self.group)
            if any(failures):
                self.free()
                raise RuntimeError("; ".join(f for f in failures if f))
            self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._barrier()

    def _barrier(self) -> None:
        """Orders the ranks between the launch that writes exchange memory and the launch that reads it."""
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self._flag, group=self.group)       # stream-ordered: no host synchronisation
        else:
            torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)

    def _exchange_ptrs(self, parity: int):
        """(this rank's buffer, where the chunk exchanged with rank r lives) for the given buffer parity."""
        off = parity * self.local_n * FR_BYTES
        mine = self.rank * self.chunk * FR_BYTES
        return self._own + off, [self._peer[r] + off + mine for r in range(self.world)]

    def _all_to_all(self, out, inp) -> None:
        if self.world == 1:
            out.copy_(inp)
            return
        import torch.distributed as dist
        dist.all_to_all_single(out, inp, group=self.group)

    # ---- transforms ----------------------------------------------------------------------------------------
    def forward(self, coeffs, coset: bool = False):
        """coeffs: this rank's <= n/world coefficients (zero padded) -> this rank's block of evaluations."""
        coeffs = self._check_buf(coeffs, self.local_n, exact=False)
        out = self.empty()
        if self.mode == "p2p":
            parity = self._seq & 1
            self._seq += 1
            x, chunks = self._exchange_ptrs(parity)
            self.steps.forward_local(coeffs, coeffs.shape[0], coset, x)
            self._barrier()
            self.steps.forward_combine(chunks, out)
            return out
        x = self.empty()
        self.steps.forward_local(coeffs, coeffs.shape[0], coset, x)
        if self.world == 1:
            recv = x
        else:
            recv = self.empty()
            self._all_to_all(recv, x)
        self.steps.forward_combine(self._chunks_of(recv), out)
        return out

    def inverse(self, evals, coset: bool = False):
        """evals: this rank's block of n/world evaluations -> this rank's n/world coefficients."""
        evals = self._check_buf(evals, self.local_n, exact=True)
        if self.mode == "p2p":
            parity = self._seq & 1
            self._seq += 1
            x, chunks = self._exchange_ptrs(parity)
            self.steps.inverse_split(evals, chunks)
            self._barrier()
            out = self.empty()
            self.steps.inverse_local(x, coset, out)   # x is exchange memory: hand back a tensor the caller owns
            return out
        stage = self.empty()
        self.steps.inverse_split(evals, self._chunks_of(stage))
        if self.world == 1:
            x = stage
        else:
            x = self.empty()
            self._all_to_all(x, stage)
        self.steps.inverse_local(x, coset)
        return x

    def free(self) -> None:
        lib = _lib.load()
        if self.world > 1 and self._flag is not None:
            self._barrier()               # nobody still reads a buffer that is about to be unmapped
            self._flag = None
        for p in self._opened:
            lib.b2p_peer_close(p)
        self._opened = []
        if self._own:
            lib.b2p_peer_free(self._own)
            self._own = 0
        if hasattr(self.steps, "free"):
            self.steps.free()


def commit_lagrange(nt: ShardedNtt, srs, evals, coset: bool = False, group=None) -> bytes:
    """kzg.Commit of a polynomial given by its evaluations, both primitives sharded: the multi-GPU form of what
    the prover does for every wire column (Lagrange values -> iNTT -> canonical-basis MSM).

    evals : this rank's block of evaluations (ShardedNtt's evaluation distribution)
    srs   : sharded.ShardedSRS with layout="cyclic" over the same world, >= n points in total
    Returns the commitment (G1Affine memory layout), identical on every rank.  One exchange for the transform,
    one all_gather of a point per rank for the sum; the coefficients never leave the GPU that holds them."""
    import torch
    from . import sharded
    if srs.layout != "cyclic" or srs.world != nt.world or srs.rank != nt.rank or srs.curve != nt.curve:
        raise ValueError("the SRS must be this rank's cyclic shard over the same world and curve")
    if srs.count < nt.local_n:
        raise ValueError("SRS shard smaller than the local domain")
    stream = torch.cuda.ExternalStream(_lib.load().b2p_srs_stream(srs.handle), device=nt.device)
    stream.wait_stream(torch.cuda.current_stream())     # whatever produced `evals` comes first
    with torch.cuda.stream(stream):          # transform and MSM on the one stream the MSM launches on
        coeffs = nt.inverse(evals, coset=coset)
        local = srs.local_msm_dev_raw(coeffs.data_ptr(), nt.local_n)
        return sharded.g1_sum(nt.curve, sharded.all_gather_points(nt.curve, local, group, nt.device)) \
            if nt.world > 1 else local
