"""ctypes binding of the C++ CPU oracle (oracle/cpu_plonk.cpp) -- TEST INFRASTRUCTURE ONLY.

Same import rule as plonk_oracle.py: tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg only.  The C++ oracle is itself checked byte
for byte against the independent big-integer oracle (plonk_oracle.py) in
tests/test_oracle.py; it exists because Python integers cannot finish a 2^17 or
2^20 proof in test / benchmark time.

Scalars cross this boundary as 32-byte little-endian canonical integers, points as
x || y (FP_BYTES little-endian canonical each, (0,0) = infinity).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libplonk_oracle.so")
FP_BYTES = {0: 32, 1: 48}
CURVE_ID = {"BN254": 0, "BLS12_381": 1}

_lib = None


def build() -> None:
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        vp, u64, u32, i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        lib.ora_threads.restype = i
        lib.ora_set_threads.argtypes = [i]
        lib.ora_field_mul.argtypes = [i, vp, vp, vp]
        lib.ora_srs_from_tau.argtypes = [i, vp, u64, vp]
        lib.ora_msm.argtypes = [i, vp, vp, u64, vp]
        lib.ora_g1_decompress.restype = C.c_int64
        lib.ora_g1_decompress.argtypes = [i, vp, u64, vp]
        lib.ora_ntt.argtypes = [i, vp, u64, i]
        lib.ora_circuit_load.restype = vp
        lib.ora_circuit_load.argtypes = [i, u64, u32, vp, vp, vp, vp, vp, vp, u32, vp, vp, vp, u64]
        lib.ora_circuit_vk.argtypes = [vp, vp]
        lib.ora_prove.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        lib.ora_circuit_free.argtypes = [vp]
        _lib = lib
    return _lib


def threads() -> int:
    return load().ora_threads()


def set_threads(t: int) -> None:
    load().ora_set_threads(t)


def scalars_le(values: Sequence[int]) -> bytes:
    return b"".join(int(v).to_bytes(32, "little") for v in values)


def points_le(cid: int, points) -> bytes:
    nb = FP_BYTES[cid]
    out = bytearray()
    for P in points:
        out += bytes(2 * nb) if P is None else P[0].to_bytes(nb, "little") + P[1].to_bytes(nb, "little")
    return bytes(out)


def points_from_le(cid: int, data: bytes):
    nb = FP_BYTES[cid]
    out = []
    for i in range(0, len(data), 2 * nb):
        x = int.from_bytes(data[i:i + nb], "little")
        y = int.from_bytes(data[i + nb:i + 2 * nb], "little")
        out.append(None if x == 0 and y == 0 else (x, y))
    return out


def field_mul(field: int, a: int, b: int) -> int:
    """field: 0 Fr-BN254, 1 Fp-BN254, 2 Fr-BLS12-381, 3 Fp-BLS12-381."""
    nb = 48 if field == 3 else 32
    out = C.create_string_buffer(nb)
    assert load().ora_field_mul(field, a.to_bytes(nb, "little"), b.to_bytes(nb, "little"), out) == 0
    return int.from_bytes(out.raw, "little")


def srs_from_tau_bytes(cid: int, tau: int, n: int) -> bytes:
    out = C.create_string_buffer(n * 2 * FP_BYTES[cid])
    assert load().ora_srs_from_tau(cid, tau.to_bytes(32, "little"), n, out) == 0
    return out.raw


def srs_from_tau(cid: int, tau: int, n: int):
    return points_from_le(cid, srs_from_tau_bytes(cid, tau, n))


def g1_decompress_bytes(cid: int, compressed: bytes) -> bytes:
    """gnark compressed G1 stream (a pk.bin payload, setup/setup.go:196-228) -> x || y little-endian points."""
    nb = FP_BYTES[cid]
    n = len(compressed) // nb
    out = C.create_string_buffer(n * 2 * nb)
    rc = load().ora_g1_decompress(cid, compressed, n, out)
    if rc != 0:
        raise ValueError(f"compressed point {rc - 1} is invalid")
    return out.raw


def msm_bytes(cid: int, points: bytes, scalars: bytes):
    n = len(scalars) // 32
    out = C.create_string_buffer(2 * FP_BYTES[cid])
    assert load().ora_msm(cid, points, scalars, n, out) == 0
    return points_from_le(cid, out.raw)[0]


def msm(cid: int, points, scalars: Sequence[int]):
    return msm_bytes(cid, points_le(cid, points), scalars_le(scalars))


def ntt_bytes(cid: int, data: bytes, inverse: bool = False, coset: bool = False) -> bytes:
    buf = C.create_string_buffer(data, len(data))
    assert load().ora_ntt(cid, buf, len(data) // 32, (1 if inverse else 0) | (2 if coset else 0)) == 0
    return buf.raw


def ntt(cid: int, values: Sequence[int], inverse: bool = False, coset: bool = False) -> List[int]:
    raw = ntt_bytes(cid, scalars_le(values), inverse, coset)
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


class Circuit:
    """Trace + SRS loaded into the C++ oracle; prove() returns MarshalProof bytes."""

    def __init__(self, cid: int, n: int, nb_public: int, ql, qr, qm, qo, qk, perm, qcp=(), cidx=(),
                 srs_points_le: Optional[bytes] = None):
        lib = load()
        self.cid, self.n, self.k = cid, n, len(qcp)

        def col(v):
            return v if isinstance(v, (bytes, bytearray)) else scalars_le(v)

        cols = [col(c) for c in (ql, qr, qm, qo, qk)]
        permarr = perm if isinstance(perm, C.Array) else (C.c_int64 * (3 * n))(*perm)
        qbufs = [C.create_string_buffer(col(c), 32 * n) for c in qcp]
        qarr = (C.c_void_p * max(self.k, 1))(*[C.cast(b, C.c_void_p) for b in qbufs]) if self.k else None
        carr = (C.c_uint64 * max(self.k, 1))(*cidx) if self.k else None
        nsrs = len(srs_points_le) // (2 * FP_BYTES[cid])
        self.h = lib.ora_circuit_load(cid, n, nb_public, *cols, permarr, self.k, qarr, carr, srs_points_le, nsrs)
        if not self.h:
            raise ValueError("ora_circuit_load failed (bad sizes?)")

    def vk_points(self):
        out = C.create_string_buffer((8 + self.k) * 2 * FP_BYTES[self.cid])
        load().ora_circuit_vk(self.h, out)
        return points_from_le(self.cid, out.raw)

    def proof_size(self) -> int:
        return (24 + 3 * self.k) * 32 if self.cid == 0 else (33 + 4 * self.k) * 32

    def prove(self, L, R, O, blinding, pi2=(), bsb22_points=()) -> bytes:
        def col(v):
            return v if isinstance(v, (bytes, bytearray)) else scalars_le(v)

        pbufs = [C.create_string_buffer(col(c), 32 * self.n) for c in pi2]
        parr = (C.c_void_p * max(self.k, 1))(*[C.cast(b, C.c_void_p) for b in pbufs]) if self.k else None
        bsb = points_le(self.cid, bsb22_points) if self.k else None
        out = C.create_string_buffer(self.proof_size())
        rc = load().ora_prove(self.h, col(L), col(R), col(O), parr, bsb, col(blinding), out)
        if rc != 0:
            raise ArithmeticError(f"ora_prove failed ({rc}): constraints not satisfied?")
        return out.raw

    def free(self):
        if self.h:
            load().ora_circuit_free(self.h)
            self.h = None
