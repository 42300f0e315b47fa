"""Host-side mirror of AlgoPlonk's public API for the proving path, bound to the
CUDA library through the C ABI.

Reference surface mirrored (same names, argument meaning and error behaviour):
  Compile(circuit, curve, setup)            /root/reference/algoplonk.go:37-59
  (*CompiledCircuit).Verify(assignment)     /root/reference/algoplonk.go:79-98  (witness -> Prove -> plonk.Verify)
  plonk.Verify                              /root/reference/algoplonk.go:93     (verify / verify_batch / VerifyProof:
                                            b2p_verify, host arithmetic of the library, no GPU)
  srs.Vk.ReadFrom(vk.bin)                   /root/reference/setup/setup.go:174,190  (kzg_vk_load)
  MarshalProof / MarshalPublicInputs        /root/reference/helper.go:13-24,91-110
  setup names                                /root/reference/setup/setup.go:23-36
In the real integration these stay Go (go/gpuplonk); this module
is the harness tests and bench.py use to drive the same C entry points.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

from . import _lib
from . import frontend as fe

# ---- curve ids / setup registry (setup/setup.go:23-36) -----------------------
BN254, BLS12_381 = "BN254", "BLS12_381"
CURVE_ID = {BN254: _lib.B2P_BN254, BLS12_381: _lib.B2P_BLS12_381}
R_MOD = fe.R_MOD
P_MOD = {
    BN254: 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    BLS12_381: 4002409555221667393417789825735904156556882819939007885332058136124031650490837864442687629129015664037894272559787,
}
FP_BYTES = {BN254: 32, BLS12_381: 48}


class SetupName:
    TestOnlyBN254 = 0
    TestOnlyBLS12381 = 1
    PerpetualPowersOfTauBN254 = 2
    EthereumKzgCeremonyBLS12381 = 3
    DuskBLS12381 = 4


_SETUPS = {
    SetupName.TestOnlyBN254: (BN254, False),
    SetupName.TestOnlyBLS12381: (BLS12_381, False),
    SetupName.PerpetualPowersOfTauBN254: (BN254, True),
    SetupName.EthereumKzgCeremonyBLS12381: (BLS12_381, True),
    SetupName.DuskBLS12381: (BLS12_381, True),
}

# fixed tau of the TestOnly setups in this harness (unsafekzg draws a random one)
TEST_TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF


# ---- conversions between Python ints and gnark's in-memory layout -------------
def fr_to_mont_bytes(curve: str, values: Sequence[int]) -> bytes:
    r = R_MOD[curve]
    R = 1 << 256
    return b"".join(((v % r) * R % r).to_bytes(32, "little") for v in values)


def fr_from_mont_bytes(curve: str, data: bytes) -> List[int]:
    r = R_MOD[curve]
    Rinv = pow(1 << 256, -1, r)
    return [int.from_bytes(data[i:i + 32], "little") * Rinv % r for i in range(0, len(data), 32)]


def points_to_mont_bytes(curve: str, points) -> bytes:
    """affine (x, y) ints or None (infinity) -> G1Affine memory layout."""
    p, nb = P_MOD[curve], FP_BYTES[curve]
    R = 1 << (8 * nb)
    out = bytearray()
    for P in points:
        if P is None:
            out += bytes(2 * nb)
        else:
            out += (P[0] * R % p).to_bytes(nb, "little") + (P[1] * R % p).to_bytes(nb, "little")
    return bytes(out)


def points_from_mont_bytes(curve: str, data: bytes):
    p, nb = P_MOD[curve], FP_BYTES[curve]
    Rinv = pow(1 << (8 * nb), -1, p)
    out = []
    for i in range(0, len(data), 2 * nb):
        x = int.from_bytes(data[i:i + nb], "little") * Rinv % p
        y = int.from_bytes(data[i + nb:i + 2 * nb], "little") * Rinv % p
        out.append(None if x == 0 and y == 0 else (x, y))
    return out


def _buf(data: bytes):
    return C.create_string_buffer(data, len(data))


def g2_to_mont_bytes(curve: str, points) -> bytes:
    """affine ((x.A0, x.A1), (y.A0, y.A1)) ints or None -> G2Affine memory layout (X.A0 X.A1 Y.A0 Y.A1)."""
    p, nb = P_MOD[curve], FP_BYTES[curve]
    R = 1 << (8 * nb)
    out = bytearray()
    for Q in points:
        if Q is None:
            out += bytes(4 * nb)
        else:
            for c in (Q[0][0], Q[0][1], Q[1][0], Q[1][1]):
                out += (c * R % p).to_bytes(nb, "little")
    return bytes(out)


def g2_from_mont_bytes(curve: str, data: bytes):
    p, nb = P_MOD[curve], FP_BYTES[curve]
    Rinv = pow(1 << (8 * nb), -1, p)
    out = []
    for i in range(0, len(data), 4 * nb):
        c = [int.from_bytes(data[i + j * nb:i + (j + 1) * nb], "little") * Rinv % p for j in range(4)]
        out.append(None if not any(c) else ((c[0], c[1]), (c[2], c[3])))
    return out


# ---- verification: plonk.Verify and its pairing check, host arithmetic of the library (no GPU) --------------
def g2_unsafe(curve: str, tau: int = TEST_TAU) -> bytes:
    """[1]_2, [tau]_2 in G2Affine layout: the G2 half of unsafekzg.NewSRS (setup/setup.go:124)."""
    out = C.create_string_buffer(8 * FP_BYTES[curve])
    _lib.check(_lib.load().b2p_g2_generate_unsafe(CURVE_ID[curve], _buf(fr_to_mont_bytes(curve, [tau])), out))
    return out.raw


def kzg_vk_load(curve: str, vk_bin: bytes):
    """srs.Vk.ReadFrom on a setup's vk.bin (setup/setup.go:174,190): (Kzg.G2 as 2 G2Affine raw, Kzg.G1 raw)."""
    nb = FP_BYTES[curve]
    g2, g1 = C.create_string_buffer(8 * nb), C.create_string_buffer(2 * nb)
    _lib.check(_lib.load().b2p_kzg_vk_load(CURVE_ID[curve], _buf(bytes(vk_bin)), len(vk_bin), g2, g1))
    return g2.raw, g1.raw


def msm_g2_raw(curve: str, g2_points_raw: bytes, scalars: Sequence[int]) -> bytes:
    """G2Affine.MultiExp on the GPU (b2p_msm_g2): points in G2Affine memory layout, result likewise."""
    _lib.init()
    nb = 4 * FP_BYTES[curve]
    n = len(scalars)
    if len(g2_points_raw) != n * nb:
        raise ValueError("one G2Affine per scalar")
    out = C.create_string_buffer(nb)
    _lib.check(_lib.load().b2p_msm_g2(CURVE_ID[curve], _buf(g2_points_raw) if n else None,
                                      _buf(fr_to_mont_bytes(curve, scalars)) if n else None, n, out))
    return out.raw


def pairing_check(curve: str, g1_raw: bytes, g2_raw: bytes) -> bool:
    """prod e(P_i, Q_i) == 1 for points in G1Affine / G2Affine memory layout."""
    n = len(g1_raw) // (2 * FP_BYTES[curve])
    if len(g1_raw) != n * 2 * FP_BYTES[curve] or len(g2_raw) != n * 4 * FP_BYTES[curve]:
        raise ValueError("point buffers have the wrong length")
    ok = C.c_int(0)
    _lib.check(_lib.load().b2p_pairing_check(CURVE_ID[curve], _buf(g1_raw) if n else None, _buf(g2_raw) if n else None,
                                             n, C.byref(ok)))
    return bool(ok.value)


def verify(curve: str, n: int, nb_public: int, commitment_indexes: Sequence[int], vk_points_raw: bytes,
           kzg_g1_raw: bytes, kzg_g2_raw: bytes, proof: bytes, public_inputs: bytes) -> None:
    """plonk.Verify (algoplonk.go:93) on the marshalled proof / public inputs; raises ValueError("error verifying
    proof: ...") when the proof is rejected, as the reference returns an error."""
    k = len(commitment_indexes)
    cidx = (C.c_uint64 * max(k, 1))(*commitment_indexes) if k else None
    rc = _lib.load().b2p_verify(CURVE_ID[curve], n, nb_public, k, cidx, _buf(vk_points_raw), _buf(kzg_g1_raw),
                                _buf(kzg_g2_raw), _buf(proof), len(proof),
                                _buf(public_inputs) if public_inputs else None, len(public_inputs))
    if rc == _lib.ERR_VERIFY:
        raise ValueError(_lib.load().b2p_last_error().decode())
    _lib.check(rc)


def verify_batch(curve: str, n: int, nb_public: int, commitment_indexes: Sequence[int], vk_points_raw: bytes,
                 kzg_g1_raw: bytes, kzg_g2_raw: bytes, proofs: Sequence[bytes], public_inputs: Sequence[bytes],
                 device: bool = False) -> None:
    """Many proofs of one circuit, one pairing check (b2p_verify_batch; device=True: b2p_verify_batch_dev, the point
    combinations on the GPU).  Raises ValueError naming the first proof rejected before the pairing, or "batch" when
    only the folded pairing check failed."""
    if len(proofs) != len(public_inputs):
        raise ValueError("one public-input blob per proof")
    if len({len(p) for p in proofs}) > 1 or len({len(p) for p in public_inputs}) > 1:
        raise ValueError("proofs of one circuit have one length")
    k = len(commitment_indexes)
    cidx = (C.c_uint64 * max(k, 1))(*commitment_indexes) if k else None
    pl = len(proofs[0]) if proofs else 0
    ql = len(public_inputs[0]) if public_inputs else 0
    pj, qj = b"".join(proofs), b"".join(public_inputs)
    bad = C.c_uint64(0)
    if device:
        _lib.init()
    fn = _lib.load().b2p_verify_batch_dev if device else _lib.load().b2p_verify_batch
    rc = fn(CURVE_ID[curve], n, nb_public, k, cidx, _buf(vk_points_raw), _buf(kzg_g1_raw),
            _buf(kzg_g2_raw), _buf(pj) if pj else None, pl, _buf(qj) if qj else None, ql, len(proofs), C.byref(bad))
    if rc == _lib.ERR_VERIFY:
        raise ValueError(_lib.load().b2p_last_error().decode())
    _lib.check(rc)


# ---- SRS ------------------------------------------------------------------------
class SRS:
    """kzg.SRS resident on the GPU (canonical basis + windowed multiples)."""

    def __init__(self, curve: str, handle: int, tau: Optional[int] = None, g2: Optional[bytes] = None):
        self._g2 = g2            # Kzg.G2[0], Kzg.G2[1] in G2Affine layout (a trusted setup's vk.bin, decoded)
        self.curve, self.handle, self.tau = curve, handle, tau

    @classmethod
    def from_points(cls, curve: str, points, g2: Optional[bytes] = None) -> "SRS":
        """points: affine int pairs, or bytes already in G1Affine layout; g2: the setup's two G2 points
        (G2Affine layout), needed only to verify proofs made on this SRS."""
        _lib.init()
        data = points if isinstance(points, (bytes, bytearray)) else points_to_mont_bytes(curve, points)
        n = len(data) // (2 * FP_BYTES[curve])
        h = C.c_void_p()
        buf = _buf(bytes(data))
        _lib.check(_lib.load().b2p_srs_load(CURVE_ID[curve], buf, n, None, 0, C.byref(h)))
        return cls(curve, h.value, g2=g2)

    @classmethod
    def from_pk_bin(cls, curve: str, pk_bin: bytes, count: int, vk_bin: Optional[bytes] = None) -> "SRS":
        """setup/setup.go:165-228: the first `count` points of an embedded pk.bin (u32 BE count + compressed
        G1), decompressed on the GPU; vk_bin: the setup's vk.bin (its G2 points, for plonk.Verify)."""
        _lib.init()
        h = C.c_void_p()
        _lib.check(_lib.load().b2p_srs_load_compressed(CURVE_ID[curve], _buf(bytes(pk_bin)), len(pk_bin), count,
                                                       C.byref(h)))
        return cls(curve, h.value, g2=kzg_vk_load(curve, vk_bin)[0] if vk_bin is not None else None)

    @classmethod
    def unsafe(cls, curve: str, size: int, tau: int = TEST_TAU) -> "SRS":
        """unsafekzg.NewSRS (setup/setup.go:102-108)."""
        _lib.init()
        h = C.c_void_p()
        t = _buf(fr_to_mont_bytes(curve, [tau]))
        _lib.check(_lib.load().b2p_srs_generate_unsafe(CURVE_ID[curve], t, size, C.byref(h)))
        return cls(curve, h.value, tau % R_MOD[curve])

    @property
    def g2(self) -> Optional[bytes]:
        """vk.Kzg.G2 (two G2Affine): derived from tau for a TestOnly SRS, else what the caller supplied."""
        if self._g2 is None and self.tau is not None:
            self._g2 = g2_unsafe(self.curve, self.tau)
        return self._g2

    @property
    def size(self) -> int:
        return _lib.load().b2p_srs_size(self.handle)

    def msm_params(self):
        c, w, b = C.c_int(), C.c_int(), C.c_uint64()
        _lib.check(_lib.load().b2p_srs_msm_params(self.handle, C.byref(c), C.byref(w), C.byref(b)))
        return c.value, w.value, b.value

    def points(self, first: int, count: int):
        out = C.create_string_buffer(count * 2 * FP_BYTES[self.curve])
        _lib.check(_lib.load().b2p_srs_get_points(self.handle, first, count, out))
        return points_from_mont_bytes(self.curve, out.raw)

    def to_lagrange_raw(self, n: int) -> bytes:
        """kzg.ToLagrangeG1(srs.Pk.G1[:n]) (setup/setup.go:124,138): n G1Affine in memory layout."""
        out = C.create_string_buffer(n * 2 * FP_BYTES[self.curve])
        _lib.check(_lib.load().b2p_srs_to_lagrange(self.handle, n, out))
        return out.raw

    def to_lagrange(self, n: int):
        return points_from_mont_bytes(self.curve, self.to_lagrange_raw(n))

    def msm(self, scalars: Sequence[int], basis: int = _lib.BASIS_CANONICAL):
        """G1Affine.MultiExp / kzg.Commit: returns an affine int pair (None = infinity)."""
        data = _buf(fr_to_mont_bytes(self.curve, scalars)) if len(scalars) else None
        out = C.create_string_buffer(2 * FP_BYTES[self.curve])
        _lib.check(_lib.load().b2p_msm_g1(self.handle, basis, data, len(scalars), out))
        return points_from_mont_bytes(self.curve, out.raw)[0]

    def msm_raw(self, scalars_mont: bytes, basis: int = _lib.BASIS_CANONICAL) -> bytes:
        out = C.create_string_buffer(2 * FP_BYTES[self.curve])
        _lib.check(_lib.load().b2p_msm_g1(self.handle, basis, _buf(scalars_mont), len(scalars_mont) // 32, out))
        return out.raw

    def free(self):
        if self.handle:
            _lib.load().b2p_srs_free(self.handle)
            self.handle = None


def ntt(curve: str, values: Sequence[int], inverse: bool = False, coset: bool = False) -> List[int]:
    """fft.Domain.FFT / FFTInverse (natural order in and out)."""
    _lib.init()
    buf = _buf(fr_to_mont_bytes(curve, values))
    flags = (_lib.NTT_INVERSE if inverse else 0) | (_lib.NTT_COSET if coset else 0)
    _lib.check(_lib.load().b2p_ntt(CURVE_ID[curve], buf, len(values), flags))
    return fr_from_mont_bytes(curve, buf.raw)


# ---- proofs -----------------------------------------------------------------------
@dataclass
class Proof:
    """plonk.Proof: raw = 9 G1Affine + (7+k) Fr in gnark memory layout, plus the BSB22 commitments."""
    curve: str
    k: int
    raw: bytes
    bsb22: bytes = b""


@dataclass
class VerifiedProof:
    """algoplonk.go:28-31."""
    Proof: Proof
    Witness: List[int]          # public inputs (canonical ints)

    def ExportProofAndPublicInputs(self, proof_path: str, public_inputs_path: str) -> None:
        """algoplonk.go:103-131."""
        with open(proof_path, "wb") as f:
            f.write(MarshalProof(self.Proof))
        with open(public_inputs_path, "wb") as f:
            f.write(MarshalPublicInputs(self.Proof.curve, self.Witness))


def MarshalProof(proof: Proof) -> bytes:
    """helper.go:13-24."""
    lib = _lib.load()
    out = C.create_string_buffer(lib.b2p_proof_marshal_size(CURVE_ID[proof.curve], proof.k))
    _lib.check(lib.b2p_marshal_proof(CURVE_ID[proof.curve], proof.k, _buf(proof.raw),
                                     _buf(proof.bsb22) if proof.k else None, out))
    return out.raw


def MarshalPublicInputs(curve: str, public_values: Sequence[int]) -> bytes:
    """helper.go:91-110 (public witness minus its 12-byte header)."""
    lib = _lib.load()
    n = len(public_values)
    out = C.create_string_buffer(32 * n)
    vals = _buf(fr_to_mont_bytes(curve, public_values)) if n else None
    _lib.check(lib.b2p_marshal_public_inputs(CURVE_ID[curve], vals, n, out))
    return out.raw


# ---- compiled circuit ----------------------------------------------------------------
class CompiledCircuit:
    """algoplonk.go:21-26: Ccs + Pk + Vk + Curve, with Pk resident on the GPU."""

    def __init__(self, cs: fe.SparseR1CS, trace: fe.TraceColumns, srs: SRS, handle: int):
        self.Ccs, self.trace, self.srs, self.handle = cs, trace, srs, handle
        self.Curve = cs.curve
        self._vk_points = None
        self._vk_raw = self._g1_raw = None

    # verifying-key commitments S1 S2 S3 Ql Qr Qm Qo Qk Qcp* (affine ints)
    def vk_commitments(self):
        if self._vk_points is None:
            k = len(self.trace.qcp)
            out = C.create_string_buffer((8 + k) * 2 * FP_BYTES[self.Curve])
            _lib.check(_lib.load().b2p_circuit_vk_commitments(self.handle, out))
            self._vk_points = points_from_mont_bytes(self.Curve, out.raw)
        return self._vk_points

    def set_profiling(self, on: bool) -> None:
        _lib.check(_lib.load().b2p_circuit_set_profiling(self.handle, 1 if on else 0))

    def stats(self) -> dict:
        arr = (C.c_double * _lib.STAT_COUNT)()
        _lib.check(_lib.load().b2p_circuit_stats(self.handle, arr))
        return {name: arr[i] for i, name in enumerate(_lib.STAT_NAMES)}

    def prove_raw(self, L: bytes, R: bytes, O: bytes, blinding: bytes, pi2: Sequence[bytes] = (),
                  bsb22: bytes = b"") -> Proof:
        """plonk.Prove on buffers already in gnark layout (what the Go shim passes)."""
        lib = _lib.load()
        k = len(self.trace.qcp)
        cid = CURVE_ID[self.Curve]
        out = C.create_string_buffer(lib.b2p_proof_raw_size(cid, k))
        bufs = [_buf(p) for p in pi2]
        arr = (C.c_void_p * max(k, 1))(*[C.cast(b, C.c_void_p) for b in bufs]) if k else None
        _lib.check(lib.b2p_prove(self.handle, _buf(L) if not isinstance(L, C.Array) else L,
                                 _buf(R) if not isinstance(R, C.Array) else R,
                                 _buf(O) if not isinstance(O, C.Array) else O,
                                 arr, _buf(bsb22) if k else None, _buf(blinding), out))
        return Proof(self.Curve, k, out.raw, bsb22)

    def Prove(self, L: Sequence[int], R: Sequence[int], O: Sequence[int], blinding: Sequence[int],
              pi2: Sequence[Sequence[int]] = (), bsb22_points=()) -> Proof:
        cv = self.Curve
        return self.prove_raw(fr_to_mont_bytes(cv, L), fr_to_mont_bytes(cv, R), fr_to_mont_bytes(cv, O),
                              fr_to_mont_bytes(cv, blinding), [fr_to_mont_bytes(cv, v) for v in pi2],
                              points_to_mont_bytes(cv, bsb22_points))

    def Verify(self, L, R, O, blinding, pi2=(), bsb22_points=(), verifier: Optional[Callable] = None,
               verify: bool = True) -> VerifiedProof:
        """algoplonk.go:79-98: prove, then plonk.Verify (algoplonk.go:93) -- the library's host verifier
        b2p_verify against this circuit's verifying key.  The reference ALWAYS verifies and returns an error when
        that fails, so an SRS whose G2 points are unknown (a trusted setup loaded without its vk.bin / g2=) is an
        error here too, unless a `verifier` callback does the checking or the caller opts out with verify=False
        (then this is plonk.Prove only).  `verifier(proof_bytes, public_bytes)`, if given, runs in addition to
        b2p_verify when the G2 points are known (tests inject the oracle's restatement of the AVM verifier)."""
        if verify and verifier is None and self.srs.g2 is None:
            raise ValueError("error verifying proof: the SRS's G2 points are unknown (load the setup's vk.bin: "
                             "SRS.from_pk_bin(vk_bin=...) / SRS.from_points(g2=...)), or pass verify=False")
        proof = self.Prove(L, R, O, blinding, pi2, bsb22_points)
        public = [v % R_MOD[self.Curve] for v in L[: self.trace.nb_public]]
        if not verify:
            return VerifiedProof(proof, public)
        blob, pub = MarshalProof(proof), MarshalPublicInputs(self.Curve, public)
        if self.srs.g2 is not None:
            self.VerifyProof(blob, pub)
        if verifier is not None and not verifier(blob, pub):
            raise ValueError("error verifying proof")
        return VerifiedProof(proof, public)

    def VerifyProofs(self, proofs: Sequence[bytes], publics: Sequence[bytes], device: bool = False) -> None:
        """Many proofs of this circuit, one folded pairing check (b2p_verify_batch, or b2p_verify_batch_dev with
        device=True); raises ValueError when rejected."""
        self._key_material()
        verify_batch(self.Curve, self.trace.n, self.trace.nb_public, self.trace.commitment_constraint_indexes,
                     self._vk_raw, self._g1_raw, self.srs.g2, proofs, publics, device=device)

    def VerifyProof(self, proof_bytes: bytes, public_bytes: bytes) -> None:
        """plonk.Verify(proof, cc.Vk, publicWitness) on marshalled bytes; raises ValueError when rejected."""
        self._key_material()
        verify(self.Curve, self.trace.n, self.trace.nb_public, self.trace.commitment_constraint_indexes,
               self._vk_raw, self._g1_raw, self.srs.g2, proof_bytes, public_bytes)

    def _key_material(self) -> None:
        if self.srs.g2 is None:
            raise ValueError("the SRS's G2 points are unknown: pass g2= to SRS.from_points")
        if self._vk_raw is None:
            k = len(self.trace.qcp)
            out = C.create_string_buffer((8 + k) * 2 * FP_BYTES[self.Curve])
            _lib.check(_lib.load().b2p_circuit_vk_commitments(self.handle, out))
            self._vk_raw = out.raw
            self._g1_raw = points_to_mont_bytes(self.Curve, self.srs.points(0, 1))

    def free(self):
        if self.handle:
            _lib.load().b2p_circuit_free(self.handle)
            self.handle = None


def Compile(cs: fe.SparseR1CS, curve: str, setup_name: int, srs: Optional[SRS] = None,
            vk_transcript: Optional[bytes] = None) -> CompiledCircuit:
    """algoplonk.go:37-59.  `cs` is the already-built constraint system (gnark's frontend.Compile
    stays on the CPU); a trusted setup needs its SRS passed in (the embedded pk.bin files of the
    reference are loaded by the caller, setup/setup.go:196-228)."""
    if curve not in CURVE_ID:
        raise ValueError(f"unsupported curve: {curve}")
    if setup_name not in _SETUPS:
        raise ValueError(f"unknown setup: {setup_name}")
    s_curve, trusted = _SETUPS[setup_name]
    if s_curve != curve:
        raise ValueError("curve and trusted setup do not match")
    if cs.curve != curve:
        raise ValueError("constraint system was built for another curve")
    _lib.init()
    tc = fe.build_trace(cs)
    n = tc.n
    if srs is None:
        if trusted:
            raise ValueError("trusted setups need their SRS points (pk.bin) passed in")
        srs = SRS.unsafe(curve, n + 3)
    if srs.size < n + 3:
        raise ValueError(f"pk.bin too small for {n + 3} elements")      # setup/setup.go:219-223
    lib = _lib.load()
    cols = [_buf(fr_to_mont_bytes(curve, c)) for c in (tc.ql, tc.qr, tc.qm, tc.qo, tc.qk)]
    perm = (C.c_int64 * (3 * n))(*tc.perm)
    k = len(tc.qcp)
    qcp_bufs = [_buf(fr_to_mont_bytes(curve, c)) for c in tc.qcp]
    qcp_arr = (C.c_void_p * max(k, 1))(*[C.cast(b, C.c_void_p) for b in qcp_bufs]) if k else None
    cidx = (C.c_uint64 * max(k, 1))(*tc.commitment_constraint_indexes) if k else None
    h = C.c_void_p()
    # vk_transcript: the verifying key's commitments as gnark binds them into the transcript (S1 S2 S3 Ql Qr Qm Qo
    # Qk Qcp*, G1Affine.Marshal() each) -- what the Go shim passes from pk.Vk; None: the library commits itself
    vkb = _buf(vk_transcript) if vk_transcript is not None else None
    _lib.check(lib.b2p_circuit_load(srs.handle, n, tc.nb_public, *cols, perm, k, qcp_arr, cidx, vkb,
                                    len(vk_transcript) if vk_transcript is not None else 0, C.byref(h)))
    return CompiledCircuit(cs, tc, srs, h.value)


# ---- witness solver (spr.Solve inside plonk.Prove, algoplonk.go:81-89) ---------------------------------------
class Solver:
    """b2p_solver_*: assigns every internal variable from the public + secret inputs, level by level on the GPU
    (or on one host thread when the circuit is a dependency chain), and hands L, R, O to the prover."""
    INFO = ["levels", "widest_level", "solved_rows", "launches", "est_host_us", "est_device_us", "last_us", "last_where"]

    def __init__(self, cs: fe.SparseR1CS, trace: Optional[fe.TraceColumns] = None,
                 hint_fn: Optional[Callable[[int, List[int], int], Sequence[int]]] = None):
        """hint_fn(hint id, input values, n_out) -> n_out output values: the caller's hint functions (gnark:
        solver.WithHints); std_hint_fn below covers the front end's NBits.  Needed only when the constraint system
        records hints."""
        if cs.input_vars is None:
            raise ValueError("the constraint system does not say which variables are inputs")
        if cs.commitments and not cs.hints:
            raise ValueError("circuits with BSB22 commitments need the commitment hint recorded in the constraint system")
        if cs.hints and hint_fn is None:
            raise ValueError("the circuit uses solver hints: pass hint_fn")
        _lib.init()
        tc = trace if trace is not None else fe.build_trace(cs)
        self.curve, self.n, self.nb_public, self.nb_inputs = cs.curve, tc.n, cs.nb_public, len(cs.input_vars)
        n = tc.n
        cols = [_buf(fr_to_mont_bytes(cs.curve, c)) for c in (tc.ql, tc.qr, tc.qm, tc.qo, tc.qk)]
        xa, xb, xc = ((C.c_uint32 * n)(*w) for w in fe.solver_wires(cs, n))
        ids = (C.c_uint32 * max(self.nb_inputs, 1))(*cs.input_vars)
        # hints: b2p_hint records pointing into arrays this object keeps alive
        self._hint_arrays = [((C.c_uint32 * max(len(h.in_vars), 1))(*h.in_vars), (C.c_uint32 * len(h.out_vars))(*h.out_vars))
                             for h in cs.hints]
        hints = (_lib.Hint * max(len(cs.hints), 1))()
        for rec, h, (ia, oa) in zip(hints, cs.hints, self._hint_arrays):
            rec.id, rec.n_in, rec.n_out = h.id, len(h.in_vars), len(h.out_vars)
            rec.in_vars, rec.out_vars = C.cast(ia, C.POINTER(C.c_uint32)), C.cast(oa, C.POINTER(C.c_uint32))
        unchecked = None
        if cs.unchecked_rows:
            mask = bytearray(n)
            for j in cs.unchecked_rows:
                mask[cs.nb_public + j] = 1
            unchecked = _buf(bytes(mask))
        h = C.c_void_p()
        _lib.check(_lib.load().b2p_solver_create_hinted(CURVE_ID[cs.curve], n, cs.nb_public, cs.nb_variables, ids,
                                                        self.nb_inputs, *cols, xa, xb, xc,
                                                        hints if cs.hints else None, len(cs.hints), unchecked, C.byref(h)))
        self.handle = h.value
        self._hint_cb = None
        if cs.hints:
            curve = cs.curve

            def trampoline(_ctx, hid, inp, n_in, outp, n_out):
                try:
                    vals = fr_from_mont_bytes(curve, C.string_at(inp, 32 * n_in)) if n_in else []
                    got = list(hint_fn(hid, vals, n_out))
                    if len(got) != n_out:
                        return 1
                    C.memmove(outp, fr_to_mont_bytes(curve, got), 32 * n_out)
                    return 0
                except Exception:  # noqa: BLE001 -- reported by the library as a failed hint
                    return 2
            self._hint_cb = _lib.HINT_FN(trampoline)
            _lib.check(_lib.load().b2p_solver_set_hint_fn(self.handle, C.cast(self._hint_cb, C.c_void_p), None))

    def info(self) -> dict:
        out = (C.c_uint64 * 8)()
        _lib.check(_lib.load().b2p_solver_info(self.handle, out))
        return dict(zip(self.INFO, out))

    def solve_raw(self, inputs: Sequence[int], where: int = _lib.SOLVE_AUTO):
        """-> (L, R, O) as n*32 bytes each, Montgomery form: what b2p_prove takes."""
        if len(inputs) != self.nb_inputs:
            raise ValueError("one value per input variable")
        bufs = [C.create_string_buffer(32 * self.n) for _ in range(3)]
        _lib.check(_lib.load().b2p_solver_solve(self.handle, _buf(fr_to_mont_bytes(self.curve, inputs)), where, *bufs))
        return tuple(b.raw for b in bufs)

    def solve(self, inputs: Sequence[int], where: int = _lib.SOLVE_AUTO):
        return tuple(fr_from_mont_bytes(self.curve, b) for b in self.solve_raw(inputs, where))

    def solve_dev(self, inputs: Sequence[int], where: int = _lib.SOLVE_AUTO):
        """-> device pointers of L, R, O (owned by the solver, valid until its next solve) for b2p_prove_dev."""
        ptrs = [C.c_void_p() for _ in range(3)]
        _lib.check(_lib.load().b2p_solver_solve_dev(self.handle, _buf(fr_to_mont_bytes(self.curve, inputs)), where,
                                                    *[C.byref(p) for p in ptrs]))
        return tuple(p.value for p in ptrs)

    def free(self):
        if self.handle:
            _lib.load().b2p_solver_free(self.handle)
            self.handle = None


def std_hint_fn(extra: Optional[Callable[[int, List[int], int], Sequence[int]]] = None):
    """hint_fn for api.Solver covering the front end's standard hints (HINT_NBITS: the low n_out bits of the input,
    gnark's bits.NBits); other ids go to `extra` (e.g. the BSB22 commitment hint, which needs the SRS)."""
    def fn(hid: int, vals: List[int], n_out: int):
        if hid == fe.HINT_NBITS:
            return [(vals[0] >> i) & 1 for i in range(n_out)]
        if extra is None:
            raise ValueError(f"no function for hint {hid}")
        return extra(hid, vals, n_out)
    return fn


def VerifyFromInputs(cc: CompiledCircuit, solver: Solver, inputs: Sequence[int], blinding: Sequence[int],
                     where: int = _lib.SOLVE_AUTO, verify: bool = True) -> VerifiedProof:
    """(*CompiledCircuit).Verify as the reference runs it (algoplonk.go:79-98): the caller assigns the circuit's
    inputs only; solving, proving and plonk.Verify happen in the library, and L, R, O never leave the GPU."""
    lib = _lib.load()
    dL, dR, dO = solver.solve_dev(inputs, where)
    cid = CURVE_ID[cc.Curve]
    out = C.create_string_buffer(lib.b2p_proof_raw_size(cid, 0))
    _lib.check(lib.b2p_prove_dev(cc.handle, dL, dR, dO, None, None, _buf(fr_to_mont_bytes(cc.Curve, blinding)), out))
    proof = Proof(cc.Curve, 0, out.raw, b"")
    public = [v % R_MOD[cc.Curve] for v in inputs[: cc.trace.nb_public]]
    if verify:
        cc.VerifyProof(MarshalProof(proof), MarshalPublicInputs(cc.Curve, public))
    return VerifiedProof(proof, public)


# ---- persisted keys (utils/utils.go:66-157) ------------------------------------------------------------------
def ShouldRecompile(target_path: str, *source_paths: str) -> bool:
    """utils.ShouldRecompile (utils/utils.go:68-86): True when the target is missing or older than any source."""
    import os
    try:
        t = os.stat(target_path).st_mtime_ns
        return any(os.stat(sp).st_mtime_ns > t for sp in source_paths)
    except OSError:
        return True


def SerializeCompiledCircuit(cc: CompiledCircuit, filepath: str, vk_transcript: Optional[bytes] = None) -> None:
    """utils.SerializeCompiledCircuit (utils/utils.go:97-121) for a key resident on the GPU: the circuit half of the
    proving key (selector columns, permutation, BSB22 columns) as the library's own snapshot (b2p_circuit_save); the
    SRS half stays where it came from (pk.bin / a gnark key file / a known-tau setup)."""
    tc, curve = cc.trace, cc.Curve
    n, k = tc.n, len(tc.qcp)
    cols = [_buf(fr_to_mont_bytes(curve, c)) for c in (tc.ql, tc.qr, tc.qm, tc.qo, tc.qk)]
    perm = (C.c_int64 * (3 * n))(*tc.perm)
    qcp_bufs = [_buf(fr_to_mont_bytes(curve, c)) for c in tc.qcp]
    qcp_arr = (C.c_void_p * max(k, 1))(*[C.cast(b, C.c_void_p) for b in qcp_bufs]) if k else None
    cidx = (C.c_uint64 * max(k, 1))(*tc.commitment_constraint_indexes) if k else None
    vkb = _buf(vk_transcript) if vk_transcript else None
    _lib.check(_lib.load().b2p_circuit_save(filepath.encode(), CURVE_ID[curve], n, tc.nb_public, *cols, perm, k,
                                            qcp_arr, cidx, vkb, len(vk_transcript) if vk_transcript else 0))


def DeserializeCompiledCircuit(filepath: str, cs: fe.SparseR1CS, srs: SRS) -> CompiledCircuit:
    """utils.DeserializeCompiledCircuit (utils/utils.go:124-157): the proving key goes from the file's page cache
    straight to HBM (b2p_circuit_load_file); `cs` is the constraint system the solver keeps on the CPU."""
    _lib.init()
    h = C.c_void_p()
    _lib.check(_lib.load().b2p_circuit_load_file(srs.handle, filepath.encode(), C.byref(h)))
    return CompiledCircuit(cs, fe.build_trace(cs), srs, h.value)


def parse_gnark_file(data: bytes) -> _lib.GnarkFile:
    """The gob stream utils.SerializeCompiledCircuit writes -> curve and the byte ranges of Ccs / Pk / Vk."""
    out = _lib.GnarkFile()
    _lib.check(_lib.load().b2p_gnark_file_parse(_buf(data), len(data), C.byref(out)))
    return out


def parse_gnark_vk(curve: str, data: bytes) -> _lib.GnarkVk:
    """plonk.VerifyingKey.WriteTo bytes -> the fields b2p_verify / b2p_circuit_load take."""
    out = _lib.GnarkVk()
    _lib.check(_lib.load().b2p_gnark_vk_parse(CURVE_ID[curve], _buf(data), len(data), C.byref(out)))
    return out


def parse_gnark_pk(curve: str, data: bytes) -> _lib.GnarkPk:
    """plonk.ProvingKey.WriteTo bytes -> verifying key + where pk.Kzg / pk.KzgLagrange sit."""
    out = _lib.GnarkPk()
    _lib.check(_lib.load().b2p_gnark_pk_parse(CURVE_ID[curve], _buf(data), len(data), C.byref(out)))
    return out


def srs_from_gnark_pk(curve: str, pk_bytes: bytes, g2: Optional[bytes] = None) -> SRS:
    """pk.Kzg of a persisted gnark proving key, decompressed on the GPU straight from the key's bytes."""
    info = parse_gnark_pk(curve, pk_bytes)
    _lib.init()
    h = C.c_void_p()
    sect = pk_bytes[info.kzg_off:info.lagrange_off]
    _lib.check(_lib.load().b2p_srs_load_compressed(CURVE_ID[curve], _buf(sect), len(sect), info.kzg_count, C.byref(h)))
    return SRS(curve, h.value, g2=g2 if g2 is not None else bytes(info.vk.kzg_g2)[:8 * FP_BYTES[curve]])

