"""CPU oracle for the PLONK proving hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this module.  The product (algoplonk_b200/) never
does: it fails loudly when its CUDA library is missing.

What this file restates (plain Python big integers, slow on purpose):

  * the reference's *verification algorithm*, transcribed line by line from
    the verifier templates  /root/reference/verifier/templateLogicSigBN254.go:126-397
    and templateLogicSigBLS12_381.go:139-407  (``verify_proof`` below);
  * the proof byte layout of /root/reference/helper.go:13-88 (``marshal_proof``);
  * the SRS file format of /root/reference/setup/setup.go:196-228 and the
    compressed-point flags pinned by /root/reference/setup/trusted_setup_test.go
    (``load_srs_g1`` / ``g1_decompress``);
  * the PLONK *prover* (``prove``) that gnark v0.15.0 backend/plonk/{bn254,
    bls12-381}/prove.go + gnark-crypto v0.20.1 (kzg, fft, iop, fiat-shamir,
    hash_to_field) implement.  gnark is an un-vendored go.mod dependency
    (/root/reference/go.mod:8-9) and is absent from this machine, so the prover
    half restates the published algorithm and is anchored on the in-repo
    verifier: every proof produced here must be accepted by ``verify_proof``.

PARITY STATUS: "parity unpinned" at proof-value level (the reference holds no
golden proofs and gnark cannot be run here).  Pinned: SRS decoding against the
reference's known-answer points, proof layout/length, transcript order and the
accept/reject behaviour of the reference verifier algorithm.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# Curve parameters  (templateLogicSigBN254.go:15,18 / templateLogicSigBLS12_381.go:15,18)
# --------------------------------------------------------------------------


@dataclass(frozen=True)
class CurveParams:
    name: str
    cid: int            # 0 = BN254, 1 = BLS12-381 (matches include/b200plonk.h)
    p: int              # base field modulus
    r: int              # scalar field modulus
    b: int              # y^2 = x^3 + b
    g1: Tuple[int, int]
    fp_bytes: int       # 32 / 48
    two_adicity: int
    root: int           # primitive 2^two_adicity-th root of unity in Fr
    coset_shift: int    # FrMultiplicativeGen (VK_COSET_SHIFT)


BN254 = CurveParams(
    name="BN254", cid=0,
    p=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    b=3, g1=(1, 2), fp_bytes=32, two_adicity=28,
    root=19103219067921713944291392827692070036145651957329286315305642004821462161904,
    coset_shift=5,
)

BLS12_381 = CurveParams(
    name="BLS12_381", cid=1,
    p=4002409555221667393417789825735904156556882819939007885332058136124031650490837864442687629129015664037894272559787,
    r=52435875175126190479447740508185965837690552500527637822603658699938581184513,
    b=4,
    g1=(3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
        1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569),
    fp_bytes=48, two_adicity=32,
    root=10238227357739495823651030575849232062558860180284477541189508159991286009131,
    coset_shift=7,
)

CURVES = {"BN254": BN254, "BLS12_381": BLS12_381, 0: BN254, 1: BLS12_381}


def domain_generator(cv: CurveParams, n: int) -> int:
    """omega of order n (power of two): root^(2^(s-log2 n)) -- gnark-crypto fft.NewDomain."""
    assert n & (n - 1) == 0 and n >= 1
    lg = n.bit_length() - 1
    assert lg <= cv.two_adicity
    return pow(cv.root, 1 << (cv.two_adicity - lg), cv.r)


# --------------------------------------------------------------------------
# G1 arithmetic (Jacobian, a = 0).  None == point at infinity in affine form.
# --------------------------------------------------------------------------

Affine = Optional[Tuple[int, int]]
Jac = Tuple[int, int, int]
JAC_INF: Jac = (1, 1, 0)


def is_on_curve(cv: CurveParams, P: Affine) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - cv.b) % cv.p == 0


def to_jac(P: Affine) -> Jac:
    return JAC_INF if P is None else (P[0], P[1], 1)


def jac_double(cv: CurveParams, P: Jac) -> Jac:
    p = cv.p
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return JAC_INF
    A = X * X % p
    B = Y * Y % p
    C = B * B % p
    D = 2 * ((X + B) * (X + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def jac_add(cv: CurveParams, P: Jac, Q: Jac) -> Jac:
    p = cv.p
    if P[2] == 0:
        return Q
    if Q[2] == 0:
        return P
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return jac_double(cv, P)
        return JAC_INF
    H = (U2 - U1) % p
    R = (S2 - S1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = U1 * HH % p
    X3 = (R * R - HHH - 2 * V) % p
    Y3 = (R * (V - X3) - S1 * HHH) % p
    Z3 = Z1 * Z2 * H % p
    return (X3, Y3, Z3)


def jac_to_affine(cv: CurveParams, P: Jac) -> Affine:
    if P[2] == 0:
        return None
    p = cv.p
    zi = pow(P[2], -1, p)
    zi2 = zi * zi % p
    return (P[0] * zi2 % p, P[1] * zi2 * zi % p)


def g1_neg(cv: CurveParams, P: Affine) -> Affine:
    if P is None:
        return None
    return (P[0], (-P[1]) % cv.p)


def g1_add(cv: CurveParams, P: Affine, Q: Affine) -> Affine:
    return jac_to_affine(cv, jac_add(cv, to_jac(P), to_jac(Q)))


def jac_mul(cv: CurveParams, P: Jac, k: int) -> Jac:
    k %= cv.r
    acc = JAC_INF
    for bit in bin(k)[2:] if k else "":
        acc = jac_double(cv, acc)
        if bit == "1":
            acc = jac_add(cv, acc, P)
    return acc


def g1_mul(cv: CurveParams, P: Affine, k: int) -> Affine:
    return jac_to_affine(cv, jac_mul(cv, to_jac(P), k))


def msm_naive(cv: CurveParams, points: Sequence[Affine], scalars: Sequence[int]) -> Affine:
    """sum s_i * P_i by a small-window bucket method (what G1Affine.MultiExp computes)."""
    assert len(points) >= len(scalars)
    n = len(scalars)
    if n == 0:
        return None
    c = 4 if n < 64 else 8
    nwin = (cv.r.bit_length() + c - 1) // c
    acc = JAC_INF
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            acc = jac_double(cv, acc)
        buckets = [JAC_INF] * (1 << c)
        for s, P in zip(scalars, points):
            d = ((s % cv.r) >> (w * c)) & ((1 << c) - 1)
            if d and P is not None:
                buckets[d] = jac_add(cv, buckets[d], to_jac(P))
        run = JAC_INF
        tot = JAC_INF
        for d in range((1 << c) - 1, 0, -1):
            run = jac_add(cv, run, buckets[d])
            tot = jac_add(cv, tot, run)
        acc = jac_add(cv, acc, tot)
    return jac_to_affine(cv, acc)


# --------------------------------------------------------------------------
# Point encodings
# --------------------------------------------------------------------------

def g1_raw_bytes(cv: CurveParams, P: Affine, gnark_infinity_flag: bool = False) -> bytes:
    """gnark G1Affine.RawBytes(): X || Y big-endian canonical (helper.go:35).
    Infinity is all zero; for BLS12-381 gnark sets the 0x40 flag on byte 0
    (verifier/verifier.go:95-99) when gnark_infinity_flag is set."""
    if P is None:
        out = bytearray(2 * cv.fp_bytes)
        if gnark_infinity_flag and cv.cid == 1:
            out[0] = 0x40
        return bytes(out)
    return P[0].to_bytes(cv.fp_bytes, "big") + P[1].to_bytes(cv.fp_bytes, "big")


def g1_from_raw_bytes(cv: CurveParams, b: bytes) -> Affine:
    assert len(b) == 2 * cv.fp_bytes
    x = int.from_bytes(b[: cv.fp_bytes], "big")
    y = int.from_bytes(b[cv.fp_bytes:], "big")
    if x == 0 and y == 0:
        return None
    return (x, y)


def fp_sqrt(cv: CurveParams, a: int) -> Optional[int]:
    # both moduli are 3 mod 4
    s = pow(a, (cv.p + 1) // 4, cv.p)
    return s if s * s % cv.p == a % cv.p else None


def g1_decompress(cv: CurveParams, b: bytes) -> Affine:
    """Decode a gnark compressed G1 point (the format of setup/*/pk.bin).
    BN254: top 2 bits of byte 0: 10 = smallest y, 11 = largest y, 01 = infinity.
    BLS12-381: top 3 bits: 100 = smallest y, 101 = largest y, 110 = infinity
    (the 3-bit mask is the one trusted_setup_test.go:290-303 strips)."""
    assert len(b) == cv.fp_bytes
    if cv.cid == 0:
        flag = b[0] >> 6
        x = int.from_bytes(bytes([b[0] & 0x3F]) + b[1:], "big")
        if flag == 0b01:
            return None
        if flag not in (0b10, 0b11):
            raise ValueError("uncompressed/invalid flag in compressed stream")
        largest = flag == 0b11
    else:
        flag = b[0] >> 5
        x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
        if flag == 0b110:
            return None
        if flag not in (0b100, 0b101):
            raise ValueError("uncompressed/invalid flag in compressed stream")
        largest = flag == 0b101
    y = fp_sqrt(cv, (x * x * x + cv.b) % cv.p)
    if y is None:
        raise ValueError("x not on curve")
    lexicographically_largest = y > (cv.p - 1) // 2
    if lexicographically_largest != largest:
        y = cv.p - y
    return (x, y)


def g1_compress(cv: CurveParams, P: Affine) -> bytes:
    if P is None:
        out = bytearray(cv.fp_bytes)
        out[0] = 0x40 if cv.cid == 0 else 0xC0
        return bytes(out)
    x, y = P
    largest = y > (cv.p - 1) // 2
    out = bytearray(x.to_bytes(cv.fp_bytes, "big"))
    if cv.cid == 0:
        out[0] |= 0xC0 if largest else 0x80
    else:
        out[0] |= 0xA0 if largest else 0x80
    return bytes(out)


def load_srs_g1(cv: CurveParams, pk_bin: bytes, count: int) -> List[Affine]:
    """setup/setup.go:196-228: u32 BE count, then compressed G1 points."""
    declared = int.from_bytes(pk_bin[:4], "big")
    need = 4 + count * cv.fp_bytes
    if len(pk_bin) < need or declared < count:
        raise ValueError(f"pk.bin too small for {count} elements")
    return [g1_decompress(cv, pk_bin[4 + i * cv.fp_bytes: 4 + (i + 1) * cv.fp_bytes]) for i in range(count)]


def srs_from_tau(cv: CurveParams, tau: int, count: int) -> List[Affine]:
    """unsafekzg-style SRS: [tau^j] G1, j < count (setup/setup.go:102-108)."""
    out = []
    g = to_jac(cv.g1)
    t = 1
    for _ in range(count):
        out.append(jac_to_affine(cv, jac_mul(cv, g, t)))
        t = t * tau % cv.r
    return out


# --------------------------------------------------------------------------
# Fr helpers, NTT, polynomials (coefficient lists, low degree first)
# --------------------------------------------------------------------------

def fr_bytes(x: int) -> bytes:
    return x.to_bytes(32, "big")


def ntt(cv: CurveParams, a: Sequence[int], omega: int) -> List[int]:
    """natural-order in, natural-order out: out[i] = sum_j a[j] omega^(ij)."""
    r = cv.r
    n = len(a)
    assert n & (n - 1) == 0
    a = list(a)
    # bit reversal
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    length = 2
    while length <= n:
        w_len = pow(omega, n // length, r)
        half = length >> 1
        ws = [1] * half
        for k in range(1, half):
            ws[k] = ws[k - 1] * w_len % r
        for start in range(0, n, length):
            for k in range(half):
                u = a[start + k]
                v = a[start + k + half] * ws[k] % r
                a[start + k] = (u + v) % r
                a[start + k + half] = (u - v) % r
        length <<= 1
    return a


def intt(cv: CurveParams, a: Sequence[int], omega: int) -> List[int]:
    r = cv.r
    n = len(a)
    ninv = pow(n, -1, r)
    return [x * ninv % r for x in ntt(cv, a, pow(omega, -1, r))]


def coset_ntt(cv: CurveParams, coeffs: Sequence[int], omega: int, shift: int) -> List[int]:
    r = cv.r
    s = 1
    tmp = []
    for c in coeffs:
        tmp.append(c * s % r)
        s = s * shift % r
    return ntt(cv, tmp, omega)


def poly_eval(cv: CurveParams, coeffs: Sequence[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % cv.r
    return acc


def poly_mul(cv: CurveParams, a: Sequence[int], b: Sequence[int]) -> List[int]:
    if not a or not b:
        return []
    need = len(a) + len(b) - 1
    n = 1
    while n < need:
        n <<= 1
    if n <= 32:
        out = [0] * need
        for i, x in enumerate(a):
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % cv.r
        return out
    w = domain_generator(cv, n)
    fa = ntt(cv, list(a) + [0] * (n - len(a)), w)
    fb = ntt(cv, list(b) + [0] * (n - len(b)), w)
    return intt(cv, [x * y % cv.r for x, y in zip(fa, fb)], w)[:need]


def poly_add(cv: CurveParams, a: Sequence[int], b: Sequence[int]) -> List[int]:
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % cv.r for i in range(n)]


def poly_sub(cv: CurveParams, a: Sequence[int], b: Sequence[int]) -> List[int]:
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % cv.r for i in range(n)]


def poly_scale(cv: CurveParams, a: Sequence[int], k: int) -> List[int]:
    return [x * k % cv.r for x in a]


def poly_div_linear(cv: CurveParams, coeffs: Sequence[int], z: int) -> List[int]:
    """(p(X) - p(z)) / (X - z)  -- kzg dividePolyByXminusA."""
    n = len(coeffs)
    q = [0] * (n - 1)
    acc = 0
    for i in range(n - 1, 0, -1):
        acc = (coeffs[i] + acc * z) % cv.r
        q[i - 1] = acc
    return q


def poly_div_zh(cv: CurveParams, num: Sequence[int], n: int) -> List[int]:
    """exact division by X^n - 1; raises if there is a remainder."""
    q = [0] * len(num)
    for i in range(len(num)):
        q[i] = ((q[i - n] if i >= n else 0) - num[i]) % cv.r
    deg_q = len(num) - n
    if any(q[i] for i in range(deg_q, len(num))):
        raise ArithmeticError("numerator not divisible by X^n - 1 (constraints not satisfied?)")
    return q[:deg_q]


def blind(cv: CurveParams, coeffs: Sequence[int], b: Sequence[int], n: int) -> List[int]:
    """p(X) + b(X) (X^n - 1)  -- gnark getBlindedCoefficients."""
    out = list(coeffs) + [0] * (n + len(b) - len(coeffs))
    for i, bi in enumerate(b):
        out[i] = (out[i] - bi) % cv.r
        out[n + i] = (out[n + i] + bi) % cv.r
    return out


# --------------------------------------------------------------------------
# Fiat-Shamir / hash to field
# --------------------------------------------------------------------------

def hash_fr(cv: CurveParams, point_bytes: bytes) -> int:
    """templateLogicSigBN254.go:386-397: expand_msg_xmd(SHA-256, DST 'BSB22-Plonk', 48 B) mod r."""
    dst_prime = b"BSB22-Plonk\x0b"
    b0 = hashlib.sha256(bytes(64) + point_bytes + b"\x00\x30\x00" + dst_prime).digest()
    b1 = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    b2 = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, b1)) + b"\x02" + dst_prime).digest()
    res = (int.from_bytes(b1, "big") * (1 << 128)) % cv.r
    return (res + int.from_bytes(b2[:16], "big")) % cv.r


def fs_point(cv: CurveParams, P: Affine) -> bytes:
    """Point bytes as the PROVER binds them into the transcript: gnark's
    G1Affine.Marshal() == RawBytes(); on BLS12-381 the point at infinity carries
    gnark's 0x40 flag (verifier/verifier.go:95-104).  The reference template's
    fs() (BLS :401-407) hashes 0x80 instead for *proof* points -- an unreachable
    case for blinded commitments; the prover follows gnark (SURVEY 8f footnote)."""
    return g1_raw_bytes(cv, P, gnark_infinity_flag=True)


# --------------------------------------------------------------------------
# Circuit trace / keys / proof containers
# --------------------------------------------------------------------------

@dataclass
class Trace:
    """The part of gnark's plonk.ProvingKey trace the prover needs."""
    curve: CurveParams
    n: int
    nb_public: int
    ql: List[int]
    qr: List[int]
    qm: List[int]
    qo: List[int]
    qk: List[int]                   # WITHOUT public inputs / commitment hashes
    perm: List[int]                 # 3n positions (gnark trace.S)
    qcp: List[List[int]] = field(default_factory=list)
    commitment_constraint_indexes: List[int] = field(default_factory=list)


@dataclass
class VerifyingKey:
    curve: CurveParams
    size: int
    size_inv: int
    omega: int
    nb_public: int
    coset_shift: int
    S: List[Affine]
    Ql: Affine
    Qr: Affine
    Qm: Affine
    Qo: Affine
    Qk: Affine
    Qcp: List[Affine]
    commitment_constraint_indexes: List[int]
    g1: Affine                       # Kzg.G1 (SRS[0])
    tau: Optional[int] = None        # known-tau (TestOnly) SRS: pairing-free check
    g2: Optional[tuple] = None       # (G2[0], G2[1]) for a real pairing


@dataclass
class Proof:
    LRO: List[Affine]
    Z: Affine
    H: List[Affine]
    bsb22: List[Affine]
    batched_H: Affine
    claimed: List[int]               # [lin(zeta), l, r, o, s1, s2, qcp_0..]
    zshift_H: Affine
    zshift_claimed: int


def identity_support(cv: CurveParams, n: int) -> List[int]:
    """iop getSupportIdentityPermutation: [w^i, u w^i, u^2 w^i]."""
    w = domain_generator(cv, n)
    out = [1] * (3 * n)
    for i in range(1, n):
        out[i] = out[i - 1] * w % cv.r
    for i in range(n):
        out[n + i] = out[i] * cv.coset_shift % cv.r
        out[2 * n + i] = out[n + i] * cv.coset_shift % cv.r
    return out


def permutation_polys(tr: Trace) -> Tuple[List[int], List[int], List[int]]:
    """S1,S2,S3 in Lagrange form: id[perm[j*n+i]] (gnark computePermutationPolynomials)."""
    idv = identity_support(tr.curve, tr.n)
    n = tr.n
    s = [idv[tr.perm[i]] for i in range(3 * n)]
    return s[:n], s[n:2 * n], s[2 * n:]


def commit(cv: CurveParams, srs: Sequence[Affine], coeffs: Sequence[int]) -> Affine:
    """kzg.Commit: MSM(canonical SRS, coefficients)."""
    return msm_naive(cv, srs[: len(coeffs)], coeffs)


def setup(tr: Trace, srs: Sequence[Affine], tau: Optional[int] = None, g2=None) -> VerifyingKey:
    """plonk.Setup (setup/setup.go:149): commitments to selectors and permutation polynomials."""
    cv = tr.curve
    n = tr.n
    w = domain_generator(cv, n)
    assert len(srs) >= n + 3
    s1, s2, s3 = permutation_polys(tr)

    def C(lag):
        return commit(cv, srs, intt(cv, lag, w))

    return VerifyingKey(
        curve=cv, size=n, size_inv=pow(n, -1, cv.r), omega=w, nb_public=tr.nb_public,
        coset_shift=cv.coset_shift, S=[C(s1), C(s2), C(s3)],
        Ql=C(tr.ql), Qr=C(tr.qr), Qm=C(tr.qm), Qo=C(tr.qo), Qk=C(tr.qk),
        Qcp=[C(q) for q in tr.qcp],
        commitment_constraint_indexes=list(tr.commitment_constraint_indexes),
        g1=srs[0], tau=tau, g2=g2,
    )


def vk_transcript_bytes(vk: VerifyingKey) -> bytes:
    """Bytes bound before the public inputs in the gamma challenge
    (templateLogicSigBN254.go:131-132): S1 S2 S3 Ql Qr Qm Qo Qk Qcp*."""
    cv = vk.curve
    pts = vk.S + [vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk] + vk.Qcp
    return b"".join(g1_raw_bytes(cv, P, gnark_infinity_flag=True) for P in pts)


def bsb22_commit(tr: Trace, srs: Sequence[Affine], pi2_lagrange: Sequence[int]) -> Affine:
    """The BSB22 hint's commitment: MSM(Lagrange SRS, committed values) == commit(iNTT(values))."""
    cv = tr.curve
    return commit(cv, srs, intt(cv, pi2_lagrange, domain_generator(cv, tr.n)))


# --------------------------------------------------------------------------
# Prover (gnark v0.15.0 prove.go restated)
# --------------------------------------------------------------------------

def grand_product(tr: Trace, L, R, O, beta: int, gamma: int) -> List[int]:
    """iop.BuildRatioCopyConstraint: Z in Lagrange form."""
    cv = tr.curve
    r = cv.r
    n = tr.n
    idv = identity_support(cv, n)
    Z = [1] * n
    for i in range(n - 1):
        num = ((L[i] + beta * idv[i] + gamma) * (R[i] + beta * idv[n + i] + gamma) % r
               * (O[i] + beta * idv[2 * n + i] + gamma)) % r
        den = ((L[i] + beta * idv[tr.perm[i]] + gamma) * (R[i] + beta * idv[tr.perm[n + i]] + gamma) % r
               * (O[i] + beta * idv[tr.perm[2 * n + i]] + gamma)) % r
        Z[i + 1] = Z[i] * num % r * pow(den, -1, r) % r
    return Z


def prove(tr: Trace, vk: VerifyingKey, srs: Sequence[Affine],
          L: Sequence[int], R: Sequence[int], O: Sequence[int],
          blinding: Sequence[int],
          pi2: Sequence[Sequence[int]] = (), bsb22_commitments: Sequence[Affine] = (),
          return_debug: bool = False):
    """PLONK prover.  blinding = [bl0,bl1, br0,br1, bo0,bo1, bz0,bz1,bz2]
    (b_L = bl0 + bl1 X, ..., b_Z = bz0 + bz1 X + bz2 X^2)."""
    cv = tr.curve
    r = cv.r
    n = tr.n
    w = vk.omega
    u = cv.coset_shift
    k = len(tr.qcp)
    assert len(pi2) == k and len(bsb22_commitments) == k
    assert len(blinding) == 9
    bl, br, bo, bz = blinding[0:2], blinding[2:4], blinding[4:6], blinding[6:9]
    public_inputs = [x % r for x in L[: tr.nb_public]]

    # completeQk: public inputs and commitment hashes written into qk
    qk = list(tr.qk)
    for i, v in enumerate(public_inputs):
        qk[i] = v
    for c in range(k):
        qk[tr.nb_public + tr.commitment_constraint_indexes[c]] = hash_fr(cv, fs_point(cv, bsb22_commitments[c]))

    # round 1: wire polynomials, blinded, committed
    lc = blind(cv, intt(cv, L, w), bl, n)
    rc = blind(cv, intt(cv, R, w), br, n)
    oc = blind(cv, intt(cv, O, w), bo, n)
    com_l, com_r, com_o = commit(cv, srs, lc), commit(cv, srs, rc), commit(cv, srs, oc)

    # round 2: gamma, beta, Z
    pub_bytes = b"".join(fr_bytes(v) for v in public_inputs)
    gamma_pre = hashlib.sha256(b"gamma" + vk_transcript_bytes(vk) + pub_bytes
                               + fs_point(cv, com_l) + fs_point(cv, com_r) + fs_point(cv, com_o)).digest()
    beta_pre = hashlib.sha256(b"beta" + gamma_pre).digest()
    gamma = int.from_bytes(gamma_pre, "big") % r
    beta = int.from_bytes(beta_pre, "big") % r
    Zl = grand_product(tr, L, R, O, beta, gamma)
    zc = blind(cv, intt(cv, Zl, w), bz, n)
    com_z = commit(cv, srs, zc)

    # round 3: alpha, quotient
    alpha_pre = hashlib.sha256(b"alpha" + beta_pre + b"".join(fs_point(cv, P) for P in bsb22_commitments)
                               + fs_point(cv, com_z)).digest()
    alpha = int.from_bytes(alpha_pre, "big") % r

    s1l, s2l, s3l = permutation_polys(tr)
    s1c, s2c, s3c = intt(cv, s1l, w), intt(cv, s2l, w), intt(cv, s3l, w)
    qlc, qrc, qmc, qoc = (intt(cv, q, w) for q in (tr.ql, tr.qr, tr.qm, tr.qo))
    qkc = intt(cv, qk, w)
    qcpc = [intt(cv, q, w) for q in tr.qcp]
    pi2c = [intt(cv, v, w) for v in pi2]

    M = lambda a, b: poly_mul(cv, a, b)
    A = lambda a, b: poly_add(cv, a, b)
    gate = A(A(A(M(qlc, lc), M(qrc, rc)), A(M(qmc, M(lc, rc)), M(qoc, oc))), qkc)
    for c in range(k):
        gate = A(gate, M(qcpc[c], pi2c[c]))
    z_shift = [zc[i] * pow(w, i, r) % r for i in range(len(zc))]          # z(wX)
    g = [gamma]
    X1 = [0, beta]
    Xu = [0, beta * u % r]
    Xu2 = [0, beta * u * u % r]
    perm_a = M(M(A(A(lc, poly_scale(cv, s1c, beta)), g), A(A(rc, poly_scale(cv, s2c, beta)), g)),
               M(A(A(oc, poly_scale(cv, s3c, beta)), g), z_shift))
    perm_b = M(M(A(A(lc, X1), g), A(A(rc, Xu), g)), M(A(A(oc, Xu2), g), zc))
    perm = poly_sub(cv, perm_a, perm_b)
    ninv = pow(n, -1, r)
    l1c = [ninv] * n                                                      # L_1 = (X^n-1)/(n(X-1))
    loc = M(l1c, poly_sub(cv, zc, [1]))
    num = A(A(gate, poly_scale(cv, perm, alpha)), poly_scale(cv, loc, alpha * alpha % r))
    h = poly_div_zh(cv, num, n)
    h = h + [0] * (3 * (n + 2) - len(h))
    assert len(h) == 3 * (n + 2), "deg h must be <= 3n+5"
    h0, h1, h2 = h[: n + 2], h[n + 2: 2 * (n + 2)], h[2 * (n + 2):]
    com_h = [commit(cv, srs, h0), commit(cv, srs, h1), commit(cv, srs, h2)]

    # round 4: zeta, evaluations
    zeta_pre = hashlib.sha256(b"zeta" + alpha_pre + b"".join(fs_point(cv, P) for P in com_h)).digest()
    zeta = int.from_bytes(zeta_pre, "big") % r
    zw = zeta * w % r
    z_at_zw = poly_eval(cv, zc, zw)
    com_w_zw = commit(cv, srs, poly_div_linear(cv, zc, zw))
    l_z, r_z, o_z = poly_eval(cv, lc, zeta), poly_eval(cv, rc, zeta), poly_eval(cv, oc, zeta)
    s1_z, s2_z = poly_eval(cv, s1c, zeta), poly_eval(cv, s2c, zeta)
    qcp_z = [poly_eval(cv, q, zeta) for q in qcpc]

    # round 5: linearised polynomial (templateLogicSigBN254.go:220-278 inverted)
    zn = pow(zeta, n, r)
    zh_z = (zn - 1) % r
    l1_z = zh_z * ninv % r * pow((zeta - 1) % r, -1, r) % r
    a2l = alpha * alpha % r * l1_z % r
    s1p = alpha * beta % r * z_at_zw % r * ((l_z + beta * s1_z + gamma) % r) % r * ((r_z + beta * s2_z + gamma) % r) % r
    s2p = (-alpha * ((l_z + beta * zeta + gamma) % r) % r * ((r_z + beta * u * zeta + gamma) % r) % r
           * ((o_z + beta * u * u * zeta + gamma) % r) + a2l) % r
    qk_vk_c = intt(cv, tr.qk, w)
    lin = poly_scale(cv, qlc, l_z)
    lin = A(lin, poly_scale(cv, qrc, r_z))
    lin = A(lin, poly_scale(cv, qmc, l_z * r_z % r))
    lin = A(lin, poly_scale(cv, qoc, o_z))
    lin = A(lin, qk_vk_c)
    for c in range(k):
        lin = A(lin, poly_scale(cv, pi2c[c], qcp_z[c]))
    lin = A(lin, poly_scale(cv, s3c, s1p))
    lin = A(lin, poly_scale(cv, zc, s2p))
    zn2 = pow(zeta, n + 2, r)
    folded_h = A(A(h0, poly_scale(cv, h1, zn2)), poly_scale(cv, h2, zn2 * zn2 % r))
    lin = poly_sub(cv, lin, poly_scale(cv, folded_h, zh_z))
    lin_z = poly_eval(cv, lin, zeta)
    com_lin = commit(cv, srs, lin)

    # batch opening at zeta (kzg.BatchOpenSinglePoint); fold challenge :280-286
    polys = [lin, lc, rc, oc, s1c, s2c] + qcpc
    digests = [com_lin, com_l, com_r, com_o, vk.S[0], vk.S[1]] + list(vk.Qcp)
    claimed = [lin_z, l_z, r_z, o_z, s1_z, s2_z] + qcp_z
    v_pre = hashlib.sha256(
        b"gamma" + fr_bytes(zeta)
        + b"".join(fs_point(cv, P) for P in digests)
        + b"".join(fr_bytes(c) for c in claimed) + fr_bytes(z_at_zw)).digest()
    v = int.from_bytes(v_pre, "big") % r
    folded = []
    acc = 1
    for pcoef in polys:
        folded = A(folded, poly_scale(cv, pcoef, acc))
        acc = acc * v % r
    com_w_z = commit(cv, srs, poly_div_linear(cv, folded, zeta))

    proof = Proof(LRO=[com_l, com_r, com_o], Z=com_z, H=com_h, bsb22=list(bsb22_commitments),
                  batched_H=com_w_z, claimed=claimed, zshift_H=com_w_zw, zshift_claimed=z_at_zw)
    if return_debug:
        dbg = dict(gamma=gamma, beta=beta, alpha=alpha, zeta=zeta, v=v, Z=Zl, lc=lc, rc=rc, oc=oc, zc=zc,
                   h=h, lin=lin, com_lin=com_lin, qk_completed=qk, folded=folded)
        return proof, dbg
    return proof


# --------------------------------------------------------------------------
# Marshalling (helper.go:13-110, Appendix B)
# --------------------------------------------------------------------------

def marshal_proof(cv: CurveParams, pf: Proof) -> bytes:
    """BN254: gnark MarshalSolidity layout (helper.go:16-17: raw X || Y, infinity all zero); BLS12-381:
    helper.go:27-88, points through RawBytes(), whose infinity is 0x40 followed by zeros."""
    out = b""
    for P in pf.LRO:
        out += g1_raw_bytes(cv, P, True)
    for P in pf.H:
        out += g1_raw_bytes(cv, P, True)
    for i in range(1, 6):
        out += fr_bytes(pf.claimed[i])
    out += g1_raw_bytes(cv, pf.Z, True)
    out += fr_bytes(pf.zshift_claimed)
    out += g1_raw_bytes(cv, pf.batched_H, True)
    out += g1_raw_bytes(cv, pf.zshift_H, True)
    for i in range(len(pf.bsb22)):
        out += fr_bytes(pf.claimed[6 + i])
    for P in pf.bsb22:
        out += g1_raw_bytes(cv, P, True)
    return out


def marshal_public_inputs(values: Sequence[int]) -> bytes:
    """helper.go:91-110: nbPublic x 32 B big-endian."""
    return b"".join(fr_bytes(v) for v in values)


def proof_size(cv: CurveParams, k: int) -> int:
    return (24 + 3 * k) * 32 if cv.cid == 0 else (33 + 4 * k) * 32


# --------------------------------------------------------------------------
# Verifier: transcription of templateLogicSigBN254.go:32-356 (+ BLS12-381 twin)
# --------------------------------------------------------------------------

class _EC:
    """AVM ec_add / ec_scalar_mul on raw X||Y bytes (zero bytes == infinity)."""

    def __init__(self, cv: CurveParams):
        self.cv = cv

    def dec(self, b: bytes) -> Affine:
        P = g1_from_raw_bytes(self.cv, b)
        if not is_on_curve(self.cv, P):
            raise ValueError("point not on curve")
        return P

    def enc(self, P: Affine) -> bytes:
        return g1_raw_bytes(self.cv, P)

    def scalar_mul(self, pb: bytes, sb: bytes) -> bytes:
        # the AVM opcode takes the scalar as a big-endian integer; not reduced here on purpose
        k = int.from_bytes(sb, "big")
        return self.enc(g1_mul(self.cv, self.dec(pb), k % self.cv.r))

    def add(self, a: bytes, b: bytes) -> bytes:
        return self.enc(g1_add(self.cv, self.dec(a), self.dec(b)))


def verify_proof(vk: VerifyingKey, proof: bytes, public_inputs: bytes) -> bool:
    """Returns True iff the reference's generated verifier would accept.  A point the AVM's ec opcodes cannot decode
    (not on the curve, or the 0x40 infinity flag RawBytes() puts on a BLS12-381 point: the AVM takes raw X || Y only)
    fails the program, i.e. rejects."""
    try:
        return _verify_proof(vk, proof, public_inputs)
    except ValueError as e:
        if "point not on curve" in str(e):
            return False
        raise


def _verify_proof(vk: VerifyingKey, proof: bytes, public_inputs: bytes) -> bool:
    """The restatement proper.
    Line references are to verifier/templateLogicSigBN254.go; the BLS12-381
    template differs only in offsets, 48-byte coordinates and fs()."""
    cv = vk.curve
    q = cv.r
    ec = _EC(cv)
    PB = 2 * cv.fp_bytes
    k = len(vk.commitment_constraint_indexes)
    bls = cv.cid == 1

    def fs(pb: bytes) -> bytes:                          # BLS :401-407
        if bls and pb == bytes(96):
            return bytes([0x80]) + bytes(95)
        return pb

    def invert(pb: bytes) -> bytes:                      # :376-384
        x = pb[: cv.fp_bytes]
        y = int.from_bytes(pb[cv.fp_bytes:], "big")
        if y == 0:
            return pb
        return x + (cv.p - y).to_bytes(cv.fp_bytes, "big")

    def curvemod(b: bytes) -> int:                       # :371-374
        return int.from_bytes(b, "big") % q

    sha256 = lambda b: hashlib.sha256(b).digest()
    I = lambda b: int.from_bytes(b, "big")
    U256 = lambda x: x.to_bytes(32, "big")

    # :50-51 length checks
    if len(proof) != proof_size(cv, k):
        raise AssertionError("proof length")
    if len(public_inputs) != vk.nb_public * 32:
        raise AssertionError("public inputs length")

    VK_QL, VK_QR, VK_QO, VK_QM, VK_QK = (ec.enc(P) for P in (vk.Ql, vk.Qr, vk.Qo, vk.Qm, vk.Qk))
    VK_S1, VK_S2, VK_S3 = (ec.enc(P) for P in vk.S)
    VK_QCP = [ec.enc(P) for P in vk.Qcp]
    enc_fs = lambda P: g1_raw_bytes(cv, P, gnark_infinity_flag=True)   # hexEncoded (verifier.go:101-104)
    VK_QL_fs, VK_QR_fs, VK_QO_fs, VK_QM_fs, VK_QK_fs = (enc_fs(P) for P in (vk.Ql, vk.Qr, vk.Qo, vk.Qm, vk.Qk))
    VK_S1_fs, VK_S2_fs, VK_S3_fs = (enc_fs(P) for P in vk.S)
    VK_QCP_fs = [enc_fs(P) for P in vk.Qcp]

    # :75-107 proof slices
    off = 0

    def take(nbytes):
        nonlocal off
        s = proof[off: off + nbytes]
        off += nbytes
        return s

    L_COM, R_COM, O_COM = take(PB), take(PB), take(PB)
    H_0, H_1, H_2 = take(PB), take(PB), take(PB)
    L_AT_Z, R_AT_Z, O_AT_Z = take(32), take(32), take(32)
    S1_AT_Z, S2_AT_Z = take(32), take(32)
    GRAND_PRODUCT = take(PB)
    GRAND_PRODUCT_AT_Z_OMEGA = take(32)
    BATCH_OPENING_AT_Z = take(PB)
    OPENING_AT_Z_OMEGA = take(PB)
    QCP_AT_Z = [take(32) for _ in range(k)]
    BSB_COM = [take(PB) for _ in range(k)]
    assert off == len(proof)

    # :109-124 range checks
    for b in [L_AT_Z, R_AT_Z, O_AT_Z, S1_AT_Z, S2_AT_Z, GRAND_PRODUCT_AT_Z_OMEGA] + QCP_AT_Z:
        if I(b) >= q:
            return False
    for i in range(vk.nb_public):
        if I(public_inputs[i * 32:(i + 1) * 32]) >= q:
            return False

    # :131-140 challenges
    gamma_pre = sha256(b"gamma" + VK_S1_fs + VK_S2_fs + VK_S3_fs + VK_QL_fs + VK_QR_fs + VK_QM_fs + VK_QO_fs
                       + VK_QK_fs + b"".join(VK_QCP_fs) + public_inputs + fs(L_COM) + fs(R_COM) + fs(O_COM))
    beta_pre = sha256(b"beta" + gamma_pre)
    alpha_pre = sha256(b"alpha" + beta_pre + b"".join(fs(c) for c in BSB_COM) + fs(GRAND_PRODUCT))
    zeta_pre = sha256(b"zeta" + alpha_pre + fs(H_0) + fs(H_1) + fs(H_2))
    gamma, beta, alpha, zeta = curvemod(gamma_pre), curvemod(beta_pre), curvemod(alpha_pre), curvemod(zeta_pre)

    # :142-146
    Zz = (pow(zeta, vk.size, q) + q - 1) % q
    zn = Zz * vk.size_inv % q

    # :148-186 public input interpolation
    w_ = 1
    batch = []
    for i in range(vk.nb_public):
        batch.append((zeta + q - w_) % q)
        w_ = w_ * vk.omega % q
    temp = [1]
    prev = 1
    for x in batch:
        y = x * prev % q
        temp.append(y)
        prev = y
    inv = pow(prev, q - 2, q)
    i = vk.nb_public
    while i > 0:
        tmp = batch[i - 1]
        cur = inv * temp[i - 1] % q
        batch[i - 1] = cur
        inv = inv * tmp % q
        i -= 1
    w_ = 1
    for i in range(vk.nb_public):
        batch[i] = w_ * (batch[i] * zn % q) % q
        w_ = w_ * vk.omega % q
    PI = 0
    for i in range(vk.nb_public):
        PI = (PI + batch[i] * I(public_inputs[i * 32:(i + 1) * 32])) % q
    # :187-194 BSB22 contributions
    for c, idx in enumerate(vk.commitment_constraint_indexes):
        w_pow = pow(vk.omega, vk.nb_public + idx, q)
        tmp = pow((zeta + q - w_pow) % q, q - 2, q)
        tmp = tmp * (w_pow * zn % q) % q
        PI = (PI + hash_fr(cv, fs(BSB_COM[c])) * tmp) % q

    # :195-201
    res = (zeta + q - 1) % q
    res = pow(res, q - 2, q)
    res = res * zn % q
    res = res * alpha % q
    res = res * alpha % q
    alpha2Lagrange = res

    # :203-218
    s1 = I(S1_AT_Z) * beta % q
    s1 = (s1 + gamma + I(L_AT_Z)) % q
    s2 = I(S2_AT_Z) * beta % q
    s2 = (s2 + gamma + I(R_AT_Z)) % q
    o = (I(O_AT_Z) + gamma) % q
    s1 = s1 * s2 % q
    s1 = s1 * o % q
    s1 = s1 * alpha % q
    s1 = s1 * I(GRAND_PRODUCT_AT_Z_OMEGA) % q
    s1 = (s1 + PI + q - alpha2Lagrange) % q
    linearized_poly_at_z = (q - s1)      # note: may equal q when s1 == 0, as in the template

    # :220-229 folded H
    n2 = vk.size + 2
    zn2 = pow(zeta, n2, q)
    folded_h = ec.scalar_mul(H_2, U256(zn2))
    folded_h = ec.add(folded_h, H_1)
    folded_h = ec.scalar_mul(folded_h, U256(zn2))
    folded_h = ec.add(folded_h, H_0)
    znminus1 = (pow(zeta, vk.size, q) + q - 1) % q
    folded_h = ec.scalar_mul(folded_h, U256(znminus1))
    folded_h = invert(folded_h)

    # :231-278 linearisation commitment
    u = I(GRAND_PRODUCT_AT_Z_OMEGA) * beta % q
    v = I(S1_AT_Z) * beta % q
    v = (v + I(L_AT_Z) + gamma) % q
    w = I(S2_AT_Z) * beta % q
    w = (w + I(R_AT_Z) + gamma) % q
    s1 = u * v % q
    s1 = s1 * w % q
    s1 = s1 * alpha % q
    coset_square = vk.coset_shift * vk.coset_shift % q
    betazeta = beta * zeta % q
    u = (betazeta + I(L_AT_Z) + gamma) % q
    v = betazeta * vk.coset_shift % q
    v = (v + I(R_AT_Z) + gamma) % q
    w = betazeta * coset_square % q
    w = (w + I(O_AT_Z) + gamma) % q
    s2 = u * v % q
    s2 = q - (s2 * w % q)
    s2 = (s2 * alpha + alpha2Lagrange) % q

    lin_poly_com = ec.scalar_mul(VK_QL, L_AT_Z)
    lin_poly_com = ec.add(lin_poly_com, ec.scalar_mul(VK_QR, R_AT_Z))
    lin_poly_com = ec.add(lin_poly_com, ec.scalar_mul(VK_QO, O_AT_Z))
    ab = I(L_AT_Z) * I(R_AT_Z) % q
    lin_poly_com = ec.add(lin_poly_com, ec.scalar_mul(VK_QM, U256(ab)))
    lin_poly_com = ec.add(lin_poly_com, VK_QK)
    for c in range(k):
        lin_poly_com = ec.add(lin_poly_com, ec.scalar_mul(BSB_COM[c], QCP_AT_Z[c]))
    lin_poly_com = ec.add(lin_poly_com, ec.scalar_mul(VK_S3, U256(s1)))
    lin_poly_com = ec.add(lin_poly_com, ec.scalar_mul(GRAND_PRODUCT, U256(s2)))
    lin_poly_com = ec.add(lin_poly_com, folded_h)

    # :280-286 fold challenge
    linearized_poly_at_z_bytes = U256(linearized_poly_at_z % (1 << 256))
    r_pre = sha256(b"gamma" + U256(zeta) + lin_poly_com
                   + fs(L_COM) + fs(R_COM) + fs(O_COM) + VK_S1_fs + VK_S2_fs + b"".join(VK_QCP_fs)
                   + linearized_poly_at_z_bytes + L_AT_Z + R_AT_Z + O_AT_Z + S1_AT_Z + S2_AT_Z
                   + b"".join(QCP_AT_Z) + GRAND_PRODUCT_AT_Z_OMEGA)
    r = curvemod(r_pre)
    r_acc = r

    # :288-321 fold
    digest = lin_poly_com
    claims = linearized_poly_at_z
    for com, val in [(L_COM, L_AT_Z), (R_COM, R_AT_Z), (O_COM, O_AT_Z), (VK_S1, S1_AT_Z), (VK_S2, S2_AT_Z)] \
            + list(zip(VK_QCP, QCP_AT_Z)):
        digest = ec.add(digest, ec.scalar_mul(com, U256(r_acc)))
        claims = (claims + I(val) * r_acc) % q
        r_acc = r_acc * r % q

    # :323-356 final check
    r_pre = sha256(digest + BATCH_OPENING_AT_Z + fs(GRAND_PRODUCT) + OPENING_AT_Z_OMEGA + U256(zeta) + U256(r))
    r = curvemod(r_pre)
    quotient = BATCH_OPENING_AT_Z
    quotient = ec.add(quotient, ec.scalar_mul(OPENING_AT_Z_OMEGA, U256(r)))
    digest = ec.add(digest, ec.scalar_mul(GRAND_PRODUCT, U256(r)))
    claims = (claims + I(GRAND_PRODUCT_AT_Z_OMEGA) * r) % q
    G1_SRS = ec.enc(vk.g1)
    claims_com = ec.scalar_mul(G1_SRS, U256(claims))
    digest = ec.add(digest, invert(claims_com))
    points_quotient = ec.scalar_mul(BATCH_OPENING_AT_Z, U256(zeta))
    zeta_omega = zeta * vk.omega % q
    r = r * zeta_omega % q
    points_quotient = ec.add(points_quotient, ec.scalar_mul(OPENING_AT_Z_OMEGA, U256(r)))
    digest = ec.add(digest, points_quotient)
    quotient = invert(quotient)

    # ec.pairing_check(digest + quotient, g2):  e(digest, G2[0]) * e(quotient, G2[1]) == 1
    return pairing_check(vk, ec.dec(digest), ec.dec(quotient))


def pairing_check(vk: VerifyingKey, A: Affine, B: Affine) -> bool:
    """e(A, [1]_2) e(B, [tau]_2) == 1.  With a known-tau (TestOnly) SRS this is
    A + tau B == O in G1 (G1 has prime order r, e is non-degenerate).  With a real
    SRS it needs the G2 points: see oracle/pairing.py."""
    cv = vk.curve
    if vk.tau is not None:
        return g1_add(cv, A, g1_mul(cv, B, vk.tau)) is None
    if vk.g2 is None:
        raise ValueError("verifying key has neither tau nor G2 points")
    from . import pairing  # local import: heavy, rarely needed
    return pairing.pairing_product_is_one(cv, [(A, vk.g2[0]), (B, vk.g2[1])])
