"""Pairing check for the restated verifier (TEST INFRASTRUCTURE, like the rest of oracle/).

The reference's generated verifiers end with `ec_pairing_check` on (digest, quotient) against the two G2
points of the verifying key (/root/reference/verifier/templateLogicSigBN254.go:326-356, 21-28).  With a
known-tau SRS plonk_oracle.pairing_check evaluates that in G1; for the REAL ceremony files (tau unknown) it
needs an actual pairing -- this module.  Nothing of it runs on the product path.

Deliberately the slow, obviously-correct construction: the degree-12 extension as polynomials modulo
w^12 - 18 w^6 + 82 (BN254) / w^12 - 2 w^6 + 2 (BLS12-381), G2 carried into E(Fp12) through the sextic twist,
a textbook Miller loop with affine line functions and a plain final exponentiation by (p^12 - 1) / r.  Any
non-degenerate bilinear map decides "product == 1" identically, so no convention has to match gnark's.
Pinned by bilinearity and by the ceremony files themselves: e([tau]_1, [1]_2) == e([1]_1, [tau]_2) with
both sides taken from the reference's pk.bin / vk.bin (tests/test_oracle.py).

Also here: gnark's compressed G2 encoding (setup/<name>/vk.bin = 2 compressed G2 + 1 compressed G1,
setup/setup.go:216-225; BLS12-381 known answers at setup/trusted_setup_test.go:93-95).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from .plonk_oracle import CurveParams

Fp2 = Tuple[int, int]                      # a0 + a1 * u,  u^2 = -1
G2Affine = Optional[Tuple[Fp2, Fp2]]


# ---------------------------------------------------------------------------------------------
# Fp2 (only what G2 decompression needs)
# ---------------------------------------------------------------------------------------------
def f2_mul(p: int, a: Fp2, b: Fp2) -> Fp2:
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def f2_add(p: int, a: Fp2, b: Fp2) -> Fp2:
    return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)


def f2_inv(p: int, a: Fp2) -> Fp2:
    n = pow((a[0] * a[0] + a[1] * a[1]) % p, -1, p)
    return (a[0] * n % p, (-a[1]) * n % p)


def _fp_sqrt(p: int, a: int) -> Optional[int]:
    s = pow(a, (p + 1) // 4, p)            # both base fields are 3 mod 4
    return s if s * s % p == a % p else None


def f2_sqrt(p: int, a: Fp2) -> Optional[Fp2]:
    """Square root in Fp[u]/(u^2+1) through the norm: x0^2 = (a0 +- sqrt(a0^2 + a1^2)) / 2, x1 = a1 / (2 x0)."""
    a0, a1 = a[0] % p, a[1] % p
    if a1 == 0:
        s = _fp_sqrt(p, a0)
        if s is not None:
            return (s, 0)
        s = _fp_sqrt(p, (-a0) % p)         # -1 is a non-residue: exactly one of a0, -a0 is a square
        return None if s is None else (0, s)
    n = _fp_sqrt(p, (a0 * a0 + a1 * a1) % p)
    if n is None:
        return None
    half = pow(2, -1, p)
    for cand in ((a0 + n) * half % p, (a0 - n) * half % p):
        x0 = _fp_sqrt(p, cand)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, p) % p
            if f2_mul(p, (x0, x1), (x0, x1)) == (a0, a1):
                return (x0, x1)
    return None


def twist_b(cv: CurveParams) -> Fp2:
    """Coefficient of the sextic twist: BN254 is a D-twist, b' = 3 / (9 + u); BLS12-381 an M-twist, b' = 4 (1 + u)."""
    if cv.cid == 0:
        inv = f2_inv(cv.p, (9, 1))
        return (cv.b * inv[0] % cv.p, cv.b * inv[1] % cv.p)
    return (cv.b % cv.p, cv.b % cv.p)


def g2_on_curve(cv: CurveParams, Q: G2Affine) -> bool:
    if Q is None:
        return True
    x, y = Q
    return f2_mul(cv.p, y, y) == f2_add(cv.p, f2_mul(cv.p, f2_mul(cv.p, x, x), x), twist_b(cv))


def _lex_largest_fp(p: int, v: int) -> bool:
    return v > (p - 1) // 2


def g2_decompress(cv: CurveParams, b: bytes) -> G2Affine:
    """gnark compressed G2: X.A1 || X.A0 big-endian, flags in the top bits of byte 0 (BN254: 10 smallest y,
    11 largest y, 01 infinity; BLS12-381: 100 / 101 / 110).  y is 'largest' by A1, or by A0 when A1 = 0."""
    nb = cv.fp_bytes
    assert len(b) == 2 * nb
    if cv.cid == 0:
        flag, mask, small, large, inf = b[0] >> 6, 0x3F, 0b10, 0b11, 0b01
    else:
        flag, mask, small, large, inf = b[0] >> 5, 0x1F, 0b100, 0b101, 0b110
    if flag == inf:
        return None
    if flag not in (small, large):
        raise ValueError("uncompressed/invalid flag in compressed G2")
    x1 = int.from_bytes(bytes([b[0] & mask]) + b[1:nb], "big")
    x0 = int.from_bytes(b[nb:], "big")
    if x0 >= cv.p or x1 >= cv.p:
        raise ValueError("G2 coordinate not reduced")
    x = (x0, x1)
    y = f2_sqrt(cv.p, f2_add(cv.p, f2_mul(cv.p, f2_mul(cv.p, x, x), x), twist_b(cv)))
    if y is None:
        raise ValueError("x not on the twist")
    largest = _lex_largest_fp(cv.p, y[1]) if y[1] else _lex_largest_fp(cv.p, y[0])
    if largest != (flag == large):
        y = ((-y[0]) % cv.p, (-y[1]) % cv.p)
    return (x, y)


def parse_vk_bin(cv: CurveParams, vk_bin: bytes):
    """setup/<name>/vk.bin (kzg.VerifyingKey.WriteTo): G2[0], G2[1] compressed, then G1 compressed."""
    from .plonk_oracle import g1_decompress
    nb = cv.fp_bytes
    if len(vk_bin) != 5 * nb:
        raise ValueError("vk.bin has the wrong length")
    return (g2_decompress(cv, vk_bin[:2 * nb]), g2_decompress(cv, vk_bin[2 * nb:4 * nb])), g1_decompress(cv, vk_bin[4 * nb:])


# ---------------------------------------------------------------------------------------------
# G2 group law in affine coordinates over Fp2 (the checker of b2p_msm_g2: G2Affine.MultiExp, gnark-crypto
# ecc/<curve>/multiexp.go, which AlgoPlonk reaches only through kzg.NewSRS -- setup/setup.go:124)
# ---------------------------------------------------------------------------------------------
def f2_sub(p: int, a: Fp2, b: Fp2) -> Fp2:
    return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)


def g2_neg(cv: CurveParams, Q: G2Affine) -> G2Affine:
    return None if Q is None else (Q[0], ((-Q[1][0]) % cv.p, (-Q[1][1]) % cv.p))


def g2_add(cv: CurveParams, P: G2Affine, Q: G2Affine) -> G2Affine:
    """Chord-and-tangent on y^2 = x^3 + b' (a = 0)."""
    p = cv.p
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if P[1] != Q[1] or P[1] == (0, 0):
            return None
        xx = f2_mul(p, P[0], P[0])
        lam = f2_mul(p, (3 * xx[0] % p, 3 * xx[1] % p), f2_inv(p, (2 * P[1][0] % p, 2 * P[1][1] % p)))
    else:
        lam = f2_mul(p, f2_sub(p, Q[1], P[1]), f2_inv(p, f2_sub(p, Q[0], P[0])))
    x3 = f2_sub(p, f2_sub(p, f2_mul(p, lam, lam), P[0]), Q[0])
    y3 = f2_sub(p, f2_mul(p, lam, f2_sub(p, P[0], x3)), P[1])
    return (x3, y3)


def g2_mul(cv: CurveParams, Q: G2Affine, k: int) -> G2Affine:
    """Double-and-add, k taken as a non-negative integer (not reduced: callers test the group order with it)."""
    acc, base = None, Q
    while k:
        if k & 1:
            acc = g2_add(cv, acc, base)
        base = g2_add(cv, base, base)
        k >>= 1
    return acc


def g2_msm_naive(cv: CurveParams, points: Sequence[G2Affine], scalars: Sequence[int]) -> G2Affine:
    acc = None
    for Q, k in zip(points, scalars):
        acc = g2_add(cv, acc, g2_mul(cv, Q, k % cv.r))
    return acc


# ---------------------------------------------------------------------------------------------
# Fp12 = Fp[w] / (w^12 - c6 w^6 + c0) as plain polynomials
# ---------------------------------------------------------------------------------------------
class _Ext:
    """Arithmetic of one degree-12 extension; elements are lists of 12 ints (low degree first)."""

    def __init__(self, p: int, c0: int, c6: int):
        # w^12 = c6 * w^6 - c0
        self.p, self.c0, self.c6 = p, c0, c6
        self.modulus = [c0, 0, 0, 0, 0, 0, (-c6) % p, 0, 0, 0, 0, 0, 1]
        self.one = [1] + [0] * 11
        self.zero = [0] * 12

    def const(self, v: int) -> List[int]:
        return [v % self.p] + [0] * 11

    def add(self, a, b):
        p = self.p
        return [(x + y) % p for x, y in zip(a, b)]

    def sub(self, a, b):
        p = self.p
        return [(x - y) % p for x, y in zip(a, b)]

    def neg(self, a):
        p = self.p
        return [(-x) % p for x in a]

    def mul(self, a, b):
        p, c0, c6 = self.p, self.c0, self.c6
        t = [0] * 23
        for i, x in enumerate(a):
            if x:
                for j, y in enumerate(b):
                    t[i + j] += x * y
        for k in range(22, 11, -1):         # w^k = c6 w^(k-6) - c0 w^(k-12)
            v = t[k]
            if v:
                t[k - 6] += c6 * v
                t[k - 12] -= c0 * v
        return [v % p for v in t[:12]]

    def muls(self, a, s: int):
        p = self.p
        return [x * s % p for x in a]

    def eq(self, a, b) -> bool:
        return all((x - y) % self.p == 0 for x, y in zip(a, b))

    def pow(self, a, e: int):
        r, base = self.one, a
        while e:
            if e & 1:
                r = self.mul(r, base)
            base = self.mul(base, base)
            e >>= 1
        return r

    @staticmethod
    def _deg(a) -> int:
        d = len(a) - 1
        while d > 0 and a[d] == 0:
            d -= 1
        return d

    def inv(self, a):
        """Extended Euclid on polynomials over Fp (the modulus is irreducible, so the gcd is a constant)."""
        p, deg = self.p, self._deg

        def divmod_(num, den):
            num, dd = num[:], deg(den)
            lead = pow(den[dd], -1, p)
            q = [0] * 13
            dn = deg(num)
            while dn >= dd and any(num):
                c, sh = num[dn] * lead % p, dn - dd
                q[sh] = c
                for k in range(dd + 1):
                    num[k + sh] = (num[k + sh] - c * den[k]) % p
                dn = deg(num)
            return q, num

        r0, r1 = self.modulus[:], [x % p for x in a] + [0]
        t0, t1 = [0] * 13, [1] + [0] * 12
        while any(r1):
            q, r = divmod_(r0, r1)
            t2 = t0[:]
            for i, qi in enumerate(q):
                if qi:
                    for j, tj in enumerate(t1):
                        if tj and i + j < 13:
                            t2[i + j] = (t2[i + j] - qi * tj) % p
            r0, r1, t0, t1 = r1, r, t1, t2
        if deg(r0) != 0 or r0[0] == 0:
            raise ZeroDivisionError("not invertible")
        s = pow(r0[0], -1, p)
        return [x * s % p for x in t0[:12]]

    def div(self, a, b):
        return self.mul(a, self.inv(b))


def _ext(cv: CurveParams) -> _Ext:
    return _Ext(cv.p, 82, 18) if cv.cid == 0 else _Ext(cv.p, 2, 2)


def _embed_fp2(F: _Ext, cv: CurveParams, a: Fp2):
    """u = w^6 - 9 (BN254: (w^6 - 9)^2 = -1) / u = w^6 - 1 (BLS12-381)."""
    k = 9 if cv.cid == 0 else 1
    out = [0] * 12
    out[0] = (a[0] - k * a[1]) % cv.p
    out[6] = a[1] % cv.p
    return out


def _twist(F: _Ext, cv: CurveParams, Q):
    """E'(Fp2) -> E(Fp12).  xi = 9 + u (BN254) / 1 + u (BLS12-381) equals w^6 in both towers.
    D-twist (b' = b / xi): (x, y) -> (x w^2, y w^3);  M-twist (b' = b xi): (x, y) -> (x / w^2, y / w^3)."""
    x, y = _embed_fp2(F, cv, Q[0]), _embed_fp2(F, cv, Q[1])
    w2 = [0, 0, 1] + [0] * 9
    w3 = [0, 0, 0, 1] + [0] * 8
    if cv.cid == 0:
        return (F.mul(x, w2), F.mul(y, w3))
    return (F.div(x, w2), F.div(y, w3))


def _double(F: _Ext, P):
    x, y = P
    m = F.div(F.muls(F.mul(x, x), 3), F.muls(y, 2))
    nx = F.sub(F.mul(m, m), F.muls(x, 2))
    return (nx, F.sub(F.mul(m, F.sub(x, nx)), y))


def _add(F: _Ext, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    (x1, y1), (x2, y2) = P, Q
    if F.eq(x1, x2):
        return _double(F, P) if F.eq(y1, y2) else None
    m = F.div(F.sub(y2, y1), F.sub(x2, x1))
    nx = F.sub(F.sub(F.mul(m, m), x1), x2)
    return (nx, F.sub(F.mul(m, F.sub(x1, nx)), y1))


def _line(F: _Ext, P1, P2, T):
    """The line through P1 and P2 (tangent when equal), evaluated at T."""
    (x1, y1), (x2, y2), (xt, yt) = P1, P2, T
    if not F.eq(x1, x2):
        m = F.div(F.sub(y2, y1), F.sub(x2, x1))
    elif F.eq(y1, y2):
        m = F.div(F.muls(F.mul(x1, x1), 3), F.muls(y1, 2))
    else:
        return F.sub(xt, x1)
    return F.sub(F.mul(m, F.sub(xt, x1)), F.sub(yt, y1))


ATE_LOOP = {0: 29793968203157093288,        # 6 x + 2, x = 4965661367192848881
            1: 15132376222941642752}        # |x| = 0xd201000000010000


def miller_loop(cv: CurveParams, Q: G2Affine, P) -> List[int]:
    """f_{loop,Q}(P) before the final exponentiation; Q on the twist over Fp2, P affine in G1."""
    F = _ext(cv)
    if Q is None or P is None:
        return F.one
    Qe = _twist(F, cv, Q)
    Pe = (F.const(P[0]), F.const(P[1]))
    R, f = Qe, F.one
    n = ATE_LOOP[cv.cid]
    for i in range(n.bit_length() - 2, -1, -1):
        f = F.mul(F.mul(f, f), _line(F, R, R, Pe))
        R = _double(F, R)
        if (n >> i) & 1:
            f = F.mul(f, _line(F, R, Qe, Pe))
            R = _add(F, R, Qe)
    if cv.cid == 0:                            # BN: two more lines through the Frobenius images of Q
        q1 = (F.pow(Qe[0], cv.p), F.pow(Qe[1], cv.p))
        nq2 = (F.pow(q1[0], cv.p), F.neg(F.pow(q1[1], cv.p)))
        f = F.mul(f, _line(F, R, q1, Pe))
        R = _add(F, R, q1)
        f = F.mul(f, _line(F, R, nq2, Pe))
    return f


def final_exponentiation(cv: CurveParams, f: List[int]) -> List[int]:
    return _ext(cv).pow(f, (cv.p ** 12 - 1) // cv.r)


def pairing(cv: CurveParams, P, Q: G2Affine) -> List[int]:
    return final_exponentiation(cv, miller_loop(cv, Q, P))


def pairing_product_is_one(cv: CurveParams, pairs: Sequence[Tuple[object, G2Affine]]) -> bool:
    """prod_i e(P_i, Q_i) == 1 with one final exponentiation (what AVM ec_pairing_check decides)."""
    F = _ext(cv)
    f = F.one
    for P, Q in pairs:
        f = F.mul(f, miller_loop(cv, Q, P))
    return F.eq(final_exponentiation(cv, f), F.one)
