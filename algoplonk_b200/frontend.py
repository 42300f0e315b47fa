"""Minimal sparse-R1CS front-end: builds the PLONK trace and the solved wire
vectors that gnark's frontend.Compile + solver hand to plonk.Prove
(/root/reference/algoplonk.go:50,81-89).

gnark's circuit compiler and witness solver stay on the CPU, in Go, in the
real integration ("stays in Go on CPU"); this module only exists so
that tests and bench.py can produce the *inputs* of the hot path without Go:
the selector columns ql,qr,qm,qo,qk, the copy permutation (gnark's
buildPermutation rule: last-seen position cycles) and L,R,O.

Row convention: rows 0..nb_public-1 are ql=-1 placeholders with
L[i] = public input i; constraint j sits on row nb_public + j and states
    ql*xa + qr*xb + qm*xa*xb + qo*xc + qk (+ qcp*pi2) == 0.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Sequence

R_MOD = {
    "BN254": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    "BLS12_381": 52435875175126190479447740508185965837690552500527637822603658699938581184513,
}
CURVE_ID = {"BN254": 0, "BLS12_381": 1}


def next_pow2(x: int) -> int:
    n = 1
    while n < x:
        n <<= 1
    return n


@dataclass
class Commitment:
    committed_rows: List[int]          # constraint indexes carrying qcp = 1
    commitment_row: int                # constraint index whose qk receives hash(commitment)
    var: int                           # variable that receives the hash value


@dataclass
class SparseR1CS:
    curve: str
    nb_public: int
    nb_variables: int
    # one tuple per constraint: (ql, qr, qm, qo, qk, xa, xb, xc)
    constraints: List[tuple]
    commitments: List[Commitment] = field(default_factory=list)

    @property
    def nb_constraints(self) -> int:
        return len(self.constraints)

    @property
    def domain_size(self) -> int:
        return next_pow2(self.nb_constraints + self.nb_public)


class Builder:
    """Eager builder: variables carry their values, so one pass yields both the
    constraint system and the solved witness (values mod r)."""

    def __init__(self, curve: str):
        self.curve = curve
        self.r = R_MOD[curve]
        self.values: List[int] = []
        self.nb_public = 0
        self.constraints: List[tuple] = []
        self.commitments: List[Commitment] = []
        self._secret_started = False
        # filled by commit(): callbacks the solver runs to obtain hash(commitment)
        self._commit_hooks: List[Callable] = []

    # -- variables ---------------------------------------------------------
    def public(self, value: int) -> int:
        assert not self._secret_started, "public variables first (gnark witness order)"
        self.values.append(value % self.r)
        self.nb_public += 1
        return len(self.values) - 1

    def secret(self, value: int) -> int:
        self._secret_started = True
        self.values.append(value % self.r)
        return len(self.values) - 1

    internal = secret

    # -- constraints -------------------------------------------------------
    def add_constraint(self, ql=0, qr=0, qm=0, qo=0, qk=0, xa=0, xb=0, xc=0) -> int:
        r = self.r
        self.constraints.append((ql % r, qr % r, qm % r, qo % r, qk % r, xa, xb, xc))
        return len(self.constraints) - 1

    def mul(self, a: int, b: int) -> int:
        c = self.internal(self.values[a] * self.values[b])
        self.add_constraint(qm=1, qo=-1, xa=a, xb=b, xc=c)
        return c

    def add(self, a: int, b: int) -> int:
        c = self.internal(self.values[a] + self.values[b])
        self.add_constraint(ql=1, qr=1, qo=-1, xa=a, xb=b, xc=c)
        return c

    def assert_is_equal(self, a: int, b: int) -> None:
        self.add_constraint(ql=1, qr=-1, xa=a, xb=b, xc=0)

    def assert_is_different_from_zero(self, a: int) -> None:
        v = self.values[a]
        inv = self.internal(pow(v, -1, self.r) if v else 0)
        self.add_constraint(qm=1, qk=-1, xa=a, xb=inv, xc=0)

    def commit(self, variables: Sequence[int], hash_of_commitment: Callable[[List[int], int, int], int]) -> int:
        """frontend.Committer.Commit (BSB22).  `hash_of_commitment(rows, values,
        commitment_row)` is the solver hint: it receives the committed constraint
        rows / values and returns hash_fr(commitment point)."""
        rows = []
        for v in variables:
            rows.append(self.add_constraint(ql=-1, xa=v, xb=0, xc=0))      # -v + qcp*pi2 = 0
        commitment_row = len(self.constraints)
        h = hash_of_commitment(rows, [self.values[v] for v in variables], commitment_row)
        cvar = self.internal(h)
        self.add_constraint(ql=-1, xa=cvar, xb=0, xc=0)                     # -cmt + qk(=hash) = 0
        self.commitments.append(Commitment(rows, commitment_row, cvar))
        return cvar

    def build(self) -> SparseR1CS:
        return SparseR1CS(self.curve, self.nb_public, len(self.values), list(self.constraints),
                          list(self.commitments))


@dataclass
class TraceColumns:
    """Lagrange-form columns + permutation, integers mod r (canonical, not Montgomery)."""
    curve: str
    n: int
    nb_public: int
    ql: List[int]
    qr: List[int]
    qm: List[int]
    qo: List[int]
    qk: List[int]
    perm: List[int]
    qcp: List[List[int]]
    commitment_constraint_indexes: List[int]


def build_permutation(cs: SparseR1CS, n: int) -> List[int]:
    """gnark buildPermutation: position -> previous position of the same variable, cycles closed
    by the last position seen.  Padding rows reference variable 0."""
    size = 3 * n
    lro = [0] * size
    for i in range(cs.nb_public):
        lro[i] = i
    off = cs.nb_public
    for j, (_, _, _, _, _, xa, xb, xc) in enumerate(cs.constraints):
        lro[off + j] = xa
        lro[n + off + j] = xb
        lro[2 * n + off + j] = xc
    perm = [-1] * size
    cycle = [-1] * max(cs.nb_variables, 1)
    for i in range(size):
        v = lro[i]
        if cycle[v] != -1:
            perm[i] = cycle[v]
        cycle[v] = i
    for i in range(size):
        if perm[i] == -1:
            perm[i] = cycle[lro[i]]
    return perm


def build_trace(cs: SparseR1CS) -> TraceColumns:
    r = R_MOD[cs.curve]
    n = cs.domain_size
    ql, qr, qm, qo, qk = ([0] * n for _ in range(5))
    for i in range(cs.nb_public):
        ql[i] = r - 1
    off = cs.nb_public
    for j, (a, b, m, o, k, _, _, _) in enumerate(cs.constraints):
        ql[off + j], qr[off + j], qm[off + j], qo[off + j], qk[off + j] = a, b, m, o, k
    qcp = []
    for c in cs.commitments:
        col = [0] * n
        for row in c.committed_rows:
            col[off + row] = 1
        qcp.append(col)
    return TraceColumns(cs.curve, n, cs.nb_public, ql, qr, qm, qo, qk, build_permutation(cs, n), qcp,
                        [c.commitment_row for c in cs.commitments])


def solve_lro(cs: SparseR1CS, values: Sequence[int], n: int):
    """L,R,O in Lagrange form (what spr.Solve returns), padding rows -> variable 0's value."""
    v0 = values[0] if values else 0
    L, R, O = [v0] * n, [v0] * n, [v0] * n
    for i in range(cs.nb_public):
        L[i] = values[i]
    off = cs.nb_public
    for j, (_, _, _, _, _, xa, xb, xc) in enumerate(cs.constraints):
        L[off + j], R[off + j], O[off + j] = values[xa], values[xb], values[xc]
    return L, R, O


def check_gates(tc: TraceColumns, L, R, O, pi2=()) -> bool:
    """Plain constraint check with public inputs written into qk (gnark completeQk)."""
    r = R_MOD[tc.curve]
    for i in range(tc.n):
        qk = L[i] if i < tc.nb_public else tc.qk[i]
        acc = tc.ql[i] * L[i] + tc.qr[i] * R[i] + tc.qm[i] * L[i] * R[i] + tc.qo[i] * O[i] + qk
        for c, col in enumerate(tc.qcp):
            acc += col[i] * pi2[c][i]
        if acc % r:
            return False
    return True


# ---------------------------------------------------------------------------
# Circuits used by the reference's examples/tests and by BASELINE.json configs
# ---------------------------------------------------------------------------

def basic_circuit(curve: str, a: int = 3, b: int = 4, c: int = 5) -> Builder:
    """examples/basic/logicsigVerifier/main.go:30-52: a*a + b*b == c*c, a,b public."""
    B = Builder(curve)
    A_, B_ = B.public(a), B.public(b)
    C_ = B.secret(c)
    aa, bb, cc = B.mul(A_, A_), B.mul(B_, B_), B.mul(C_, C_)
    B.assert_is_equal(B.add(aa, bb), cc)
    return B


def bsb22_circuit(curve: str, nb_commitments: int, hash_hint, x: int = 9, y: int = 3) -> Builder:
    """bsb22_test.go:18-39: X == Y*Y, then nb_commitments x (Commit(Y, X) != 0)."""
    B = Builder(curve)
    X = B.public(x)
    Y = B.secret(y)
    B.assert_is_equal(X, B.mul(Y, Y))
    for _ in range(nb_commitments):
        cmt = B.commit([Y, X], hash_hint)
        B.assert_is_different_from_zero(cmt)
    return B


def squaring_chain(curve: str, log2_rows: int, x0: int = 2):
    """SURVEY 8d synthetic benchmark: x_{i+1} = x_i^2, public y = x_m, sized so that
    nb_public + nb_constraints == 2^log2_rows exactly.  Returns (SparseR1CS, values)
    without going through Builder (it has to be fast at 2^20 rows)."""
    r = R_MOD[curve]
    rows = 1 << log2_rows
    m = rows - 2                       # 1 public row + m squarings + 1 equality row
    values = [0] * (m + 2)             # var 0 = y (public), var 1 = x0, var 1+i = x_i
    x = x0 % r
    values[1] = x
    for i in range(m):
        x = x * x % r
        values[2 + i] = x
    values[0] = x
    one, neg1 = 1, r - 1
    constraints = [(0, 0, one, neg1, 0, 1 + i, 1 + i, 2 + i) for i in range(m)]
    constraints.append((one, neg1, 0, 0, 0, 0, 1 + m, 0))          # y == x_m
    cs = SparseR1CS(curve, 1, m + 2, constraints)
    return cs, values


def random_dense_circuit(curve: str, log2_rows: int, seed: int = 0, nb_public: int = 3):
    """SURVEY 8d "random-dense" shape: every selector column carries full-width random field elements and the
    wires are reused at random, so the permutation has long, irregular cycles -- the squaring chain only ever
    exercises qm = 1, qo = -1.  Gate i:  ql*a + qr*b + qm*a*b + qo*c + qk = 0  with a, b drawn from the
    variables seen so far and c a fresh variable solved from the gate (one gate in eight reuses an existing
    variable as c and solves qk instead).  nb_public + nb_constraints == 2^log2_rows exactly.
    Returns (SparseR1CS, values)."""
    import random
    r = R_MOD[curve]
    rng = random.Random(seed)
    rows = 1 << log2_rows
    assert rows > nb_public >= 1
    values = [rng.randrange(r) for _ in range(nb_public)]
    values.append(rng.randrange(r))                      # one secret input
    constraints = []
    for i in range(rows - nb_public):
        a, b = rng.randrange(len(values)), rng.randrange(len(values))
        ql, qr, qm = (0 if rng.random() < 0.1 else rng.randrange(r) for _ in range(3))
        lin = (ql * values[a] + qr * values[b] + qm * values[a] * values[b]) % r
        if i % 8 == 7:
            c, qo = rng.randrange(len(values)), rng.randrange(r)
            qk = (-(lin + qo * values[c])) % r
        else:
            qo, qk = rng.randrange(1, r), (0 if rng.random() < 0.5 else rng.randrange(r))
            values.append((-(lin + qk)) * pow(qo, -1, r) % r)
            c = len(values) - 1
        constraints.append((ql, qr, qm, qo, qk, a, b, c))
    return SparseR1CS(curve, nb_public, len(values), constraints), values
