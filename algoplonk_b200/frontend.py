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
from typing import Callable, List, Optional, Sequence

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
class Hint:
    """A solver hint (gnark: solver.Hint): out_vars = f(in_vars), f identified by `id` and supplied at solve time."""
    id: int
    in_vars: List[int]
    out_vars: List[int]


HINT_NBITS = 1            # bits of one variable, least significant first (gnark: bits.NBits)
HINT_BSB22 = 0x100        # + commitment index: hash of the commitment to the committed variables (gnark: bsb22CommitmentComputePlaceholder)


@dataclass
class SparseR1CS:
    curve: str
    nb_public: int
    nb_variables: int
    # one tuple per constraint: (ql, qr, qm, qo, qk, xa, xb, xc)
    constraints: List[tuple]
    commitments: List[Commitment] = field(default_factory=list)
    # variables the caller assigns (frontend.NewWitness: public, then secret); None: the first nb_public + 1 ...
    # are not known -- circuits built by Builder / the generators below always record them
    input_vars: Optional[List[int]] = None
    hints: List[Hint] = field(default_factory=list)
    unchecked_rows: List[int] = field(default_factory=list)    # constraint indexes whose gate the prover completes (BSB22)

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
        self.input_vars: List[int] = []          # variables assigned by the caller, in declaration order
        self.hints: List[Hint] = []
        self.unchecked_rows: List[int] = []

    # -- variables ---------------------------------------------------------
    def public(self, value: int) -> int:
        assert not self._secret_started, "public variables first (gnark witness order)"
        self.values.append(value % self.r)
        self.nb_public += 1
        self.input_vars.append(len(self.values) - 1)
        return len(self.values) - 1

    def secret(self, value: int) -> int:
        self._secret_started = True
        self.values.append(value % self.r)
        self.input_vars.append(len(self.values) - 1)
        return len(self.values) - 1

    def internal(self, value: int) -> int:
        """A variable the solver determines from a constraint (not part of the witness assignment)."""
        self._secret_started = True
        self.values.append(value % self.r)
        return len(self.values) - 1

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

    def hint(self, hint_id: int, in_vars: Sequence[int], out_values: Sequence[int]) -> List[int]:
        """Variables a solver hint produces (api.Compiler.NewHint): the builder knows their values, the constraint
        system only records which function makes them from what."""
        outs = [self.internal(v) for v in out_values]
        self.hints.append(Hint(hint_id, list(in_vars), outs))
        return outs

    def to_binary(self, a: int, nbits: int) -> List[int]:
        """api.ToBinary: nbits boolean variables from the NBits hint, each constrained b*b = b, recomposing to a."""
        bits = self.hint(HINT_NBITS, [a], [(self.values[a] >> i) & 1 for i in range(nbits)])
        acc = None
        for i, b in enumerate(bits):
            self.add_constraint(qm=1, ql=-1, xa=b, xb=b)             # b * b - b = 0
            if acc is None:
                acc = b
            else:
                nxt = self.internal(self.values[acc] + (self.values[b] << i))
                self.add_constraint(ql=1, qr=1 << i, qo=-1, xa=acc, xb=b, xc=nxt)
                acc = nxt
        self.assert_is_equal(acc, a)
        return bits

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
        cvar = self.hint(HINT_BSB22 + len(self.commitments), list(variables), [h])[0]
        self.add_constraint(ql=-1, xa=cvar, xb=0, xc=0)                     # -cmt + qk(=hash) = 0
        self.commitments.append(Commitment(rows, commitment_row, cvar))
        self.unchecked_rows += rows + [commitment_row]                       # the prover adds qcp * pi2 and the hash
        return cvar

    def build(self) -> SparseR1CS:
        return SparseR1CS(self.curve, self.nb_public, len(self.values), list(self.constraints),
                          list(self.commitments), list(self.input_vars), list(self.hints), list(self.unchecked_rows))


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


def solver_wires(cs: SparseR1CS, n: int):
    """The variable of every row's L, R, O wire (xa, xb, xc of b2p_solver_create): public rows carry their variable
    on L, padding rows and unused wires variable 0 -- the same rule solve_lro applies to values."""
    xa, xb, xc = [0] * n, [0] * n, [0] * n
    for i in range(cs.nb_public):
        xa[i] = i
    off = cs.nb_public
    for j, (_, _, _, _, _, a, b, c) in enumerate(cs.constraints):
        xa[off + j], xb[off + j], xc[off + j] = a, b, c
    return xa, xb, xc


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
    cs = SparseR1CS(curve, 1, m + 2, constraints, input_vars=[0, 1])
    return cs, values


def wide_mimc_circuit(curve: str, lanes: int, rounds: int, seed: int = 1):
    """A wide, shallow circuit (SURVEY 8f rank 4's case for solving on the GPU): `lanes` independent MiMC-style
    permutations of `rounds` rounds  x <- (x + c_i)^5  (one addition gate with a constant, three multiplication
    gates), e.g. the leaves of a Merkle tree hashed side by side.  Public: the output of lane 0.  Depth 4*rounds
    levels, every level `lanes` rows wide.  Returns (SparseR1CS, values); built directly, like squaring_chain."""
    import random
    r = R_MOD[curve]
    rng = random.Random(seed)
    consts = mimc_constants(curve)
    neg1 = r - 1
    values = [0] + [rng.randrange(r) for _ in range(lanes)]      # var 0 = public output of lane 0, then the lane inputs
    cur = list(range(1, lanes + 1))
    constraints = []

    def gate_level(make):                                        # one row per lane: row order = level order
        for j in range(lanes):
            row, val = make(j)
            values.append(val % r)
            constraints.append(row + (len(values) - 1,))
            cur_next[j] = len(values) - 1

    for i in range(rounds):
        c = consts[i % len(consts)]
        cur_next = [0] * lanes
        gate_level(lambda j: ((1, 0, 0, neg1, c, cur[j], 0), values[cur[j]] + c))                 # t = x + c
        t = cur_next
        cur_next = [0] * lanes
        gate_level(lambda j: ((0, 0, 1, neg1, 0, t[j], t[j]), values[t[j]] * values[t[j]]))       # t^2
        t2 = cur_next
        cur_next = [0] * lanes
        gate_level(lambda j: ((0, 0, 1, neg1, 0, t2[j], t2[j]), values[t2[j]] * values[t2[j]]))   # t^4
        t4 = cur_next
        cur_next = [0] * lanes
        gate_level(lambda j: ((0, 0, 1, neg1, 0, t4[j], t[j]), values[t4[j]] * values[t[j]]))     # t^5
        cur = cur_next
    values[0] = values[cur[0]]
    constraints.append((1, neg1, 0, 0, 0, 0, cur[0], 0))         # y == lane 0's output
    cs = SparseR1CS(curve, 1, len(values), constraints, input_vars=list(range(0, lanes + 1)))
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
    return SparseR1CS(curve, nb_public, len(values), constraints, input_vars=list(range(nb_public + 1))), values


# ---------------------------------------------------------------------------
# MiMC + Merkle proof: the circuit of the reference's examples/merkle and of its integration tests
# (examples/merkle/logicsigVerifier/main.go:35-61, testutils/verifier_integration_test.go)
# ---------------------------------------------------------------------------
_KECCAK_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
              0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
              0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
              0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
              0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_KECCAK_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def keccak256_legacy(data: bytes) -> bytes:
    """Keccak-256 with the original 0x01 padding (Go's sha3.NewLegacyKeccak256; hashlib only has the NIST 0x06
    variant).  Used for the MiMC round constants exactly as gnark-crypto derives them."""
    rate, M = 136, (1 << 64) - 1
    msg = bytearray(data) + b"\x01" + bytes((-len(data) - 2) % rate) + b"\x80" if (len(data) + 1) % rate else \
        bytearray(data) + b"\x81"
    A = [[0] * 5 for _ in range(5)]
    rol = lambda v, r: ((v << r) | (v >> (64 - r))) & M if r else v
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        for rc in _KECCAK_RC:
            Cc = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
            D = [Cc[(x - 1) % 5] ^ rol(Cc[(x + 1) % 5], 1) for x in range(5)]
            A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
            B = [[0] * 5 for _ in range(5)]
            for x in range(5):
                for y in range(5):
                    B[y][(2 * x + 3 * y) % 5] = rol(A[x][y], _KECCAK_ROT[x][y])
            A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
            A[0][0] ^= rc
    return b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


_MIMC_ROUNDS = {"BN254": 110, "BLS12_381": 111}
_MIMC_CONSTANTS: dict = {}


def mimc_constants(curve: str) -> List[int]:
    """gnark-crypto ecc/<curve>/fr/mimc initConstants [UPSTREAM-RECALL: un-vendored dependency, go.mod:9]: seed "seed",
    rnd = Keccak(seed); then repeatedly rnd = Keccak(rnd), constant_i = rnd mod r."""
    if curve not in _MIMC_CONSTANTS:
        rnd = keccak256_legacy(b"seed")
        out = []
        for _ in range(_MIMC_ROUNDS[curve]):
            rnd = keccak256_legacy(rnd)
            out.append(int.from_bytes(rnd, "big") % R_MOD[curve])
        _MIMC_CONSTANTS[curve] = out
    return _MIMC_CONSTANTS[curve]


def mimc_hash(curve: str, blocks: Sequence[int]) -> int:
    """MiMC (x^5 rounds) in Miyaguchi-Preneel mode over field elements: what the Merkle example hashes with
    (examples/merkle/logicsigVerifier/main.go:19,177-190 `mimcHash`), outside any circuit."""
    r, cs_ = R_MOD[curve], mimc_constants(curve)
    h = 0
    for m in blocks:
        x = m % r
        for c in cs_:
            t = (x + h + c) % r
            x = pow(t, 5, r)
        x = (x + h) % r
        h = (h + x + m) % r
    return h


def _mimc_in_circuit(B: Builder, h: int, blocks: Sequence[int]) -> int:
    """std/hash/mimc: per block  r = encrypt(block, key = h);  h = h + r + block.  Returns the variable of the digest.
    A round is  t = m + h + c_i  (one addition gate with a constant) and  m = t^5  (three multiplication gates)."""
    cs_ = mimc_constants(B.curve)
    for m0 in blocks:
        m = m0
        for c in cs_:
            t = B.internal(B.values[m] + B.values[h] + c)
            B.add_constraint(ql=1, qr=1, qo=-1, qk=c, xa=m, xb=h, xc=t)
            t2 = B.mul(t, t)
            t4 = B.mul(t2, t2)
            m = B.mul(t4, t)
        enc = B.add(m, h)
        h = B.add(B.add(h, enc), m0)
    return h


def merkle_circuit(curve: str, depth: int = 16, nb_leaves: int = 6, index: int = 3, bits_from_hint: bool = False):
    """MerkleCircuit (examples/merkle/logicsigVerifier/main.go:45-61): public RootHash, secret Path[depth + 1] (the
    unhashed leaf, then the siblings up to the root) and Index; Define() = merkle.MerkleProof.VerifyProof with MiMC:
    the leaf is hashed, the index is split into bits, every level selects (left, right) by its bit and hashes them,
    the result must equal RootHash.  The tree of main.go:63-92: `nb_leaves` leaves "leaf<i>", zero elsewhere, proof for
    leaf `index`.  The gate-level shape is this front end's (gnark's compiler is not available here); the
    statement proved and the witness are the example's.  Returns (Builder, root)."""
    r = R_MOD[curve]
    leaves = [int.from_bytes(b"leaf%d" % i, "big") % r for i in range(nb_leaves)]
    H = lambda *xs: mimc_hash(curve, xs)
    level = [H(x) for x in leaves]
    zero = [H(0)]                                # zero[i]: node at level i over uninitialised leaves (main.go:197-208)
    for _ in range(depth):
        zero.append(H(zero[-1], zero[-1]))
    path, pos = [leaves[index]], index
    for lvl in range(depth):
        sib = pos ^ 1
        path.append(level[sib] if sib < len(level) else zero[lvl])
        nxt = []
        for i in range(0, len(level), 2):
            nxt.append(H(level[i], level[i + 1] if i + 1 < len(level) else zero[lvl]))
        level, pos = nxt, pos >> 1
    root = level[0]

    B = Builder(curve)
    v_root = B.public(root)
    v_path = [B.secret(p) for p in path]
    v_index = B.secret(index)
    zero_h = B.secret(0)
    B.add_constraint(ql=1, xa=zero_h)            # the initial chaining value is the constant 0
    total = _mimc_in_circuit(B, zero_h, [v_path[0]])          # leafSum
    # api.ToBinary(index, depth): boolean bits that recompose to the index
    bits, acc = [], None
    hinted = B.hint(HINT_NBITS, [v_index], [(index >> i) & 1 for i in range(depth)]) if bits_from_hint else None
    for i in range(depth):
        # bits_from_hint: the bits are what gnark's NBits hint produces at solve time (same variables, same rows --
        # only who assigns them differs); else the caller supplies them with the witness
        b = hinted[i] if bits_from_hint else B.secret((index >> i) & 1)
        B.add_constraint(qm=1, ql=-1, xa=b, xb=b)             # b * b - b = 0
        bits.append(b)
        if acc is None:
            acc = b
        else:
            nxt = B.internal(B.values[acc] + (B.values[b] << i))
            B.add_constraint(ql=1, qr=1 << i, qo=-1, xa=acc, xb=b, xc=nxt)
            acc = nxt
    B.assert_is_equal(acc, v_index)
    for i in range(depth):
        b, sib = bits[i], v_path[i + 1]
        # d1 = Select(b, sibling, sum), d2 = Select(b, sum, sibling):  d1 = sum + b (sib - sum), d2 = sib + sum - d1
        diff = B.internal(B.values[sib] - B.values[total])
        B.add_constraint(ql=1, qr=-1, qo=-1, xa=sib, xb=total, xc=diff)
        bd = B.mul(b, diff)
        d1 = B.add(total, bd)
        s = B.add(sib, total)
        d2 = B.internal(B.values[s] - B.values[d1])
        B.add_constraint(ql=1, qr=-1, qo=-1, xa=s, xb=d1, xc=d2)
        total = _mimc_in_circuit(B, zero_h, [d1, d2])         # nodeSum
    B.assert_is_equal(total, v_root)
    return B, root
