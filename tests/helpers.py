"""Shared test plumbing: builds the same circuit for the CUDA path and for the oracles,
and reloads the committed golden fixtures (tests/golden/, made by tools/gen_golden.py)."""
import functools
import json
import os
import random

from oracle import plonk_oracle as po
from algoplonk_b200 import frontend as fe

TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_trace(tc: fe.TraceColumns) -> po.Trace:
    return po.Trace(po.CURVES[tc.curve], tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm,
                    tc.qcp, tc.commitment_constraint_indexes)


def build_bsb22(curve: str, k: int, commit_fn, blind_seed: int = 7):
    """Builds bsb22Circuit (bsb22_test.go:18-39) with k commitments.  commit_fn(col) -> affine
    point commits to a Lagrange column (GPU or oracle).  Returns (cs, values, pi2 columns, commitments)."""
    cv = po.CURVES[curve]
    dry = fe.bsb22_circuit(curve, k, lambda rows, vals, crow: 1).build()
    n, off, nbc = dry.domain_size, dry.nb_public, dry.nb_constraints
    rng = random.Random(blind_seed)
    pi2s, coms = [], []

    def hint(rows, vals, crow):
        col = [0] * n
        for rr, v in zip(rows, vals):
            col[off + rr] = v
        col[off + crow] = rng.randrange(cv.r)          # gnark: SetRandom on the commitment row
        col[off + nbc - 1] = rng.randrange(cv.r)       # ... and on the last constraint row
        com = commit_fn(col)
        pi2s.append(col)
        coms.append(com)
        return po.hash_fr(cv, po.fs_point(cv, com))

    B = fe.bsb22_circuit(curve, k, hint)
    return B.build(), B.values, pi2s, coms


def vk_from_points(tc: fe.TraceColumns, pts, g1, tau=None, g2=None) -> po.VerifyingKey:
    """pts: S1 S2 S3 Ql Qr Qm Qo Qk Qcp* (affine ints)."""
    cv = po.CURVES[tc.curve]
    return po.VerifyingKey(curve=cv, size=tc.n, size_inv=pow(tc.n, -1, cv.r), omega=po.domain_generator(cv, tc.n),
                           nb_public=tc.nb_public, coset_shift=cv.coset_shift, S=list(pts[0:3]), Ql=pts[3],
                           Qr=pts[4], Qm=pts[5], Qo=pts[6], Qk=pts[7], Qcp=list(pts[8:]),
                           commitment_constraint_indexes=list(tc.commitment_constraint_indexes), g1=g1, tau=tau,
                           g2=g2)


# ---- golden fixtures ---------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def srs_kat():
    with open(os.path.join(GOLDEN, "srs_kat.json")) as f:
        return json.load(f)


@functools.lru_cache(maxsize=None)
def golden_proofs():
    with open(os.path.join(GOLDEN, "proofs.json")) as f:
        return json.load(f)


@functools.lru_cache(maxsize=None)
def real_srs_points(name: str):
    """Decompressed G1 points of the committed slice of a reference SRS file."""
    ent = srs_kat()[name]
    cv = po.CURVES[ent["curve"]]
    raw = bytes.fromhex(ent["first"])
    return [po.g1_decompress(cv, raw[i * cv.fp_bytes:(i + 1) * cv.fp_bytes]) for i in range(ent["count"])]


@functools.lru_cache(maxsize=None)
def real_srs_g2(name: str):
    """(G2[0], G2[1]) = ([1]_2, [tau]_2) of a reference setup, from its committed vk.bin bytes."""
    from oracle import pairing
    ent = srs_kat()[name]
    cv = po.CURVES[ent["curve"]]
    g2, g1 = pairing.parse_vk_bin(cv, bytes.fromhex(ent["vk_bin"]))
    assert g1 == cv.g1
    return g2


def case_id(case) -> str:
    return f"{case['curve']}-{case['name']}-{case['srs'][:4]}"


def build_case(case):
    """Rebuilds the prover inputs of a golden case.  Returns dict(cs, tc, L, R, O, pi2, coms, srs, tau)."""
    curve = case["curve"]
    cv = po.CURVES[curve]
    if case["name"] == "basic":
        B = fe.basic_circuit(curve)
        cs, values, pi2s, coms = B.build(), B.values, [], []
    elif case["name"].startswith("bsb22"):
        k = case["k"]
        n_dry = fe.bsb22_circuit(curve, k, lambda a, b, c: 1).build().domain_size
        srs_o = (po.srs_from_tau(cv, TAU, n_dry + 3) if case["srs"] == "tau"
                 else real_srs_points(case["srs"])[: n_dry + 3])
        trd = type("T", (), {"curve": cv, "n": n_dry})
        cs, values, pi2s, coms = build_bsb22(curve, k, lambda col: po.bsb22_commit(trd, srs_o, col))
        assert [po.g1_raw_bytes(cv, P).hex() for P in coms] == case["bsb22"]
    elif case["name"].startswith("squaring_2p"):
        cs, values = fe.squaring_chain(curve, int(case["name"].split("2p")[1]), x0=3)
        pi2s, coms = [], []
    else:
        raise KeyError(case["name"])
    tc = fe.build_trace(cs)
    assert tc.n == case["n"]
    L, R, O = fe.solve_lro(cs, values, tc.n)
    if case["srs"] == "tau":
        srs, tau = po.srs_from_tau(cv, TAU, tc.n + 3), TAU
    else:
        srs, tau = real_srs_points(case["srs"])[: tc.n + 3], None
    return dict(cs=cs, tc=tc, L=L, R=R, O=O, pi2=pi2s, coms=coms, srs=srs, tau=tau, cv=cv,
                blinding=case["blinding"])


# ---- scalar distributions -------------------------------------------------
def scalars_uniform(r: int, n: int, seed: int):
    rng = random.Random(seed)
    return [rng.randrange(r) for _ in range(n)]


def scalars_witness_like(r: int, n: int, seed: int):
    """40 % zero, 20 % one, 10 % < 2^16, 30 % uniform: the bucket-skew case."""
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        t = rng.random()
        if t < 0.4:
            out.append(0)
        elif t < 0.6:
            out.append(1)
        elif t < 0.7:
            out.append(rng.randrange(1 << 16))
        else:
            out.append(rng.randrange(r))
    return out
