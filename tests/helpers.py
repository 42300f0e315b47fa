"""Shared test plumbing: builds the same circuit for the CUDA path and for the oracle."""
import random

from oracle import plonk_oracle as po
from algoplonk_b200 import frontend as fe

TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF


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
