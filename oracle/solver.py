"""Witness-solver oracle (TEST INFRASTRUCTURE, like the rest of oracle/): gnark's solving rule for a sparse R1CS,
restated in plain Python integers.  The reference reaches it through plonk.Prove (/root/reference/algoplonk.go:81-89:
frontend.NewWitness assigns the public and secret inputs, the prover's first step is spr.Solve, gnark v0.15.0
constraint/solver [UPSTREAM-RECALL]): rows are visited in order; a row  ql*a + qr*b + qm*a*b + qo*c + qk = 0  whose
wires are all assigned is an assertion, a row with exactly one unassigned wire defines it.  Nothing here runs on the
product path: it checks b2p_solver_solve (csrc/solver.cuh)."""
from typing import List, Sequence, Tuple


class Unsatisfied(Exception):
    pass


def solve(r: int, nb_public: int, nb_variables: int, constraints: Sequence[tuple], input_vars: Sequence[int],
          inputs: Sequence[int], hints: Sequence[tuple] = (), hint_fn=None, unchecked: Sequence[int] = ()
          ) -> Tuple[List[int], List[int]]:
    """Returns (values of every variable, level of every constraint: 0 = assertion, else 1 + deepest wire it reads).
    hints: (id, in_vars, out_vars) triples; hint_fn(id, input values, n_out) -> output values runs a hint the first time a
    constraint needs one of its outputs (gnark runs hint instructions in stream order, which comes to the same values).
    unchecked: constraint indexes left out of the final check (BSB22 rows the prover completes)."""
    val = [None] * nb_variables
    lvl = [0] * nb_variables
    for v, x in zip(input_vars, inputs):
        val[v] = x % r
    hint_of = {}
    for h, (_, _, outs) in enumerate(hints):
        for v in outs:
            hint_of[v] = h
    done = set()

    def need(v):
        h = hint_of.get(v)
        if h is None or h in done:
            return
        hid, ins, outs = hints[h]
        for i in ins:                       # an input may itself come out of a hint that has not run yet
            if val[i] is None:
                need(i)
        if any(val[i] is None for i in ins):
            raise ValueError(f"hint {h} needed before its inputs are assigned")
        got = hint_fn(hid, [val[i] for i in ins], len(outs))
        d = max([lvl[i] for i in ins], default=0) + 1
        for v2, x in zip(outs, got):
            val[v2] = x % r
            lvl[v2] = d
        done.add(h)

    levels = []
    for j, (ql, qr, qm, qo, qk, a, b, c) in enumerate(constraints):
        used = [(a, bool(ql or qm)), (b, bool(qr or qm)), (c, bool(qo))]
        for w, u in used:
            if u and val[w] is None:
                need(w)
        unknown = sorted({w for w, u in used if u and val[w] is None})
        depth = max([lvl[w] for w, u in used if u and val[w] is not None], default=0)
        if not unknown:
            levels.append(0)
            continue
        if len(unknown) > 1:
            raise ValueError(f"constraint {j} has two unassigned wires")
        u = unknown[0]
        if sum(1 for w, us in used if us and w == u) > 1:
            raise ValueError(f"constraint {j} uses its unassigned wire twice")
        if u == c and qo:
            num = ql * val[a] + qr * val[b] + qm * val[a] * val[b] + qk
            den = qo
        elif u == a:
            num = qr * val[b] + (qo * val[c] if qo else 0) + qk
            den = ql + qm * val[b]
        else:
            num = ql * val[a] + (qo * val[c] if qo else 0) + qk
            den = qr + qm * val[a]
        if den % r == 0:
            raise Unsatisfied(f"constraint {j} cannot determine its wire")
        val[u] = (-num) * pow(den, -1, r) % r
        lvl[u] = depth + 1
        levels.append(depth + 1)
    skip = set(unchecked)
    for j, (ql, qr, qm, qo, qk, a, b, c) in enumerate(constraints):
        for w in (a, b, c):
            if val[w] is None:
                need(w)
            if val[w] is None:
                raise ValueError(f"variable {w} is never assigned")
        if j not in skip and (ql * val[a] + qr * val[b] + qm * val[a] * val[b] + qo * val[c] + qk) % r:
            raise Unsatisfied(f"constraint #{nb_public + j} is not satisfied")
    return val, levels
