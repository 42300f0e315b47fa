"""GPU parity: b2p_solver_* (spr.Solve inside plonk.Prove, /root/reference/algoplonk.go:81-89) against
the oracle's restatement of gnark's solving rule and against the witness the front end computed while building the
circuit -- bit-exact L, R, O on the device path, the host path and whatever B2P_SOLVE_AUTO picks."""
import ctypes as C

import pytest

import helpers as H
from algoplonk_b200 import _lib, api, frontend as fe
from oracle import plonk_oracle as po
from oracle import solver as osolver

pytestmark = pytest.mark.gpu
CURVES = ("BN254", "BLS12_381")
SETUP = {"BN254": api.SetupName.TestOnlyBN254, "BLS12_381": api.SetupName.TestOnlyBLS12381}
WHERE = (_lib.SOLVE_HOST, _lib.SOLVE_DEVICE, _lib.SOLVE_AUTO)


def _circuits(curve):
    B = fe.basic_circuit(curve)
    yield "basic", B.build(), B.values
    B, _ = fe.merkle_circuit(curve, depth=3)
    yield "merkle3", B.build(), B.values
    yield ("chain10",) + fe.squaring_chain(curve, 10, x0=7)
    yield ("dense10",) + fe.random_dense_circuit(curve, 10, seed=3)
    yield ("wide",) + fe.wide_mimc_circuit(curve, 1500, 5)            # levels wider than one block's share
    yield ("wide_narrow",) + fe.wide_mimc_circuit(curve, 40, 9)
    B = fe.Builder(curve)                                                # a division: the R wire is the unknown
    x = B.public(12345)
    y = B.secret(777)
    B.assert_is_different_from_zero(B.add(x, y))
    inv = B.internal(pow(B.values[y], -1, B.r) * 5 % B.r)                # 5 / y through the L wire: ql = 0, qm = 1
    B.add_constraint(qm=1, qk=-5, xa=inv, xb=y)
    yield "division", B.build(), B.values


@pytest.mark.parametrize("curve", CURVES)
def test_solver_matches_oracle_and_front_end(gpu, curve):
    cv = po.CURVES[curve]
    for name, cs, values in _circuits(curve):
        tc = fe.build_trace(cs)
        want = fe.solve_lro(cs, values, tc.n)
        inputs = [values[v] for v in cs.input_vars]
        oval, olevels = osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, inputs)
        assert oval == [v % cv.r for v in values], name                    # the oracle against the eager builder
        s = api.Solver(cs, tc)
        info = s.info()
        assert info["levels"] == max(olevels, default=0), name
        assert info["solved_rows"] == sum(1 for l in olevels if l), name
        for where in WHERE:
            got = s.solve(inputs, where)
            assert got == want, (name, where)
            assert fe.check_gates(tc, *got)
        assert s.info()["last_where"] in (_lib.SOLVE_HOST, _lib.SOLVE_DEVICE)
        s.free()


@pytest.mark.parametrize("curve", CURVES)
def test_solver_reports_what_gnark_reports(gpu, curve):
    cv = po.CURVES[curve]
    B = fe.basic_circuit(curve)
    cs = B.build()
    s = api.Solver(cs)
    for where in (_lib.SOLVE_HOST, _lib.SOLVE_DEVICE):
        with pytest.raises(_lib.B200PlonkError, match=r"constraint #\d+ is not satisfied") as e:
            s.solve([3, 4, 6], where)                                       # 9 + 16 != 36
        assert e.value.code == _lib.ERR_VERIFY
        assert s.solve([3, 4, cv.r - 5], where) == fe.solve_lro(cs, fe.basic_circuit(curve, 3, 4, -5).values, 8)
    with pytest.raises(ValueError):
        s.solve([3, 4])
    s.free()
    # x != 0 with x = 0: the row cannot determine the inverse
    B = fe.Builder(curve)
    x = B.public(0)
    B.assert_is_different_from_zero(x)
    s = api.Solver(B.build())
    for where in (_lib.SOLVE_HOST, _lib.SOLVE_DEVICE):
        with pytest.raises(_lib.B200PlonkError, match="division by zero"):
            s.solve([0], where)
        assert s.solve([2], where)[1][1] == pow(2, -1, cv.r)
    s.free()


def test_solver_create_refuses_what_needs_a_hint(gpu):
    curve = "BN254"
    B = fe.Builder(curve)
    x = B.public(3)
    a, b = B.internal(1), B.internal(2)
    B.add_constraint(ql=1, qr=1, qo=-1, xa=a, xb=b, xc=x)                  # two unassigned wires
    with pytest.raises(_lib.B200PlonkError, match="two unassigned wires"):
        api.Solver(B.build())
    B = fe.Builder(curve)
    x = B.public(9)
    a = B.internal(3)
    B.add_constraint(qm=1, qo=-1, xa=a, xb=a, xc=x)                        # a * a = x: a square root, not linear
    with pytest.raises(_lib.B200PlonkError, match="twice"):
        api.Solver(B.build())
    B = fe.Builder(curve)
    x = B.public(9)
    a = B.internal(3)
    B.add_constraint(ql=1, xa=x, xb=0, xc=a)                               # `a` sits on a wire with a zero selector
    with pytest.raises(_lib.B200PlonkError, match="neither an input nor determined"):
        api.Solver(B.build())
    cs, _, _, _ = H.build_bsb22(curve, 1, lambda col: po.CURVES[curve].g1)
    with pytest.raises(ValueError, match="hint"):
        api.Solver(cs)
    lib = _lib.load()
    out = C.c_void_p()
    assert lib.b2p_solver_create(0, 8, 1, 4, None, 0, None, None, None, None, None, None, None, None, C.byref(out)) == _lib.ERR_ARG
    assert lib.b2p_solver_solve(None, None, 0, None, None, None) == _lib.ERR_ARG


@pytest.mark.parametrize("curve,build", [("BN254", lambda c: fe.squaring_chain(c, 10, x0=3)),
                                         ("BLS12_381", lambda c: fe.wide_mimc_circuit(c, 255, 4)),
                                         ("BN254", lambda c: fe.wide_mimc_circuit(c, 4095, 4))])
def test_inputs_to_verified_proof_on_the_device(gpu, curve, build):
    """cc.Verify as the reference runs it: inputs in, verified proof out; L, R, O go from the solver to the prover
    inside HBM.  Same bytes as proving the front end's own witness through host buffers."""
    cv = po.CURVES[curve]
    cs, values = build(curve)
    srs = api.SRS.unsafe(curve, cs.domain_size + 3, H.TAU)
    cc = api.Compile(cs, curve, SETUP[curve], srs=srs)
    blinding = H.scalars_uniform(cv.r, 9, 5)
    want = api.MarshalProof(cc.Prove(*fe.solve_lro(cs, values, cc.trace.n), blinding))
    s = api.Solver(cs, cc.trace)
    inputs = [values[v] for v in cs.input_vars]
    for where in WHERE:
        vp = api.VerifyFromInputs(cc, s, inputs, blinding, where)            # b2p_verify runs inside
        assert api.MarshalProof(vp.Proof) == want
        assert vp.Witness == [values[0] % cv.r]
    bad = list(inputs)
    bad[1] = (bad[1] + 1) % cv.r
    with pytest.raises(_lib.B200PlonkError, match="not satisfied"):
        api.VerifyFromInputs(cc, s, bad, blinding)
    s.free()
    cc.free()
    srs.free()


# ---- hints -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("curve", CURVES)
def test_solver_runs_the_callers_hints(gpu, curve):
    """gnark's hint functions are the caller's: the NBits hint behind api.ToBinary (the reference's Merkle circuit takes
    the bits of its leaf index that way, examples/merkle/logicsigVerifier/main.go:45-61) placed where its input is
    known, on the device path (a synchronisation point) and on the host path; same witness as the oracle's solver."""
    cv = po.CURVES[curve]
    cases = []
    B = fe.Builder(curve)
    x = B.public(0b1011001)
    y = B.secret(77)
    bits = B.to_binary(B.mul(x, y), 14)                       # the hint's input is itself solved from a row
    B.assert_is_equal(B.add(bits[0], bits[3]), B.add(bits[3], bits[0]))
    cases.append(("to_binary", B.build(), B.values))
    M, _ = fe.merkle_circuit(curve, depth=3, bits_from_hint=True)
    cases.append(("merkle_hinted", M.build(), M.values))
    W = fe.Builder(curve)                                      # hints between wide levels: the launch plan is cut there
    xs = [W.secret(1000 + i) for i in range(600)]
    sq = [W.mul(v, v) for v in xs]
    lows = [W.to_binary(W.hint(fe.HINT_NBITS + 0, [q], [W.values[q] & 1])[0], 1)[0] for q in sq[:3]]
    W.assert_is_equal(W.add(lows[0], lows[1]), W.add(lows[1], lows[0]))
    cases.append(("wide_with_hints", W.build(), W.values))
    for name, cs, values in cases:
        tc = fe.build_trace(cs)
        want = fe.solve_lro(cs, values, tc.n)
        inputs = [values[v] for v in cs.input_vars]
        hs = [(h.id, h.in_vars, h.out_vars) for h in cs.hints]
        oval, olevels = osolver.solve(cv.r, cs.nb_public, cs.nb_variables, cs.constraints, cs.input_vars, inputs, hs,
                                      api.std_hint_fn())
        assert oval == [v % cv.r for v in values], name
        s = api.Solver(cs, tc, hint_fn=api.std_hint_fn())
        for where in WHERE:
            assert s.solve(inputs, where) == want, (name, where)
        s.free()
        with pytest.raises(ValueError, match="hint"):
            api.Solver(cs, tc)                                  # hints recorded, no function given
    # a hint function that fails, and one that lies (the rows that constrain its outputs catch it)
    cs, values = cases[0][1], cases[0][2]
    inputs = [values[v] for v in cs.input_vars]

    def failing(hid, vals, n_out):
        raise RuntimeError("no such hint")
    s = api.Solver(cs, hint_fn=failing)
    for where in (_lib.SOLVE_HOST, _lib.SOLVE_DEVICE):
        with pytest.raises(_lib.B200PlonkError, match="hint function reported a failure") as e:
            s.solve(inputs, where)
        assert e.value.code == _lib.ERR_INTERNAL
    s.free()
    s = api.Solver(cs, hint_fn=lambda hid, vals, n_out: [1] * n_out)
    for where in (_lib.SOLVE_HOST, _lib.SOLVE_DEVICE):
        with pytest.raises(_lib.B200PlonkError, match="is not satisfied"):
            s.solve(inputs, where)
    s.free()


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("k", (1, 2))
def test_solver_with_the_bsb22_commitment_hint(gpu, curve, k):
    """bsb22Circuit (bsb22_test.go:18-39): the commitment hint -- commit to the committed wires on the Lagrange basis
    (b2p_msm_g1), hash to the field -- runs as a solver hint; the committed rows and the commitment row, whose gates the
    prover completes, are left out of the solver's check.  L, R, O then prove to the golden bytes."""
    cv = po.CURVES[curve]
    n_dry = fe.bsb22_circuit(curve, k, lambda a, b, c: 1).build().domain_size
    srs = api.SRS.unsafe(curve, n_dry + 3, H.TAU)
    cs, values, pi2s, coms = H.build_bsb22(curve, k, lambda col: srs.msm(col, basis=_lib.BASIS_LAGRANGE))
    case = next(c for c in H.golden_proofs() if c["curve"] == curve and c["name"] == f"bsb22_k{k}" and c["srs"] == "tau")
    off = cs.nb_public
    calls = []

    def commitment_hint(hid, vals, n_out):
        c = hid - fe.HINT_BSB22
        col = pi2s[c]                                            # the committed values + gnark's two blinding slots
        assert [col[off + r] for r in cs.commitments[c].committed_rows] == vals
        com = srs.msm(col, basis=_lib.BASIS_LAGRANGE)            # on the GPU, inside the solve
        assert com == coms[c]
        calls.append(c)
        return [po.hash_fr(cv, po.fs_point(cv, com))]

    s = api.Solver(cs, hint_fn=api.std_hint_fn(commitment_hint))
    tc = fe.build_trace(cs)
    want = fe.solve_lro(cs, values, tc.n)
    inputs = [values[v] for v in cs.input_vars]
    for where in WHERE:
        assert s.solve(inputs, where) == want, where
    assert calls == list(range(k)) * 3
    L, R, O = s.solve(inputs)
    cc = api.Compile(cs, curve, SETUP[curve], srs=srs)
    assert api.MarshalProof(cc.Prove(L, R, O, case["blinding"], pi2s, coms)).hex() == case["proof"]
    # a wrong input: this test's hint notices that the committed values changed ...
    with pytest.raises(_lib.B200PlonkError, match="hint function reported a failure"):
        s.solve([inputs[0] + 1] + inputs[1:])
    s.free()
    # ... and with a hint that does not look, the rows the solver does check (X == Y * Y) fail
    s = api.Solver(cs, hint_fn=api.std_hint_fn(lambda hid, vals, n_out: [po.hash_fr(cv, po.fs_point(cv, coms[hid - fe.HINT_BSB22]))]))
    for where in (_lib.SOLVE_HOST, _lib.SOLVE_DEVICE):
        with pytest.raises(_lib.B200PlonkError, match="is not satisfied"):
            s.solve([inputs[0] + 1] + inputs[1:], where)
    s.free()
    cc.free()
    srs.free()

