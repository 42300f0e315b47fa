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
