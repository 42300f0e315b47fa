"""GPU parity: random-dense circuits (full-width selectors, irregular copy cycles) against the C++ oracle, the cyclic
SRS layout of the sharded MSM, and the sharded NTT's composition with it.  Written at the end of round 1 (when they
could not run on hardware any more -- hence a file of their own); green in the round-1 driver run and throughout
round 2.  Their CPU halves: `tests/test_oracle.py` (random-dense circuits), `tests/test_sharded.py` (cyclic gloo
collective), `tests/test_sharded_ntt.py` (host run of the kernel bodies).
"""
import random

import pytest

import helpers as H
from algoplonk_b200 import _lib, api, frontend as fe, sharded, sharded_ntt as sn
from oracle import cpu_oracle as co
from oracle import plonk_oracle as po
from test_sharded_ntt import to_tensor

pytestmark = pytest.mark.gpu
CURVES = ("BN254", "BLS12_381")
SETUP = {"BN254": api.SetupName.TestOnlyBN254, "BLS12_381": api.SetupName.TestOnlyBLS12381}


@pytest.mark.parametrize("curve,logn", [("BN254", 10), ("BN254", 13), ("BLS12_381", 12)])
def test_random_dense_circuit_byte_identical_to_cpp_oracle(gpu, curve, logn):
    """SURVEY 8d "random-dense" shape: full-width random values in every selector column (the squaring chain
    only has qm = 1, qo = -1), irregular copy cycles, three public inputs, past the single-tile NTT sizes."""
    cv = po.CURVES[curve]
    cs, values = fe.random_dense_circuit(curve, logn, seed=logn)
    cc = api.Compile(cs, curve, SETUP[curve])
    tc = cc.trace
    L, R, O = fe.solve_lro(cs, values, tc.n)
    blinding = H.scalars_uniform(cv.r, 9, logn)
    blob = api.MarshalProof(cc.Prove(L, R, O, blinding))
    srs_le = co.srs_from_tau_bytes(cv.cid, api.TEST_TAU, tc.n + 3)
    circ = co.Circuit(cv.cid, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), srs_le)
    assert cc.vk_commitments() == circ.vk_points()
    assert blob == circ.prove(L, R, O, blinding)
    vk = H.vk_from_points(tc, cc.vk_commitments(), cv.g1, tau=api.TEST_TAU)
    assert po.verify_proof(vk, blob, api.MarshalPublicInputs(curve, L[: tc.nb_public]))
    circ.free()
    cc.free()


@pytest.mark.parametrize("curve", CURVES)
def test_cyclic_shards_on_one_gpu_equal_whole_msm(gpu, curve):
    """layout="cyclic" (the coefficient distribution of the domain-sharded NTT): shards generated on the device
    with b2p_srs_generate_unsafe_strided and shards cut from host points hold the whole SRS's points r, r+G, ...
    and their partial sums add up to the whole MSM."""
    n, world = 1003, 4
    cv = po.CURVES[curve]
    rng = random.Random(9)
    scalars = [rng.randrange(cv.r) for _ in range(n)]
    whole = api.SRS.unsafe(curve, n)
    want = whole.msm(scalars)
    all_pts = api.points_to_mont_bytes(curve, whole.points(0, n))
    for make in ("unsafe", "from_points"):
        parts = []
        for r in range(world):
            sh = sharded.ShardedSRS.unsafe(curve, n, r, world, layout="cyclic") if make == "unsafe" else \
                sharded.ShardedSRS.from_points(curve, all_pts, r, world, layout="cyclic")
            idx = sharded.shard_indices(n, r, world, "cyclic")
            assert (sh.first, sh.count, sh.layout) == (r, len(idx), "cyclic")
            assert api.SRS(curve, sh.handle).points(0, 3) == [whole.points(i, 1)[0] for i in idx[:3]]
            assert api.SRS(curve, sh.handle).points(sh.count - 1, 1) == whole.points(idx[-1], 1)
            parts.append(sh.local_msm_raw(api.fr_to_mont_bytes(curve, [scalars[i] for i in idx])))
            sh.free()
        assert api.points_from_mont_bytes(curve, sharded.g1_sum(curve, b"".join(parts)))[0] == want
    whole.free()


@pytest.mark.parametrize("curve", CURVES)
def test_gpu_commit_lagrange_world_one(gpu, curve):
    """Sharded iNTT + cyclic-sharded MSM composed (sharded_ntt.commit_lagrange) == the single-GPU Lagrange-basis
    commitment b2p_msm_g1(B2P_BASIS_LAGRANGE) == the oracle's MSM over the Lagrange SRS, at world 1."""
    cv = po.CURVES[curve]
    n = 256
    vals = H.scalars_uniform(cv.r, n, 17)                     # evaluations in natural order
    whole = api.SRS.unsafe(curve, n + 3)
    want = whole.msm(vals, basis=_lib.BASIS_LAGRANGE)
    block = [vals[k] for k in sn.local_eval_exponents(n, 0, 1)]
    for mode in ("staged", "p2p"):
        nt = sn.ShardedNtt(curve, n, rank=0, world=1, mode=mode)
        srs = sharded.ShardedSRS.unsafe(curve, n + 3, 0, 1, layout="cyclic")
        got = sn.commit_lagrange(nt, srs, to_tensor(curve, block, "cuda"))
        assert api.points_from_mont_bytes(curve, got)[0] == want
        with pytest.raises(ValueError):
            sn.commit_lagrange(nt, sharded.ShardedSRS.unsafe(curve, n + 3, 0, 1), to_tensor(curve, block, "cuda"))
        nt.free()
        srs.free()
    # the oracle's view of the same commitment: coefficients by inverse NTT, then the canonical-basis sum
    coeffs = po.intt(cv, vals, po.domain_generator(cv, n))
    assert want == po.commit(cv, whole.points(0, n), coeffs)
    whole.free()


# ---- one proof over several GPUs: the commit hook (algoplonk_b200/sharded_prover.py) --------------------------
@pytest.mark.parametrize("curve", CURVES)
def test_whole_proof_through_the_commit_hook(gpu, curve):
    """b2p_prove with every kzg.Commit delegated through b2p_srs_set_commit_hook gives the plain prover's bytes:
    (1) ShardedProver at world 1 (one shard = the whole SRS, no process group); (2) a world of 3 simulated on one
    device -- three ragged SRS blocks, the committer's own slicing, partial sums added with b2p_g1_sum."""
    import torch
    from algoplonk_b200 import sharded_prover as sp
    cv = po.CURVES[curve]
    cs, values = fe.squaring_chain(curve, 10)
    plain = api.Compile(cs, curve, SETUP[curve])
    tc = plain.trace
    L, R, O = fe.solve_lro(cs, values, tc.n)
    blinding = H.scalars_uniform(cv.r, 9, 3)
    want = api.MarshalProof(plain.Prove(L, R, O, blinding))

    prover = sp.ShardedProver(cs, curve, SETUP[curve])
    assert api.MarshalProof(prover.Prove(L, R, O, blinding)) == want
    assert prover.committer.commits == 9                      # L R O Z H0 H1 H2 W_zeta W_{omega zeta}
    prover.close()

    world, total = 3, tc.n + 3
    shards = [sharded.ShardedSRS.unsafe(curve, total, r, world) for r in range(world)]
    seen = []

    def all_ranks(scalars, offset, count):                    # what ranks 0..2 would each do, one after the other
        assert (offset, count) == (0, len(scalars) // 32)     # world 1 from the committer's point of view
        n, parts = count, []
        for r in range(world):
            off, cnt = sp.slice_for(total, r, world, n)
            parts.append(shards[r].local_msm_dev_raw(scalars.data_ptr() + 32 * off, cnt) if cnt
                         else bytes(2 * cv.fp_bytes))
        seen.append(n)
        return sharded.g1_sum(curve, b"".join(parts))
    device = torch.device("cuda", torch.cuda.current_device())
    hook = sp.CommitHook(plain.srs, sp.ShardedCommitter(curve, total, device=device, local_msm=all_ranks), device)
    assert api.MarshalProof(plain.Prove(L, R, O, blinding)) == want
    assert seen == [tc.n + 2] * 3 + [tc.n + 3] + [tc.n + 2] * 5
    hook.remove()
    assert api.MarshalProof(plain.Prove(L, R, O, blinding)) == want      # and the handle commits by itself again
    for s in shards:
        s.free()
    plain.free()
