"""The reference's examples/merkle (examples/merkle/logicsigVerifier/main.go:34-175) through this repo's mirror of
AlgoPlonk's API, end to end on the GPU -- everything main.go does around plonk.Prove except the Algorand side
(PuyaPy verifier generation, localnet simulation: out of scope):

    ap.Compile(&circuit, ecc.BN254, setup.PerpetualPowersOfTauBN254)      -> api.Compile on the real PPoT points
    compiledCircuit.Verify(&assignment)                                    -> inputs -> solver (NBits hint) -> prover -> plonk.Verify
    verifiedProof.ExportProofAndPublicInputs(proofFile, publicInputsFile)  -> same files, same bytes layout
    utils.SerializeCompiledCircuit / DeserializeCompiledCircuit            -> the key snapshot, and a second proof from it

    python examples/python/merkle.py [output folder]

Needs a B200 (the library has no CPU fallback).  The PPoT points are the committed slice of the reference's own
setup/PerpetualPowersOfTauBN254/pk.bin (tests/golden/ppot_bn254_first_131075.bin), its G2 points the setup's vk.bin."""
import json
import os
import random
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from algoplonk_b200 import api, frontend as fe      # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main(out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    curve = "BN254"
    # the MerkleCircuit of main.go:45-61 for the tree of main.go:63-92 (six leaves, proof for leaf 3, depth 16); the
    # bits of the leaf index come from gnark's NBits hint at solve time
    builder, root = fe.merkle_circuit(curve, depth=16, nb_leaves=6, index=3, bits_from_hint=True)
    cs = builder.build()
    assignment = [builder.values[v] for v in cs.input_vars]          # RootHash, Path[17], Index (+ the constant 0)

    print("Compiling circuit (trace on the host, proving key resident on the GPU)")
    with open(os.path.join(GOLDEN, "ppot_bn254_first_131075.bin"), "rb") as f:
        pk_bin = f.read()
    with open(os.path.join(GOLDEN, "srs_kat.json")) as f:
        vk_bin = bytes.fromhex(json.load(f)["PerpetualPowersOfTauBN254"]["vk_bin"])
    srs = api.SRS.from_pk_bin(curve, pk_bin, cs.domain_size + 3, vk_bin=vk_bin)      # setup/setup.go:113-114: n + 3 points
    cc = api.Compile(cs, curve, api.SetupName.PerpetualPowersOfTauBN254, srs=srs)

    print("Verifying: inputs -> witness solver -> plonk.Prove -> plonk.Verify, all in the library")
    solver = api.Solver(cs, cc.trace, hint_fn=api.std_hint_fn())
    rng = random.SystemRandom()
    blinding = [rng.randrange(api.R_MOD[curve]) for _ in range(9)]   # gnark draws these with fr.SetRandom
    verified = api.VerifyFromInputs(cc, solver, assignment, blinding)
    assert verified.Witness == [root]
    print("  solver:", {k: v for k, v in solver.info().items() if k in ("levels", "widest_level", "last_us")},
          "ran on", "device" if solver.info()["last_where"] == 2 else "a host thread")

    proof_file = os.path.join(out_dir, "MerkleVerifier.proof")
    public_file = os.path.join(out_dir, "MerkleVerifier.public_inputs")
    print(f"Writing proof to {proof_file} and public inputs to {public_file}")
    verified.ExportProofAndPublicInputs(proof_file, public_file)

    key_file = os.path.join(out_dir, "MerkleVerifier.b2pk")
    print(f"Persisting the proving key's circuit half to {key_file}, reloading it, proving again from the reloaded key")
    api.SerializeCompiledCircuit(cc, key_file)
    assert not api.ShouldRecompile(key_file, os.path.abspath(__file__))
    cc2 = api.DeserializeCompiledCircuit(key_file, cs, srs)
    again = api.VerifyFromInputs(cc2, solver, assignment, blinding)
    with open(proof_file, "rb") as f:
        assert f.read() == api.MarshalProof(again.Proof)            # same inputs, same blinding: same bytes

    print("Checking the exported files the way a third party would (plonk.Verify on the bytes), alone and in a batch")
    with open(proof_file, "rb") as f, open(public_file, "rb") as g:
        proof, public = f.read(), g.read()
    cc.VerifyProof(proof, public)
    cc.VerifyProofs([proof] * 8, [public] * 8, device=True)
    try:
        cc.VerifyProof(proof, bytes(31) + b"\x01")
        raise SystemExit("a proof for another root was accepted")
    except ValueError:
        pass
    solver.free(); cc2.free(); cc.free(); srs.free()
    print(f"ok: {len(proof)}-byte proof of a depth-16 MiMC Merkle path, root {root:#x}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "generated")
