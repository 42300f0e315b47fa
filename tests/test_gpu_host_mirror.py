"""The C++ mirror of AlgoPlonk's API (algoplonk_b200/host/algoplonk.hpp: Compile / Verify / MarshalProof, the
surface of /root/reference/algoplonk.go:37-131 and helper.go:13-110) driven from a compiled program with no
Python in the loop: examples/cpp/basic.cpp proves the reference's BasicCircuit and must print the committed
golden proof byte for byte."""
import os
import subprocess

import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build(tmp_path_factory, name="basic"):
    out = tmp_path_factory.mktemp("cpp") / name
    lib_dir = os.path.join(ROOT, "algoplonk_b200")
    subprocess.run([GXX, "-O2", "-std=c++17", "-o", str(out), os.path.join(ROOT, "examples", "cpp", name + ".cpp"),
                    "-L" + lib_dir, "-lb200plonk", "-Wl,-rpath," + lib_dir], check=True)
    return str(out)


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path_factory):
    """CPU: the header builds against the C ABI; without a device the program reports the library's error."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    exe = _build(tmp_path_factory)
    out = subprocess.run([exe, "BN254", "12"] + ["1"] * 9, capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "b2p_init" in out.stderr and out.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_cpp_mirror_reproduces_golden_basic_proof(gpu, tmp_path_factory, curve):
    case = next(c for c in H.golden_proofs() if c["curve"] == curve and c["name"] == "basic" and c["srs"] == "tau")
    exe = _build(tmp_path_factory)
    args = [exe, curve, format(H.TAU, "x")] + [format(b, "x") for b in case["blinding"]]
    out = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    proof_hex, public_hex = out.stdout.split()
    assert proof_hex == case["proof"]
    assert public_hex == case["public_inputs"]


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_cpp_caller_solver_and_persisted_key(gpu, tmp_path_factory, tmp_path, curve):
    """examples/cpp/bench_prove.cpp: a compiled caller proves the squaring chain from pageable columns, from the
    circuit's inputs alone (library solver, L R O in HBM) and from a reloaded key snapshot -- one proof, three ways --
    and a wrong input is refused with gnark's message."""
    import json
    exe = _build(tmp_path_factory, "bench_prove")
    out = subprocess.run([exe, curve, "12", "3", str(tmp_path / "key.b2pk")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout)
    assert line["same_bytes"] and line["bad_input_rejected"] and line["log2"] == 12
    assert os.path.getsize(tmp_path / "key.b2pk") == 64 + 5 * 32 * 4096 + 3 * 8 * 4096

