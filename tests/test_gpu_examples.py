"""The example programs run as a user would run them (a subprocess, no test fixtures)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_merkle_example_of_the_reference(gpu, tmp_path):
    """examples/python/merkle.py = the reference's examples/merkle/logicsigVerifier/main.go minus the Algorand side."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "python", "merkle.py"), str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ok: 768-byte proof" in out.stdout
    assert os.path.getsize(tmp_path / "MerkleVerifier.proof") == 768            # bsb22_test.go:70: (24 + 3k) * 32, k = 0
    assert os.path.getsize(tmp_path / "MerkleVerifier.public_inputs") == 32
