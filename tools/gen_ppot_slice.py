#!/usr/bin/env python3
"""Generates the BASELINE config-2 fixtures (run in the dev container: reads /root/reference once, the tests never do).

  tests/golden/ppot_bn254_first_131075.bin
        the first 2^17 + 3 compressed G1 points of /root/reference/setup/PerpetualPowersOfTauBN254/pk.bin, copied
        verbatim behind a fresh 4-byte big-endian count -- the bytes `setup.Run` reads for a 2^17-row circuit
        (setup/setup.go:85-90,113-114,196-228).  Reference-held data, test fixture only.
  tests/golden/config2_ppot_2p17.json
        the proof of the 2^17-row squaring chain (x0 = 3, blinding 1..9) over those points, written by the C++ CPU
        oracle (oracle/cpu_plonk.cpp); freezes the oracle's output so that a later change to oracle or CUDA path
        shows up as a diff against a committed value, and carries the sha256 of the .bin.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cpu_oracle as co  # noqa: E402
from algoplonk_b200 import frontend as fe  # noqa: E402

SRC = "/root/reference/setup/PerpetualPowersOfTauBN254/pk.bin"
OUT = os.path.join(ROOT, "tests", "golden")
LOG2, COUNT = 17, (1 << 17) + 3


def main():
    with open(SRC, "rb") as f:
        declared = int.from_bytes(f.read(4), "big")
        assert declared >= COUNT
        payload = f.read(COUNT * 32)
    blob = COUNT.to_bytes(4, "big") + payload
    with open(os.path.join(OUT, "ppot_bn254_first_131075.bin"), "wb") as f:
        f.write(blob)
    pts = co.g1_decompress_bytes(0, payload)
    cs, values = fe.squaring_chain("BN254", LOG2, x0=3)
    tc = fe.build_trace(cs)
    L, R, O = fe.solve_lro(cs, values, tc.n)
    circ = co.Circuit(0, tc.n, tc.nb_public, tc.ql, tc.qr, tc.qm, tc.qo, tc.qk, tc.perm, (), (), pts)
    blinding = list(range(1, 10))
    proof = circ.prove(L, R, O, blinding)
    vk = co.points_le(0, circ.vk_points())
    with open(os.path.join(OUT, "config2_ppot_2p17.json"), "w") as f:
        json.dump({"curve": "BN254", "log2": LOG2, "x0": 3, "blinding": blinding, "srs_points": COUNT,
                   "srs_sha256": hashlib.sha256(blob).hexdigest(), "vk_points_le": vk.hex(), "proof": proof.hex(),
                   "public_inputs": L[0].to_bytes(32, "big").hex(),
                   "generator": "tools/gen_ppot_slice.py (oracle/cpu_plonk.cpp)"}, f, indent=1)
    print("wrote", len(blob), "bytes and a", len(proof), "byte proof")


if __name__ == "__main__":
    main()
