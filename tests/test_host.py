"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, host-only entry points (marshalling) match the reference layout, the field
templates' host emulation matches Python integers, and the front-end / API mirror behaves
like the reference (compile_test.go, setup/registry_test.go)."""
import ctypes as C
import os
import random
import re
import subprocess

import pytest

import helpers as H
from algoplonk_b200 import _lib, api, frontend as fe
from oracle import plonk_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---- C ABI ------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "b200plonk.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(b2p_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, f"header/binding mismatch: {declared ^ bound}"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.b2p_version()


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (cgo parses it with a C compiler) and a C caller
    must link against the library with nothing but the header."""
    hdr = os.path.join(ROOT, "include", "b200plonk.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                   check=True)
    src = tmp_path / "caller.c"
    src.write_text('#include "b200plonk.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%s %d\\n", b2p_version(), (int)b2p_proof_marshal_size(B2P_BN254, 0)); '
                   'return b2p_ntt(7, (void*)0, 1, 0) == B2P_ERR_ARG ? 0 : 1; }\n')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libb200plonk.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "sm_100a" in out.stdout and out.stdout.split()[-1] == "768"


def _call_arities(src: str, prefix: str):
    """(name, number of top-level arguments) of every `prefix name(...)` call in a source text."""
    out = []
    for m in re.finditer(re.escape(prefix) + r"(b2p_[a-z0-9_]+)\s*\(", src):
        depth, args, i, seen = 1, 0, m.end(), False
        while depth:
            ch = src[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                args += 1
            if depth and not ch.isspace():
                seen = True
            i += 1
        if src[m.end():i - 1].strip() == "void":
            seen = False
        out.append((m.group(1), args + 1 if seen else 0))
    return out


def test_go_shim_calls_match_the_header():
    """The cgo shim cannot be compiled here (no Go toolchain): at least every C.b2p_* call in go/gpuplonk must
    name a function the header declares, with the declared number of arguments."""
    with open(os.path.join(ROOT, "include", "b200plonk.h")) as f:
        header = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = {}
    for name, n in _call_arities(header, ""):
        declared[name] = n
    assert declared["b2p_prove"] == 8 and declared["b2p_last_error"] == 0 and declared["b2p_circuit_load"] == 15
    godir = os.path.join(ROOT, "go", "gpuplonk")
    calls = []
    for fn in sorted(os.listdir(godir)):
        if fn.endswith(".go"):
            with open(os.path.join(godir, fn)) as f:
                calls += [(fn, name, n) for name, n in _call_arities(f.read(), "C.")]
    assert len(calls) >= 15
    for fn, name, n in calls:
        assert name in declared, f"{fn}: C.{name} is not declared in b200plonk.h"
        assert declared[name] == n, f"{fn}: C.{name} called with {n} arguments, header declares {declared[name]}"


def test_sizes():
    lib = _lib.load()
    for k in (0, 1, 2):
        assert lib.b2p_proof_marshal_size(0, k) == (24 + 3 * k) * 32      # templateLogicSigBN254.go:50
        assert lib.b2p_proof_marshal_size(1, k) == (33 + 4 * k) * 32      # templateLogicSigBLS12_381.go:50
        assert lib.b2p_proof_raw_size(0, k) == 9 * 64 + (7 + k) * 32
        assert lib.b2p_proof_raw_size(1, k) == 9 * 96 + (7 + k) * 32


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU error path")
def test_init_fails_loudly_without_gpu():
    with pytest.raises(_lib.B200PlonkError) as e:
        _lib.init()
    assert e.value.code == -2 and "CUDA" in str(e.value)
    # ... and so does the public API: there is no CPU fallback
    with pytest.raises(_lib.B200PlonkError):
        api.SRS.unsafe("BN254", 8)


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    assert lib.b2p_ntt(7, C.create_string_buffer(32), 1, 0) == -1
    assert b"unsupported curve" in lib.b2p_last_error()
    assert lib.b2p_marshal_proof(0, 0, None, None, None) == -1
    assert lib.b2p_srs_size(None) == 0
    assert lib.b2p_srs_set_commit_hook(None, None, None) == -1 and b"null" in lib.b2p_last_error()
    assert lib.b2p_kzg_vk_load(0, None, 0, None, None) == -1
    assert lib.b2p_verify_batch(5, 8, 0, 0, None, None, None, None, None, 0, None, 0, 0, None) == -1
    assert lib.b2p_g2_generate_unsafe(0, None, None) == -1
    assert lib.b2p_device_copy(None, None, 16) == -1 and lib.b2p_device_copy(None, None, 0) == 0


@pytest.mark.parametrize("case", H.golden_proofs(), ids=H.case_id)
def test_marshal_proof_matches_reference_layout(case):
    """b2p_marshal_proof is host-only: raw gnark-layout proof -> helper.go:27-88 byte layout."""
    curve, k = case["curve"], case["k"]
    cv = po.CURVES[curve]
    blob = bytes.fromhex(case["proof"])
    pb = 2 * cv.fp_bytes
    # parse the golden blob (Appendix B) back into points / scalars
    pos = 0

    def P():
        nonlocal pos
        pt = po.g1_from_raw_bytes(cv, blob[pos:pos + pb])
        pos += pb
        return pt

    def S():
        nonlocal pos
        v = int.from_bytes(blob[pos:pos + 32], "big")
        pos += 32
        return v

    lro = [P(), P(), P()]
    h = [P(), P(), P()]
    claimed = [S() for _ in range(5)]
    z = P()
    zs = S()
    wz, wzw = P(), P()
    qcp = [S() for _ in range(k)]
    bsb = [P() for _ in range(k)]
    raw = api.points_to_mont_bytes(curve, lro + [z] + h + [wz, wzw]) + \
        api.fr_to_mont_bytes(curve, [12345] + claimed + qcp + [zs])
    out = api.MarshalProof(api.Proof(curve, k, raw, api.points_to_mont_bytes(curve, bsb)))
    assert out == blob


def test_marshal_public_inputs():
    vals = [0, 1, 35, po.BN254.r - 1]
    assert api.MarshalPublicInputs("BN254", vals) == po.marshal_public_inputs(vals)   # helper.go:96-109


# ---- field templates (host emulation of the PTX carry chains) ---------------------------------
@pytest.fixture(scope="module", params=["emulated-device-code", "host-u128"])
def hostfield(request, tmp_path_factory):
    """field.cuh compiled for the host twice: with the device multiplication code running on the carry-flag
    emulation of ptx.cuh (-DB2P_HOST_EMULATE_DEVICE_MUL), and with the 64-bit-limb host product the library's
    host side uses."""
    out = tmp_path_factory.mktemp("hf") / "hostfield.so"
    src = os.path.join(ROOT, "tests", "csrc", "hostfield_shim.cpp")
    flags = ["-DB2P_HOST_EMULATE_DEVICE_MUL"] if request.param == "emulated-device-code" else []
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O1", "-std=c++17", "-shared",
                    "-fPIC", *flags, "-x", "c++", src, "-o", str(out)], check=True)
    return C.CDLL(str(out))


@pytest.mark.parametrize("fid,mod,nlimb", [(0, po.BN254.r, 8), (1, po.BN254.p, 8), (2, po.BLS12_381.r, 8),
                                          (3, po.BLS12_381.p, 12)])
def test_device_field_code_on_host(hostfield, fid, mod, nlimb):
    """The very field.cuh the kernels use, run through its host emulation, against Python ints."""
    R = 1 << (32 * nlimb)
    rng = random.Random(fid)

    def call(op, a, b=0):
        A = (C.c_uint32 * nlimb)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(nlimb)])
        B = (C.c_uint32 * nlimb)(*[(b >> (32 * i)) & 0xFFFFFFFF for i in range(nlimb)])
        O = (C.c_uint32 * nlimb)()
        hostfield.ht_field_op(fid, op, A, B, O)
        return sum(int(O[i]) << (32 * i) for i in range(nlimb))

    Rinv = pow(R, -1, mod)
    vals = [0, 1, mod - 1, mod - 2, (1 << 32) - 1, 1 << 32] + [rng.randrange(mod) for _ in range(60)]
    for a in vals:
        b = rng.choice(vals)
        assert call(0, a, b) == a * b * Rinv % mod          # Montgomery product
        assert call(1, a, b) == (a + b) % mod
        assert call(2, a, b) == (a - b) % mod
        assert call(3, a) == (-a) % mod
        assert call(7, a) == a * a * Rinv % mod
        if mod.bit_length() + 2 <= 32 * nlimb:               # base fields: a*b - c*d with one reduction
            assert call(9, a, b) == (-b) % mod                 # x*y - y*(x + R) = -y*R, reduced once
            assert call(10, a, b) == (a * a - b * b) * Rinv % mod
        assert call(5, a) == a * R % mod
        assert call(6, a) == a * Rinv % mod
    # the separated product / reduction paths (squaring, a*b - c*d) on many more operands
    for _ in range(1500):
        a, b = rng.randrange(mod), rng.randrange(mod)
        assert call(7, a) == a * a * Rinv % mod
        if mod.bit_length() + 2 <= 32 * nlimb:
            assert call(10, a, b) == (a * a - b * b) * Rinv % mod
    # carry-propagation corner cases: operands built from saturated / empty limbs make limbs of the partial
    # sums hit 0xffffffff, where a dropped carry shows (a 2^-32 event per row on random operands: an earlier
    # mul_sub that added its second product after the first lost one carry per ~10^9 additions)
    structured_c = (mod - 1) // 3

    def structured():
        limbs = [rng.choice((0, 1, 0xFFFFFFFF, 0xFFFFFFFE, 0x80000000, rng.randrange(1 << 32))) for _ in range(nlimb)]
        return sum(v << (32 * i) for i, v in enumerate(limbs)) % mod
    for _ in range(3000):
        a, b = structured(), structured()
        assert call(0, a, b) == a * b * Rinv % mod
        assert call(7, a) == a * a * Rinv % mod
        if mod.bit_length() + 2 <= 32 * nlimb:
            assert call(10, a, b) == (a * a - b * b) * Rinv % mod
            assert call(11, a, b) == (a * b - b * structured_c) * Rinv % mod
    # hashes are reduced mod r from raw 256-bit values (Fiat-Shamir challenges): any N-limb input
    for _ in range(300):
        a = rng.randrange(R)
        assert call(8, a) == a * R % mod
    assert call(8, R - 1) == (R - 1) * R % mod
    for a in vals[:12]:
        am = a * R % mod
        inv = call(4, am)
        assert inv == (pow(a, -1, mod) * R % mod if a else 0)


# ---- front-end ---------------------------------------------------------------------------------
@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_frontend_basic_and_permutation(curve):
    B = fe.basic_circuit(curve)
    cs, values = B.build(), B.values
    tc = fe.build_trace(cs)
    assert tc.n == 8 and tc.nb_public == 2
    L, R, O = fe.solve_lro(cs, values, tc.n)
    assert fe.check_gates(tc, L, R, O)
    assert sorted(tc.perm) == list(range(3 * tc.n))
    # positions on one cycle carry one value
    wires = L + R + O
    assert all(wires[i] == wires[tc.perm[i]] for i in range(3 * tc.n))


def test_frontend_squaring_chain_sizes():
    for lg in (3, 6, 10):
        cs, values = fe.squaring_chain("BN254", lg)
        assert cs.nb_public + cs.nb_constraints == 1 << lg == cs.domain_size
        tc = fe.build_trace(cs)
        L, R, O = fe.solve_lro(cs, values, tc.n)
        assert fe.check_gates(tc, L, R, O)


def test_compile_rejects_like_the_reference():
    B = fe.basic_circuit("BN254")
    cs = B.build()
    with pytest.raises(ValueError, match="unknown setup"):          # compile_test.go:22-30
        api.Compile(cs, api.BN254, 999)
    with pytest.raises(ValueError, match="unsupported curve"):      # algoplonk.go:39-41
        api.Compile(cs, "BW6_761", api.SetupName.TestOnlyBN254)
    with pytest.raises(ValueError, match="do not match"):           # algoplonk.go:46-49
        api.Compile(cs, api.BN254, api.SetupName.DuskBLS12381)


def test_conversions_roundtrip():
    for curve in ("BN254", "BLS12_381"):
        cv = po.CURVES[curve]
        vals = H.scalars_uniform(cv.r, 5, 1)
        assert api.fr_from_mont_bytes(curve, api.fr_to_mont_bytes(curve, vals)) == vals
        pts = po.srs_from_tau(cv, 5, 3) + [None]
        assert api.points_from_mont_bytes(curve, api.points_to_mont_bytes(curve, pts)) == pts


@pytest.mark.parametrize("curve", ["BN254", "BLS12_381"])
def test_device_curve_formulas_on_host(hostfield, curve):
    """ec.cuh -- the XYZZ mixed addition, doubling and general addition the MSM kernels execute -- through the same
    host emulation, against the big-int oracle, including the cases a bucket sees rarely: P + P, P - P, infinity
    on either side, small multiples."""
    from algoplonk_b200 import api
    cv = po.CURVES[curve]
    cid = 0 if curve == "BN254" else 1
    nw = 2 * (8 if curve == "BN254" else 12)
    rng = random.Random(41 + cid)

    def call(op, P, Q=None, k=0):
        a = (C.c_uint32 * nw).from_buffer_copy(api.points_to_mont_bytes(curve, [P]))
        b = (C.c_uint32 * nw).from_buffer_copy(api.points_to_mont_bytes(curve, [Q]))
        o = (C.c_uint32 * nw)()
        hostfield.ht_ec_op(cid, op, a, b, C.c_uint64(k), o)
        return api.points_from_mont_bytes(curve, bytes(o))[0]

    pts = [po.g1_mul(cv, cv.g1, rng.randrange(1, cv.r)) for _ in range(6)] + [cv.g1]
    for P in pts:
        Q = rng.choice(pts)
        want = po.g1_add(cv, P, Q)
        assert call(0, P, Q) == want
        assert call(1, P, Q) == want                       # includes P == Q for some draws: the doubling branch
        assert call(4, P, Q) == want
        assert call(3, P, Q) == po.g1_add(cv, P, po.g1_neg(cv, Q))
        assert call(2, P) == po.g1_add(cv, P, P) == call(6, P)
        assert call(7, P) is None
        assert call(8, P, Q) == po.g1_add(cv, po.g1_mul(cv, P, 3), po.g1_mul(cv, Q, 2))
        for k in (0, 1, 2, 3, 7, 255, 65537):
            assert call(5, P, None, k) == (po.g1_mul(cv, P, k) if k else None)
    P = pts[0]
    assert call(1, P, P) == po.g1_add(cv, P, P) == call(0, P, P) == call(4, P, P)
    assert call(1, P, po.g1_neg(cv, P)) is None and call(0, P, po.g1_neg(cv, P)) is None and call(3, P, P) is None
    assert call(0, None, P) == P and call(0, P, None) == P and call(4, None, P) == P and call(4, P, None) == P
    assert call(1, None, P) == P and call(3, None, P) == po.g1_neg(cv, P)
    assert call(2, None) is None and call(0, None, None) is None and call(5, None, None, 5) is None


def test_msm_signed_digits_recompose_the_scalar(tmp_path_factory):
    """msm_digits.cuh on the host: for every window size the plan can choose and scalars built to stress the carry
    chain (all windows at / just above half, all ones, r - 1, powers of two around window edges), the signed digits
    sum back to the scalar, every bucket index is inside [0, 2^(c-1)), and the last window absorbs the final carry."""
    out = tmp_path_factory.mktemp("md") / "msm_digits.so"
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O1", "-std=c++17", "-shared", "-fPIC",
                    "-x", "c++", os.path.join(ROOT, "tests", "csrc", "msm_digits_shim.cpp"), "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    rng = random.Random(77)
    for curve, bits in (("BN254", 254), ("BLS12_381", 255)):
        r = po.CURVES[curve].r
        assert r.bit_length() == bits
        for c in range(2, 23):
            cc, W, nb = C.c_int(), C.c_int(), C.c_uint32()
            lib.ht_msm_plan(C.c_uint64(1 << 20), bits, c, C.byref(cc), C.byref(W), C.byref(nb))
            assert (cc.value, nb.value) == (c, 1 << (c - 1)) and W.value * c >= bits + 1 > (W.value - 1) * c
            half = 1 << (c - 1)
            pattern = lambda d: sum(d << (c * w) for w in range(W.value)) % r
            scalars = [0, 1, r - 1, r - 2, (1 << (bits - 1)) - 1, (1 << (bits - 1)), pattern(half), pattern(half + 1),
                       pattern(half - 1), pattern((1 << c) - 1), (1 << c) - 1, 1 << c, (1 << (c * (W.value - 1))) - 1]
            scalars += [rng.randrange(r) for _ in range(6)]
            for s in scalars:
                words = (C.c_uint32 * 8)(*[(s >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
                win, bucket, neg = (C.c_int * 64)(), (C.c_uint32 * 64)(), (C.c_int * 64)()
                assert W.value <= 64 or c < 4
                if W.value > 64:
                    win, bucket, neg = (C.c_int * 128)(), (C.c_uint32 * 128)(), (C.c_int * 128)()
                cnt = lib.ht_msm_digits(words, c, W.value, win, bucket, neg)
                assert 0 <= cnt <= W.value
                total = 0
                for i in range(cnt):
                    assert 0 <= bucket[i] < half and 0 <= win[i] < W.value
                    assert i == 0 or win[i] > win[i - 1]
                    total += (-1 if neg[i] else 1) * (bucket[i] + 1) << (c * win[i])
                assert total == s, (curve, c, hex(s))
    # the plan picks the window the cost model says: 2^20 points -> c = 20, 13 windows for BN254
    cc, W, nb = C.c_int(), C.c_int(), C.c_uint32()
    lib.ht_msm_plan(C.c_uint64((1 << 20) + 3), 254, 0, C.byref(cc), C.byref(W), C.byref(nb))
    assert (cc.value, W.value, nb.value) == (20, 13, 1 << 19)


def test_transcript_sha256_against_hashlib(tmp_path_factory):
    """csrc/sha256.hpp (every Fiat-Shamir challenge of prover and verifier goes through it) against hashlib at the
    padding boundaries (55, 56, 63, 64, 119, 120 bytes ...), long messages, and with the message split across two
    update() calls at every kind of offset."""
    import hashlib
    out = tmp_path_factory.mktemp("sha") / "sha256.so"
    subprocess.run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O1", "-std=c++17", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "csrc", "sha256_shim.cpp"), "-o", str(out)], check=True)
    lib = C.CDLL(str(out))
    rng = random.Random(256)
    lengths = list(range(0, 70)) + [111, 112, 119, 120, 127, 128, 129, 191, 192, 1000, 4096, 65537]
    for n in lengths:
        msg = bytes(rng.randrange(256) for _ in range(n))
        want = hashlib.sha256(msg).digest()
        for split in sorted({0, 1, n // 2, max(n - 1, 0), n, min(n, 55), min(n, 56), min(n, 64), min(n, 65)}):
            got = C.create_string_buffer(32)
            lib.ht_sha256(msg, C.c_uint64(n), C.c_uint64(split), got)
            assert got.raw == want, (n, split)



@pytest.mark.parametrize("curve", ("BN254", "BLS12_381"))
def test_fp2_device_code_on_host(hostfield, curve):
    """fp2.cuh (Fp[u]/(u^2+1), the coordinate field of G2 -- what the G2 MSM computes with) through the same two host
    builds as the base fields, against Python integers: products as two fused a b - c d, squares, inverses."""
    cv = po.CURVES[curve]
    p, n = cv.p, (8 if cv.cid == 0 else 12)
    R = 1 << (32 * n)
    rng = random.Random(cv.cid + 2)

    def enc(z):
        out = []
        for c in z:
            m = c * R % p
            out += [(m >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
        return (C.c_uint32 * (2 * n))(*out)

    def call(op, a, b=(0, 0)):
        O = (C.c_uint32 * (2 * n))()
        hostfield.ht_fp2_op(cv.cid, op, enc(a), enc(b), O)
        vals = [sum(int(O[k * n + i]) << (32 * i) for i in range(n)) * pow(R, -1, p) % p for k in range(2)]
        return tuple(vals)

    mul = lambda a, b: ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)
    special = [0, 1, p - 1, 2, (p - 1) // 2]
    for _ in range(300):
        a = tuple(rng.choice(special) if rng.random() < 0.3 else rng.randrange(p) for _ in range(2))
        b = tuple(rng.choice(special) if rng.random() < 0.3 else rng.randrange(p) for _ in range(2))
        assert call(0, a, b) == mul(a, b)
        assert call(1, a) == mul(a, a)
        assert call(2, a, b) == ((a[0] + b[0]) % p, (a[1] + b[1]) % p)
        assert call(3, a, b) == ((a[0] - b[0]) % p, (a[1] - b[1]) % p)
        assert call(4, a) == ((-a[0]) % p, (-a[1]) % p)
        assert call(6, a, b) == ((-b[0]) % p, (-b[1]) % p)
        assert call(7, a) == (2 * a[0] % p, 2 * a[1] % p)
        inv = call(5, a)
        assert (a == (0, 0) and inv == (0, 0)) or mul(a, inv) == (1, 0)


@pytest.mark.skipif(_has_gpu(), reason="the point is the behaviour on a box without a GPU")
def test_device_entry_points_fail_loudly_without_a_gpu():
    """No CPU fallback behind the device entry points added in round 2: the device batch verifier does not quietly run
    the host batch, and the solver (whose host path still keeps its key material in HBM) refuses to be created."""
    import helpers as H
    import test_verify_host as tvh
    case = H.golden_proofs()[0]
    (curve, n, nbp, cidx, vk, g1, g2), _, _ = tvh._verify_args(case)
    proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
    lib = _lib.load()
    bad = C.c_uint64()
    args = (api.CURVE_ID[curve], n, nbp, 0, None, vk, g1, g2, proof, len(proof), pub, len(pub), 1, C.byref(bad))
    assert lib.b2p_verify_batch(*args) == 0                                  # host arithmetic: works anywhere
    assert lib.b2p_verify_batch_dev(*args) == _lib.ERR_CUDA and b"CUDA" in lib.b2p_last_error()
    out = C.c_void_p()
    z = b"\0" * 256
    w = (C.c_uint32 * 8)()
    assert lib.b2p_solver_create(0, 8, 1, 4, (C.c_uint32 * 1)(0), 1, z, z, z, z, z, w, w, w, C.byref(out)) == _lib.ERR_CUDA
    with pytest.raises(_lib.B200PlonkError):
        api.msm_g2_raw("BN254", b"", [])

