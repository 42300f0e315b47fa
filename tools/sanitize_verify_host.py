"""AddressSanitizer + UndefinedBehaviorSanitizer, then ThreadSanitizer, over the host verifier (csrc/verify_host.hpp, pairing_host.hpp,
verify.cu): the translation unit is rebuilt with g++ -fsanitize=address,undefined next to a few extern "C" wrappers,
and every golden proof, a tampered copy of each, the pairing on the ceremony files, the vk.bin decoder and the
persisted-key parsers of keyfile.hpp (good and damaged input) go through it.  CPU only.

    python tools/sanitize_verify_host.py          # prints "sanitizers: 0 reports" and exits 0 when clean
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "algoplonk_b200", "csrc")
SHIM = r"""
#include "iface.hpp"
using namespace b2p;
extern "C" {
int s_verify(int curve, uint64_t n, uint32_t nbp, uint32_t k, const uint64_t* cidx, const void* vk, const void* g1,
             const void* g2, const void* proof, uint64_t plen, const void* pub, uint64_t publen) {
    HostVerifyKey key{n, nbp, k, cidx, vk, g1, g2};
    std::string why;
    static const unsigned char none = 0;
    return host_verify(curve, key, proof, plen, pub ? pub : &none, publen, &why) ? 1 : 0;
}
int s_verify_batch(int curve, uint64_t n, uint32_t nbp, uint32_t k, const uint64_t* cidx, const void* vk, const void* g1,
                   const void* g2, const void* proofs, uint64_t plen, const void* pubs, uint64_t publen, uint64_t count) {
    HostVerifyKey key{n, nbp, k, cidx, vk, g1, g2};
    std::string why;
    uint64_t bad = 0;
    static const unsigned char none = 0;
    return host_verify_batch(curve, key, proofs, plen, pubs ? pubs : &none, publen, count, &bad, &why) ? 1 : 0;
}
int s_pairing(int curve, const void* g1s, const void* g2s, uint64_t n) {
    std::string why;
    return host_pairing_check(curve, g1s, g2s, n, &why) ? 1 : 0;
}
int s_vk_load(int curve, const void* in, uint64_t len, void* g2, void* g1) { return host_kzg_vk_load(curve, in, len, g2, g1) ? 1 : 0; }
void s_g2_unsafe(int curve, const void* tau, void* out) { host_g2_unsafe(curve, tau, out); }
// persisted keys (keyfile.hpp): parsers of file input
int s_file_parse(const void* f, uint64_t len, b2p_gnark_file* out) { return host_gnark_file_parse(f, len, out) ? 1 : 0; }
int s_vk_parse(int curve, const void* b, uint64_t len, b2p_gnark_vk* out) { return host_gnark_vk_parse(curve, b, len, out) ? 1 : 0; }
int s_pk_parse(int curve, const void* b, uint64_t len, b2p_gnark_pk* out) { return host_gnark_pk_parse(curve, b, len, out) ? 1 : 0; }
}
"""


def build(tmp, sanitize="address,undefined"):
    shim = os.path.join(tmp, "shim.cpp")
    with open(shim, "w") as f:
        f.write(SHIM)
    lib = os.path.join(tmp, "libverify_san.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", f"-fsanitize={sanitize}",
                    *(["-fno-sanitize-recover=undefined"] if "undefined" in sanitize else []), "-DHD=inline", "-I", CSRC, "-x", "c++",
                    os.path.join(CSRC, "verify.cu"), shim, "-o", lib], check=True)
    return lib


def workload(lib_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    import test_verify_host as T
    from algoplonk_b200 import api
    from oracle import plonk_oracle as po
    lib = C.CDLL(lib_path)
    buf = lambda b: C.create_string_buffer(bytes(b), len(b))
    n_ok = n_bad = 0
    for case in H.golden_proofs():
        (curve, n, nbp, cidx, vk, g1, g2), c, _ = T._verify_args(case)
        cid = api.CURVE_ID[curve]
        proof, pub = bytes.fromhex(case["proof"]), bytes.fromhex(case["public_inputs"])
        ci = (C.c_uint64 * max(len(cidx), 1))(*cidx)
        call = lambda p, q: lib.s_verify(cid, C.c_uint64(n), nbp, len(cidx), ci, buf(vk), buf(g1), buf(g2), buf(p),
                                         C.c_uint64(len(p)), buf(q) if q else None, C.c_uint64(len(q)))
        assert call(proof, pub) == 1
        n_ok += 1
        for pos in range(0, len(proof), 37):
            bad = bytearray(proof)
            bad[pos] ^= 0x10
            assert call(bytes(bad), pub) == 0
            n_bad += 1
        import random
        rng = random.Random(len(proof))
        for _ in range(8):                                  # garbage proof / public inputs / key: any verdict, no report
            call(bytes(rng.randrange(256) for _ in range(len(proof))), pub)
            call(proof, bytes(rng.randrange(256) for _ in range(len(pub))))
            junk_vk = bytes(rng.randrange(256) for _ in range(len(vk)))
            assert lib.s_verify(cid, C.c_uint64(n), nbp, len(cidx), ci, buf(junk_vk), buf(g1), buf(g2), buf(proof),
                                C.c_uint64(len(proof)), buf(pub) if pub else None, C.c_uint64(len(pub))) == 0
        assert lib.s_verify_batch(cid, C.c_uint64(n), nbp, len(cidx), ci, buf(vk), buf(g1), buf(g2), buf(proof * 6),
                                  C.c_uint64(len(proof)), buf(pub * 6) if pub else None, C.c_uint64(len(pub)),
                                  C.c_uint64(6)) == 1      # >= 4 proofs: the threaded path
    for name in ("PerpetualPowersOfTauBN254", "DuskBLS12_381"):
        ent = H.srs_kat()[name]
        curve = ent["curve"]
        cv, cid, nb = po.CURVES[curve], api.CURVE_ID[curve], po.CURVES[curve].fp_bytes
        vk_bin = bytes.fromhex(ent["vk_bin"])
        g2, g1 = C.create_string_buffer(8 * nb), C.create_string_buffer(2 * nb)
        assert lib.s_vk_load(cid, buf(vk_bin), C.c_uint64(len(vk_bin)), g2, g1) == 0
        for pos in range(0, len(vk_bin), 7):
            bad = bytearray(vk_bin)
            bad[pos] ^= 0xA5
            lib.s_vk_load(cid, buf(bad), C.c_uint64(len(bad)), g2, g1)       # any verdict, no report
        assert lib.s_vk_load(cid, buf(vk_bin), C.c_uint64(len(vk_bin)), g2, g1) == 0
        pts = H.real_srs_points(name)
        g1s = api.points_to_mont_bytes(curve, [pts[1], po.g1_neg(cv, pts[0])])
        assert lib.s_pairing(cid, buf(g1s), g2, C.c_uint64(2)) == 1
        out = C.create_string_buffer(8 * nb)
        lib.s_g2_unsafe(cid, buf(api.fr_to_mont_bytes(curve, [cv.r - 1])), out)
    # persisted keys: the gob container and gnark's key layouts, good input and a few thousand damaged copies
    # (bit flips, truncations, garbage): any verdict, no report; exact-length heap copies so that an over-read shows
    import gnark_format as gf
    import test_keyfile as TK
    from algoplonk_b200 import _lib
    n_keys = 0
    for curve in ("BN254", "BLS12_381"):
        cid = api.CURVE_ID[curve]
        key = TK._key(curve, 2)
        blob = gf.compiled_circuit_bytes(b"\xa1ccs-bytes", key["pk"], key["vk"], gf.ECC_ID[curve])
        fi, vki, pki = _lib.GnarkFile(), _lib.GnarkVk(), _lib.GnarkPk()
        assert lib.s_file_parse(buf(blob), C.c_uint64(len(blob)), C.byref(fi)) == 0
        assert lib.s_vk_parse(cid, buf(key["vk"]), C.c_uint64(len(key["vk"])), C.byref(vki)) == 0 and vki.k == 2
        assert lib.s_pk_parse(cid, buf(key["pk"]), C.c_uint64(len(key["pk"])), C.byref(pki)) == 0
        rng = random.Random(cid)
        for data, fn in ((blob, lambda d: lib.s_file_parse(buf(d), C.c_uint64(len(d)), C.byref(fi))),
                         (key["vk"], lambda d: lib.s_vk_parse(cid, buf(d), C.c_uint64(len(d)), C.byref(vki))),
                         (key["pk"], lambda d: lib.s_pk_parse(cid, buf(d), C.c_uint64(len(d)), C.byref(pki)))):
            for cut in list(range(0, min(len(data), 300))) + [len(data) - j for j in range(1, 40)]:
                fn(data[:cut]) if cut else None
                n_keys += 1
            for _ in range(400):
                bad = bytearray(data)
                for _ in range(rng.randrange(1, 4)):
                    bad[rng.randrange(len(bad))] ^= 1 << rng.randrange(8)
                fn(bytes(bad))
                n_keys += 1
            for _ in range(50):
                fn(bytes(rng.randrange(256) for _ in range(rng.randrange(1, 400))))
                n_keys += 1
    print(f"sanitizers: 0 reports ({n_ok} proofs accepted, {n_bad} tampered proofs rejected, vk.bin fuzzed, "
          f"ceremony pairings checked, {n_keys} damaged key files / keys parsed)")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        workload(sys.argv[1])
        sys.exit(0)
    rc = 0
    for sanitize, libs in (("address,undefined", ("libasan.so", "libubsan.so")), ("thread", ("libtsan.so",))):
        with tempfile.TemporaryDirectory() as tmp:
            lib = build(tmp, sanitize)
            pre = [subprocess.run(["gcc", f"-print-file-name={n}"], capture_output=True, text=True).stdout.strip()
                   for n in libs]
            env = dict(os.environ, LD_PRELOAD=":".join(pre), ASAN_OPTIONS="detect_leaks=0:abort_on_error=1",
                       UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1", TSAN_OPTIONS="halt_on_error=1",
                       B2P_VERIFY_THREADS="4")
            print(f"-fsanitize={sanitize}:", flush=True)
            rc |= subprocess.run([sys.executable, os.path.abspath(__file__), lib], env=env).returncode
    sys.exit(rc)
