"""Persisted keys (utils/utils.go:89-157): the gob container, gnark's VerifyingKey / ProvingKey
encodings and the library's own snapshot header.  Host code only -- no GPU.  The Kzg section and the G2 points are the
reference's own pk.bin / vk.bin bytes; the field order of the VerifyingKey is recalled (see tests/gnark_format.py)."""
import ctypes as C
import os
import random

import pytest

import gnark_format as gf
import helpers as H
from algoplonk_b200 import _lib, api
from oracle import plonk_oracle as po

REAL = {"BN254": "PerpetualPowersOfTauBN254", "BLS12_381": "DuskBLS12_381"}
CURVES = ("BN254", "BLS12_381")


def _key(curve, k=0, size=8, with_lines=True):
    ent = H.srs_kat()[REAL[curve]]
    cv = po.CURVES[curve]
    pts = H.real_srs_points(REAL[curve])
    raw = bytes.fromhex(ent["first"])
    vk_bin = bytes.fromhex(ent["vk_bin"])
    vk_points = [pts[3 + i] for i in range(8 + k)]
    cidx = [2 + 3 * i for i in range(k)]
    size = size if k == 0 else max(size, 16)
    vk = gf.plonk_vk_bytes(curve, size, 1, vk_points, pts[0], vk_bin[:4 * cv.fp_bytes], cidx, with_lines=with_lines)
    kzg = gf.kzg_pk_bytes(raw[:(size + 3) * cv.fp_bytes], size + 3)
    lag = gf.kzg_pk_bytes(raw[:size * cv.fp_bytes], size)       # any `size` valid points (not Lagrange: never decoded here)
    return dict(vk=vk, pk=gf.plonk_pk_bytes(vk, kzg, lag), vk_points=vk_points, cidx=cidx, size=size, pts=pts,
                g2=H.real_srs_g2(REAL[curve]), kzg=kzg)


def test_gob_primitives_against_the_package_documentation():
    """encoding/gob's documented examples: 7 -> 07, 256 -> FE 01 00, -129 as int -> FE 01 01; and the Point value."""
    assert gf.gob_uint(7) == b"\x07" and gf.gob_uint(256) == b"\xfe\x01\x00"
    assert gf.gob_int(-129) == b"\xfe\x01\x01" and gf.gob_int(65) == b"\xff\x82" and gf.gob_int(-65) == b"\xff\x81"
    # type Point struct{X, Y int}: the 31-byte descriptor of the documentation
    td = gf.gob_struct_typedef(65, "Point", [("X", 2), ("Y", 2)])
    assert td.hex() == "1fff810301010550" "6f696e7401ff8200" "0102010158010400" "0101590104000000"


@pytest.mark.parametrize("curve", CURVES)
def test_container_ranges(curve):
    rng = random.Random(1)
    ccs = bytes(rng.randrange(256) for _ in range(301))
    key = _key(curve)
    blob = gf.compiled_circuit_bytes(ccs, key["pk"], key["vk"], gf.ECC_ID[curve])
    info = api.parse_gnark_file(blob)
    assert info.curve == api.CURVE_ID[curve] and info.ecc_id == gf.ECC_ID[curve]
    assert blob[info.ccs_off:info.ccs_off + info.ccs_len] == ccs
    assert blob[info.pk_off:info.pk_off + info.pk_len] == key["pk"]
    assert blob[info.vk_off:info.vk_off + info.vk_len] == key["vk"]
    # an empty Ccs is omitted by gob (zero value): the next field's delta is 2
    info = api.parse_gnark_file(gf.compiled_circuit_bytes(b"", key["pk"], key["vk"], gf.ECC_ID[curve]))
    assert info.ccs_len == 0 and info.pk_len == len(key["pk"]) and info.vk_len == len(key["vk"])


def test_container_rejects():
    key = _key("BN254")
    good = gf.compiled_circuit_bytes(b"abc", key["pk"], key["vk"], 1)
    for bad in (good[:-1], good[:40], good + b"\x00", b"", b"\x05abc",
                gf.compiled_circuit_bytes(b"abc", key["pk"], key["vk"], 2),       # BLS12-377: not an AlgoPlonk curve
                gf.compiled_circuit_bytes(b"abc", key["pk"], key["vk"], 0)):      # UNKNOWN
        with pytest.raises(_lib.B200PlonkError) as e:
            api.parse_gnark_file(bad)
        assert e.value.code == _lib.ERR_ARG
    # a slice length that runs past its message
    cut = bytearray(good)
    i = good.index(b"abc") - 1
    cut[i] = 0x7F
    with pytest.raises(_lib.B200PlonkError):
        api.parse_gnark_file(bytes(cut))


@pytest.mark.parametrize("curve", CURVES)
@pytest.mark.parametrize("k", (0, 1, 2))
@pytest.mark.parametrize("with_lines", (True, False))
def test_vk_fields(curve, k, with_lines):
    cv = po.CURVES[curve]
    key = _key(curve, k, with_lines=with_lines)
    vk = api.parse_gnark_vk(curve, key["vk"])
    assert (vk.size, vk.nb_public, vk.k, vk.has_lines, vk.encoded_len) == (key["size"], 1, k, int(with_lines), len(key["vk"]))
    assert list(vk.commitment_indexes)[:k] == key["cidx"]
    nb = 2 * cv.fp_bytes
    assert api.points_from_mont_bytes(curve, bytes(vk.points)[:(8 + k) * nb]) == key["vk_points"]
    assert api.points_from_mont_bytes(curve, bytes(vk.kzg_g1)[:nb]) == [key["pts"][0]]
    assert tuple(api.g2_from_mont_bytes(curve, bytes(vk.kzg_g2)[:4 * nb])) == key["g2"]
    assert api.fr_from_mont_bytes(curve, bytes(vk.size_inv)) == [pow(key["size"], -1, cv.r)]
    assert api.fr_from_mont_bytes(curve, bytes(vk.generator)) == [po.domain_generator(cv, key["size"])]
    assert api.fr_from_mont_bytes(curve, bytes(vk.coset_shift)) == [5 if cv.cid == 0 else 7]


@pytest.mark.parametrize("curve", CURVES)
def test_vk_rejects_what_does_not_decode(curve):
    cv = po.CURVES[curve]
    key = _key(curve, 1)
    good = key["vk"]
    assert api.parse_gnark_vk(curve, good).k == 1

    def bad(b):
        with pytest.raises(_lib.B200PlonkError) as e:
            api.parse_gnark_vk(curve, bytes(b))
        assert e.value.code == _lib.ERR_ARG
        return str(e.value)

    assert "truncated" in bad(good[:-1]) or "differs" in bad(good[:-1])
    bad(good + b"\x00")                                                        # left-over bytes shift the tail
    b = bytearray(good); b[7] = 9                                               # Size 9: not a power of two
    assert "power of two" in bad(b)
    b = bytearray(good); b[8 + 31] ^= 1                                          # SizeInv
    assert "SizeInv" in bad(b)
    b = bytearray(good); b[8 + 32 + 31] ^= 1                                     # Generator
    assert "Generator" in bad(b)
    b = bytearray(good); b[8:8 + 32] = cv.r.to_bytes(32, "big")                  # unreduced scalar
    assert "reduced" in bad(b)
    first_point = 8 + 32 + 32 + 8 + 32
    b = bytearray(good); b[first_point] &= 0x1F                                  # flag bits cleared: not a compressed point
    bad(b)
    b = bytearray(good); b[first_point + cv.fp_bytes - 1] ^= 1                   # another x: half of them are off the curve
    try:
        api.parse_gnark_vk(curve, bytes(b))
    except _lib.B200PlonkError:
        pass
    # the other curve's key does not parse
    other = "BLS12_381" if curve == "BN254" else "BN254"
    with pytest.raises(_lib.B200PlonkError):
        api.parse_gnark_vk(other, good)
    # Qcp count and index count disagree
    idx_tail = len(good) - (4 + 8)
    b = bytearray(good); b[idx_tail + 3] = 2
    bad(b)
    # an index beyond the rows
    b = bytearray(good); b[-8:] = (1 << 40).to_bytes(8, "big")
    assert "out of range" in bad(b)


@pytest.mark.parametrize("curve", CURVES)
def test_pk_sections(curve):
    cv = po.CURVES[curve]
    key = _key(curve, 1)
    pk = api.parse_gnark_pk(curve, key["pk"])
    n = key["size"]
    assert pk.vk.size == n and pk.vk.k == 1
    assert (pk.kzg_off, pk.kzg_count) == (len(key["vk"]), n + 3)
    assert (pk.lagrange_off, pk.lagrange_count) == (len(key["vk"]) + 4 + (n + 3) * cv.fp_bytes, n)
    assert key["pk"][pk.kzg_off:pk.lagrange_off] == key["kzg"]           # what b2p_srs_load_compressed is handed
    for badpk in (key["pk"][:-1], key["pk"] + b"\x00",
                  gf.plonk_pk_bytes(key["vk"], gf.kzg_pk_bytes(b"", 0), gf.kzg_pk_bytes(b"", 0)),       # too few points
                  gf.plonk_pk_bytes(key["vk"], key["kzg"], gf.kzg_pk_bytes(key["kzg"][4:4 + cv.fp_bytes], 1))):
        with pytest.raises(_lib.B200PlonkError):
            api.parse_gnark_pk(curve, badpk)


def test_snapshot_header_checks(tmp_path):
    """b2p_circuit_save writes without a device; a damaged file must be refused before anything is uploaded --
    exercised here through the argument checks that need no SRS handle."""
    lib = _lib.load()
    n = 8
    col = (C.c_uint8 * (32 * n))(*range(256))
    perm = (C.c_int64 * (3 * n))(*range(3 * n))
    path = str(tmp_path / "key.b2pk").encode()
    assert lib.b2p_circuit_save(path, 0, n, 1, col, col, col, col, col, perm, 0, None, None, b"xyz", 3) == 0
    data = open(path, "rb").read()
    assert data[:6] == b"B2PKEY" and len(data) == 64 + 5 * 32 * n + 3 * n * 8 + 64
    assert lib.b2p_circuit_save(path, 0, 7, 1, col, col, col, col, col, perm, 0, None, None, None, 0) == _lib.ERR_ARG
    assert lib.b2p_circuit_save(path, 0, n, 1, col, col, col, col, col, perm, 1, None, None, None, 0) == _lib.ERR_ARG
    assert lib.b2p_circuit_save(path, 5, n, 1, col, col, col, col, col, perm, 0, None, None, None, 0) == _lib.ERR_ARG
    assert lib.b2p_circuit_save(b"/nonexistent-dir/x", 0, n, 1, col, col, col, col, col, perm, 0, None, None, None, 0) == _lib.ERR_ARG
    out = C.c_void_p()
    assert lib.b2p_circuit_load_file(None, path, C.byref(out)) == _lib.ERR_ARG


def test_should_recompile(tmp_path):
    a, b = tmp_path / "a", tmp_path / "b"
    assert api.ShouldRecompile(str(a), str(b))                       # both missing
    a.write_text("x"); b.write_text("y")
    os.utime(a, ns=(1_000, 1_000)); os.utime(b, ns=(2_000, 2_000))
    assert api.ShouldRecompile(str(a), str(b))                       # source newer than target
    assert not api.ShouldRecompile(str(b), str(a))
    assert api.ShouldRecompile(str(b), str(a), str(tmp_path / "missing"))
