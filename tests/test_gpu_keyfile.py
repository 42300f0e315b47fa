"""GPU side of the persisted-key path (utils/utils.go:89-157): a key file in gnark's layout whose
Kzg section is the reference's real pk.bin bytes -> parse -> SRS decompressed on the GPU from the file's own bytes ->
proofs equal to the golden ones; and the library's own snapshot: save, load, same proof bytes."""
import pytest

import gnark_format as gf
import helpers as H
from algoplonk_b200 import _lib, api
from oracle import plonk_oracle as po

pytestmark = pytest.mark.gpu
REAL = {"BN254": ("PerpetualPowersOfTauBN254", api.SetupName.PerpetualPowersOfTauBN254),
        "BLS12_381": ("DuskBLS12_381", api.SetupName.DuskBLS12381)}


def _real_cases():
    return [c for c in H.golden_proofs() if c["srs"] in ("PerpetualPowersOfTauBN254", "DuskBLS12_381")]


@pytest.mark.parametrize("case", _real_cases(), ids=H.case_id)
def test_gnark_key_file_to_resident_key(gpu, case, tmp_path):
    curve = case["curve"]
    cv = po.CURVES[curve]
    name, setup = REAL[curve]
    ent = H.srs_kat()[name]
    c = H.build_case(case)
    tc = c["tc"]
    n, k = tc.n, len(tc.qcp)
    raw = bytes.fromhex(ent["first"])
    vk_bin = bytes.fromhex(ent["vk_bin"])
    # what plonk.Setup + SerializeCompiledCircuit would leave on disk for this circuit
    first = api.Compile(c["cs"], curve, setup, srs=api.SRS.from_points(curve, c["srs"]))
    vk_points = first.vk_commitments()
    lagrange = first.srs.to_lagrange(n)
    first.free()
    vk = gf.plonk_vk_bytes(curve, n, tc.nb_public, vk_points, c["srs"][0], vk_bin[:4 * cv.fp_bytes],
                           tc.commitment_constraint_indexes)
    kzg = gf.kzg_pk_bytes(raw[:(n + 3) * cv.fp_bytes], n + 3)
    lag = gf.kzg_pk_bytes(b"".join(po.g1_compress(cv, P) for P in lagrange), n)
    blob = gf.compiled_circuit_bytes(b"\xa1ccs", gf.plonk_pk_bytes(vk, kzg, lag), vk, gf.ECC_ID[curve])

    info = api.parse_gnark_file(blob)
    assert info.curve == api.CURVE_ID[curve]
    pk_bytes = blob[info.pk_off:info.pk_off + info.pk_len]
    pk = api.parse_gnark_pk(curve, pk_bytes)
    assert (pk.vk.size, pk.vk.nb_public, pk.vk.k, pk.kzg_count, pk.lagrange_count) == (n, tc.nb_public, k, n + 3, n)
    nb = 2 * cv.fp_bytes
    assert api.points_from_mont_bytes(curve, bytes(pk.vk.points)[:(8 + k) * nb]) == vk_points
    srs = api.srs_from_gnark_pk(curve, pk_bytes)                 # GPU decompression of the file's Kzg section
    assert srs.size == n + 3 and srs.points(0, n + 3) == c["srs"][:n + 3]
    # the verifying key's transcript bytes come from the file too (what the Go shim passes from pk.Vk)
    vkt = b"".join(po.g1_raw_bytes(cv, P, gnark_infinity_flag=True) for P in
                   api.points_from_mont_bytes(curve, bytes(pk.vk.points)[:(8 + k) * nb]))
    cc = api.Compile(c["cs"], curve, setup, srs=srs, vk_transcript=vkt)
    vp = cc.Verify(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"])      # b2p_verify against the file's G2
    assert api.MarshalProof(vp.Proof).hex() == case["proof"]
    # KzgLagrange of the file is what ToLagrangeG1 makes of its Kzg (setup/setup.go:124,138)
    assert pk_bytes[pk.lagrange_off + 4:] == b"".join(po.g1_compress(cv, P) for P in srs.to_lagrange(n))

    # the library's own snapshot of the circuit half
    path = str(tmp_path / "key.b2pk")
    api.SerializeCompiledCircuit(cc, path, vk_transcript=vkt)
    cc.free()
    cc2 = api.DeserializeCompiledCircuit(path, c["cs"], srs)
    assert cc2.vk_commitments() == vk_points
    vp2 = cc2.Verify(c["L"], c["R"], c["O"], c["blinding"], c["pi2"], c["coms"])
    assert api.MarshalProof(vp2.Proof).hex() == case["proof"]
    cc2.free()

    # a damaged snapshot is refused, with the reason
    data = bytearray(open(path, "rb").read())
    data[200] ^= 1
    open(path, "wb").write(bytes(data))
    with pytest.raises(_lib.B200PlonkError, match="checksum"):
        api.DeserializeCompiledCircuit(path, c["cs"], srs)
    open(path, "wb").write(bytes(data[:-64]))
    with pytest.raises(_lib.B200PlonkError, match="length"):
        api.DeserializeCompiledCircuit(path, c["cs"], srs)
    other = "BLS12_381" if curve == "BN254" else "BN254"
    data[200] ^= 1
    open(path, "wb").write(bytes(data))
    srs_other = api.SRS.unsafe(other, n + 3)
    with pytest.raises(_lib.B200PlonkError, match="other curve"):
        api.DeserializeCompiledCircuit(path, c["cs"], srs_other)
    srs_other.free()
    srs.free()
