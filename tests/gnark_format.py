"""Writers for the bytes gnark v0.15.0 persists (TEST INFRASTRUCTURE): the gob stream of utils.CompiledCircuitBytes
(/root/reference/utils/utils.go:89-121) and plonk.VerifyingKey / ProvingKey.WriteTo.  The gob framing follows the
encoding/gob specification (the worked `Point` example of its package documentation); the key layouts are recalled
from gnark's backend/plonk/<curve>/marshal.go -- there is no gnark-written file on this machine to diff against, so
these writers pin the parser's behaviour, not gnark's ("persisted keys")."""
from oracle import plonk_oracle as po


def gob_uint(v: int) -> bytes:
    if v < 128:
        return bytes([v])
    b = v.to_bytes((v.bit_length() + 7) // 8, "big")
    return bytes([256 - len(b)]) + b


def gob_int(i: int) -> bytes:
    return gob_uint(((~i) << 1) | 1 if i < 0 else i << 1)


def gob_message(content: bytes) -> bytes:
    return gob_uint(len(content)) + content


T_UINT, T_BYTES = 3, 5            # gob's built-in type ids


def gob_struct_typedef(type_id: int, name: str, fields) -> bytes:
    """wireType{StructT: structType{CommonType{Name, Id}, Field []fieldType{Name, Id}}} for a user struct."""
    body = gob_int(-type_id) + b"\x03" + b"\x01" + b"\x01" + gob_uint(len(name)) + name.encode() + b"\x01" + gob_int(type_id) + b"\x00"
    body += b"\x01" + gob_uint(len(fields))
    for fname, ftype in fields:
        body += b"\x01" + gob_uint(len(fname)) + fname.encode() + b"\x01" + gob_int(ftype) + b"\x00"
    body += b"\x00\x00"
    return gob_message(body)


def compiled_circuit_bytes(ccs: bytes, pk: bytes, vk: bytes, ecc_id: int) -> bytes:
    """gob.NewEncoder(&buf).Encode(CompiledCircuitBytes{Ccs, Pk, Vk, Curve}): one type definition, one value."""
    out = gob_struct_typedef(65, "CompiledCircuitBytes",
                             [("Ccs", T_BYTES), ("Pk", T_BYTES), ("Vk", T_BYTES), ("Curve", T_UINT)])
    val = gob_int(65)
    delta = 1
    for blob in (ccs, pk, vk):
        if blob:                                   # gob omits zero values
            val += gob_uint(delta) + gob_uint(len(blob)) + blob
            delta = 1
        else:
            delta += 1
    if ecc_id:
        val += gob_uint(delta) + gob_uint(ecc_id)
    val += b"\x00"
    return out + gob_message(val)


ECC_ID = {"BN254": 1, "BLS12_381": 3}


def lines_size(curve: str) -> int:
    cv = po.CURVES[curve]
    return 2 * 2 * (65 if cv.cid == 0 else 63) * 4 * cv.fp_bytes


def plonk_vk_bytes(curve: str, size: int, nb_public: int, vk_points, kzg_g1, g2_compressed: bytes,
                   commitment_indexes=(), coset_shift=None, with_lines=True) -> bytes:
    """vk_points: S1 S2 S3 Ql Qr Qm Qo Qk then Qcp* (affine ints); g2_compressed: G2[0] || G2[1] as in a vk.bin."""
    cv = po.CURVES[curve]
    k = len(commitment_indexes)
    assert len(vk_points) == 8 + k
    gen = po.domain_generator(cv, size)
    shift = coset_shift if coset_shift is not None else (5 if cv.cid == 0 else 7)
    fr = lambda v: (v % cv.r).to_bytes(32, "big")
    out = size.to_bytes(8, "big") + fr(pow(size, -1, cv.r)) + fr(gen) + nb_public.to_bytes(8, "big") + fr(shift)
    for P in vk_points[:8]:
        out += po.g1_compress(cv, P)
    out += k.to_bytes(4, "big")
    for P in vk_points[8:]:
        out += po.g1_compress(cv, P)
    out += po.g1_compress(cv, kzg_g1) + g2_compressed
    if with_lines:
        out += bytes(lines_size(curve))
    out += k.to_bytes(4, "big")
    for i in commitment_indexes:
        out += int(i).to_bytes(8, "big")
    return out


def kzg_pk_bytes(compressed_points: bytes, count: int) -> bytes:
    """kzg.ProvingKey.WriteTo: uint32 count + compressed G1 (the format of setup/<name>/pk.bin)."""
    return count.to_bytes(4, "big") + compressed_points


def plonk_pk_bytes(vk: bytes, kzg: bytes, kzg_lagrange: bytes) -> bytes:
    return vk + kzg + kzg_lagrange
