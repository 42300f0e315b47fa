// Persisted keys: the bytes utils.SerializeCompiledCircuit writes and
// utils.DeserializeCompiledCircuit reads (/root/reference/utils/utils.go:89-157) -- a gob stream holding
// CompiledCircuitBytes{Ccs, Pk, Vk []byte; Curve ecc.ID}, where Pk = plonk.ProvingKey.WriteTo and Vk =
// plonk.VerifyingKey.WriteTo of gnark v0.15.0 (go.mod:8).  Host code, no device: the parser hands back byte ranges
// and decoded key fields; b2p_srs_load_compressed decompresses the Kzg section on the GPU straight from the file.
//
// What is pinned and what is not:
//  * the gob framing follows the published encoding/gob wire format (uint / int / []byte / struct deltas);
//  * the kzg.ProvingKey section (uint32 count + compressed G1) is the format of the reference's embedded pk.bin
//    (setup/setup.go:196-228) and pinned on those files;
//  * compressed G1 / G2 and big-endian Fr are gnark-crypto's encodings, pinned by setup/trusted_setup_test.go's
//    known answers through the decoders reused here (pairing_host.hpp);
//  * the ORDER of the fields inside plonk.VerifyingKey.WriteTo is recalled from gnark's backend/plonk/<curve>/marshal.go
//    ([UPSTREAM-RECALL], no gnark source or gnark-written file on this machine): Size, SizeInv, Generator,
//    NbPublicVariables, CosetShift, S[3], Ql, Qr, Qm, Qo, Qk, Qcp, Kzg.G1, Kzg.G2[2], (Kzg.Lines), CommitmentConstraintIndexes.
//    The parser therefore VALIDATES what it reads (Size a power of two, Size * SizeInv = 1, Generator of exact order
//    Size, every point on its curve, len(Qcp) = len(CommitmentConstraintIndexes), nothing left over), so a layout
//    mismatch is an error, never a silently wrong key.  Kzg.Lines (precomputed pairing lines, fixed size) is skipped
//    by length when present.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/b200plonk.h"
#include "verify_host.hpp"

namespace b2p {
namespace keyfile {

struct Cursor {
    const uint8_t* p;
    uint64_t len, pos = 0;
    bool need(uint64_t k) const { return k <= len - pos; }
};

// encoding/gob unsigned integer: one byte below 128, else a byte holding the negated byte count and the value big-endian
inline bool gob_uint(Cursor& c, uint64_t* out) {
    if (!c.need(1)) return false;
    const uint8_t b = c.p[c.pos++];
    if (b < 0x80) { *out = b; return true; }
    const unsigned n = 256u - b;
    if (n > 8 || !c.need(n)) return false;
    uint64_t v = 0;
    for (unsigned i = 0; i < n; i++) v = (v << 8) | c.p[c.pos++];
    *out = v;
    return true;
}
// signed: bit 0 says "complemented", the value sits above it
inline bool gob_int(Cursor& c, int64_t* out) {
    uint64_t u;
    if (!gob_uint(c, &u)) return false;
    *out = (u & 1) ? (int64_t)~(u >> 1) : (int64_t)(u >> 1);
    return true;
}

// gnark-crypto's ecc.ID -> this library's curve id; -1 = a curve AlgoPlonk does not support (algoplonk.go:37-48)
inline int curve_of_ecc_id(uint64_t id) { return id == 1 ? B2P_BN254 : id == 3 ? B2P_BLS12_381 : -1; }

// The gob stream of CompiledCircuitBytes: type-definition messages (negative type id) are skipped -- []byte and uint
// are gob built-ins, the only user type is the struct itself -- then ONE value message: field deltas 1..4 with
// Ccs, Pk, Vk as (uint length, bytes) and Curve as uint; zero-valued fields are absent.  nullptr = ok
inline const char* parse_container(const uint8_t* file, uint64_t len, b2p_gnark_file* out) {
    memset(out, 0, sizeof *out);
    out->curve = -1;
    Cursor c{file, len};
    while (c.pos < c.len) {
        uint64_t mlen;
        if (!gob_uint(c, &mlen) || !c.need(mlen) || mlen == 0) return "gob: truncated message";
        Cursor m{file, c.pos + mlen, c.pos};
        c.pos += mlen;
        int64_t tid;
        if (!gob_int(m, &tid)) return "gob: bad type id";
        if (tid < 0) continue;                                   // a type definition
        int field = -1;
        for (;;) {
            uint64_t delta;
            if (!gob_uint(m, &delta)) return "gob: truncated struct";
            if (delta == 0) break;
            field += (int)delta;
            if (field > 3 || delta > 4) return "gob: not a CompiledCircuitBytes value (unknown field)";
            uint64_t v;
            if (!gob_uint(m, &v)) return "gob: truncated field";
            if (field == 3) { out->ecc_id = (uint32_t)v; out->curve = curve_of_ecc_id(v); continue; }
            if (!m.need(v)) return "gob: byte slice runs past its message";
            uint64_t* off = field == 0 ? &out->ccs_off : field == 1 ? &out->pk_off : &out->vk_off;
            uint64_t* ln = field == 0 ? &out->ccs_len : field == 1 ? &out->pk_len : &out->vk_len;
            *off = m.pos;
            *ln = v;
            m.pos += v;
        }
        if (m.pos != m.len) return "gob: trailing bytes inside the value message";
        if (c.pos != c.len) return "gob: more than one value in the file";
        if (out->curve < 0) return "compiled circuit is on a curve other than BN254 / BLS12-381";
        return nullptr;
    }
    return "gob: no value message";
}

template <class PC>
struct GnarkKeys {
    using V = hp::HostVerifier<PC>;
    using PR = hp::Pairing<PC>;
    using Fr = typename PC::Fr;
    using Fp = typename PC::Fp;
    static constexpr uint64_t FPB = PR::FPB;                 // compressed G1 = FPB bytes, G2 = 2 FPB
    // Kzg.Lines: [2][2][len(LoopCounter) - 1] LineEvaluationAff{R0, R1 E2} (gnark-crypto kzg.VerifyingKey), raw
    static constexpr uint64_t LINES = 2ull * 2 * (PC::D_TWIST ? 65 : 63) * 4 * FPB;

    static bool be64(Cursor& c, uint64_t* v) {
        if (!c.need(8)) return false;
        *v = 0;
        for (int i = 0; i < 8; i++) *v = (*v << 8) | c.p[c.pos++];
        return true;
    }
    static bool be32(Cursor& c, uint64_t* v) {
        if (!c.need(4)) return false;
        *v = 0;
        for (int i = 0; i < 4; i++) *v = (*v << 8) | c.p[c.pos++];
        return true;
    }
    static bool fr(Cursor& c, Fr* f) {
        if (!c.need(32) || !V::from_be(c.p + c.pos, 32, f, false)) return false;
        c.pos += 32;
        return true;
    }
    static const char* g1(Cursor& c, uint8_t* out) {
        if (!c.need(FPB)) return "verifying key truncated inside a G1 point";
        typename PR::G1 g;
        if (const char* e = PR::g1_decompress(c.p + c.pos, &g)) return e;
        if (!PC::D_TWIST && !g.inf && !V::in_g1_subgroup({g.x, g.y, g.inf})) return "G1 point not in the r-torsion subgroup";
        c.pos += FPB;
        if (g.inf) memset(out, 0, 2 * FPB);
        else { g.x.store(out); g.y.store(out + FPB); }
        return nullptr;
    }

    // The VerifyingKey at the head of `bytes`, read with (lines = true) or without the Kzg.Lines block
    static const char* parse_vk_prefix(const uint8_t* bytes, uint64_t len, bool lines, b2p_gnark_vk* out) {
        memset(out, 0, sizeof *out);
        Cursor c{bytes, len};
        Fr size_inv, gen, shift;
        if (!be64(c, &out->size)) return "verifying key truncated";
        if (!fr(c, &size_inv) || !fr(c, &gen)) return "verifying key: SizeInv / Generator not a reduced scalar";
        if (!be64(c, &out->nb_public)) return "verifying key truncated";
        if (!fr(c, &shift)) return "verifying key: CosetShift not a reduced scalar";
        const uint64_t n = out->size;
        if (n < 2 || (n & (n - 1)) || n > (1ull << PC::FrP::TWO_ADICITY)) return "verifying key: Size is not a power of two the field supports";
        if (!(Fr::mul(Fr::from_u64(n), size_inv) == Fr::one())) return "verifying key: SizeInv is not 1 / Size";
        {   // Generator has exact order Size
            Fr g = gen;
            for (uint64_t m = n; m > 2; m >>= 1) g = g.sqr();
            if (g == Fr::one() || !(g.sqr() == Fr::one())) return "verifying key: Generator does not have order Size";
        }
        if (out->nb_public > n) return "verifying key: more public variables than rows";
        size_inv.store(out->size_inv); gen.store(out->generator); shift.store(out->coset_shift);
        for (int i = 0; i < 8; i++)
            if (const char* e = g1(c, out->points + (uint64_t)i * 2 * FPB)) return e;
        uint64_t k;
        if (!be32(c, &k)) return "verifying key truncated before Qcp";
        if (k > B2P_MAX_COMMITMENTS) return "verifying key: more BSB22 commitments than this library supports";
        out->k = (uint32_t)k;
        for (uint64_t i = 0; i < k; i++)
            if (const char* e = g1(c, out->points + (8 + i) * 2 * FPB)) return e;
        if (const char* e = g1(c, out->kzg_g1)) return e;
        for (int i = 0; i < 2; i++) {
            if (!c.need(2 * FPB)) return "verifying key truncated inside Kzg.G2";
            typename PR::G2 q;
            if (const char* e = PR::g2_decompress(c.p + c.pos, &q)) return e;
            c.pos += 2 * FPB;
            PR::store_g2(q, out->kzg_g2 + (uint64_t)i * 4 * FPB);
        }
        if (lines) {
            if (!c.need(LINES)) return "verifying key truncated inside Kzg.Lines";
            out->has_lines = 1;
            c.pos += LINES;
        }
        uint64_t k2;
        if (!be32(c, &k2)) return "verifying key truncated before CommitmentConstraintIndexes";
        if (k2 != k) return "verifying key: len(CommitmentConstraintIndexes) differs from len(Qcp) (field order mismatch?)";
        for (uint64_t i = 0; i < k; i++) {
            if (!be64(c, &out->commitment_indexes[i])) return "verifying key truncated inside CommitmentConstraintIndexes";
            if (out->nb_public + out->commitment_indexes[i] >= n) return "verifying key: commitment constraint index out of range";
        }
        out->encoded_len = c.pos;
        return nullptr;
    }

    // A stand-alone VerifyingKey (the Vk field of the file): one of the two readings must consume every byte
    static const char* parse_vk(const uint8_t* bytes, uint64_t len, b2p_gnark_vk* out) {
        const char* first = nullptr;
        for (int lines = 1; lines >= 0; lines--) {
            const char* e = parse_vk_prefix(bytes, len, lines != 0, out);
            if (!e && out->encoded_len != len) e = "verifying key: trailing bytes";
            if (!e) return nullptr;
            if (lines) first = e;
        }
        return first;
    }

    // plonk.ProvingKey.WriteTo = VerifyingKey, Kzg, KzgLagrange (two kzg.ProvingKey: uint32 count + compressed G1)
    static const char* parse_pk(const uint8_t* bytes, uint64_t len, b2p_gnark_pk* out) {
        const char* first = nullptr;
        for (int lines = 1; lines >= 0; lines--) {      // whichever reading of the VerifyingKey makes the rest fit exactly
            const char* e = parse_pk_as(bytes, len, lines != 0, out);
            if (!e) return nullptr;
            if (lines) first = e;
        }
        return first;
    }
    static const char* parse_pk_as(const uint8_t* bytes, uint64_t len, bool lines, b2p_gnark_pk* out) {
        memset(out, 0, sizeof *out);
        if (const char* e = parse_vk_prefix(bytes, len, lines, &out->vk)) return e;
        Cursor c{bytes, len, out->vk.encoded_len};
        out->kzg_off = c.pos;
        if (!be32(c, &out->kzg_count)) return "proving key truncated before Kzg";
        if (!c.need(out->kzg_count * FPB)) return "proving key truncated inside Kzg";
        c.pos += out->kzg_count * FPB;
        out->lagrange_off = c.pos;
        if (!be32(c, &out->lagrange_count)) return "proving key truncated before KzgLagrange";
        if (!c.need(out->lagrange_count * FPB)) return "proving key truncated inside KzgLagrange";
        c.pos += out->lagrange_count * FPB;
        if (c.pos != len) return "proving key: trailing bytes";
        // plonk.Setup's sizes (setup/setup.go:113-114,124,138): n + 3 canonical points, n Lagrange points
        if (out->kzg_count < out->vk.size + 3) return "proving key: Kzg holds fewer than Size + 3 points";
        if (out->lagrange_count != out->vk.size) return "proving key: KzgLagrange does not hold Size points";
        return nullptr;
    }
};

inline const char* parse_vk(int curve, const void* b, uint64_t len, b2p_gnark_vk* out) {
    const uint8_t* p = static_cast<const uint8_t*>(b);
    return curve == 0 ? GnarkKeys<hp::Bn254Pairing>::parse_vk(p, len, out) : GnarkKeys<hp::Bls12381Pairing>::parse_vk(p, len, out);
}
inline const char* parse_pk(int curve, const void* b, uint64_t len, b2p_gnark_pk* out) {
    const uint8_t* p = static_cast<const uint8_t*>(b);
    return curve == 0 ? GnarkKeys<hp::Bn254Pairing>::parse_pk(p, len, out) : GnarkKeys<hp::Bls12381Pairing>::parse_pk(p, len, out);
}

}  // namespace keyfile
}  // namespace b2p
