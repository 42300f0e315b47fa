// Host-only translation unit: plonk.Verify, the pairing-product check and the G2 half of a known-tau SRS,
// per curve (verify_host.hpp, pairing_host.hpp).  No device code and no CUDA calls: these entry points also
// work on a box without a GPU, exactly like gnark's verifier which they stand in for.
#include "iface.hpp"
#include "verify_host.hpp"
#include "keyfile.hpp"

namespace b2p {

template <class PC>
static bool verify_t(const HostVerifyKey& vk, const uint8_t* proof, uint64_t proof_len, const uint8_t* pub,
                     uint64_t pub_len, std::string* why) {
    using V = hp::HostVerifier<PC>;
    typename V::Key k{vk.n, vk.nb_public, vk.k, vk.commit_idx, static_cast<const uint8_t*>(vk.vk_points),
                      static_cast<const uint8_t*>(vk.g1), static_cast<const uint8_t*>(vk.g2)};
    return V::verify(k, proof, proof_len, pub, pub_len, why);
}

bool host_verify(int curve, const HostVerifyKey& vk, const void* proof, uint64_t proof_len, const void* pub,
                 uint64_t pub_len, std::string* why) {
    const uint8_t* p = static_cast<const uint8_t*>(proof);
    const uint8_t* q = static_cast<const uint8_t*>(pub);
    return curve == 0 ? verify_t<hp::Bn254Pairing>(vk, p, proof_len, q, pub_len, why)
                      : verify_t<hp::Bls12381Pairing>(vk, p, proof_len, q, pub_len, why);
}

template <class PC>
static bool verify_batch_t(const HostVerifyKey& vk, const uint8_t* proofs, uint64_t proof_len, const uint8_t* pubs,
                           uint64_t pub_len, uint64_t count, uint64_t* bad, std::string* why) {
    using V = hp::HostVerifier<PC>;
    typename V::Key k{vk.n, vk.nb_public, vk.k, vk.commit_idx, static_cast<const uint8_t*>(vk.vk_points),
                      static_cast<const uint8_t*>(vk.g1), static_cast<const uint8_t*>(vk.g2)};
    return V::verify_batch(k, proofs, proof_len, pubs, pub_len, count, bad, why);
}

bool host_verify_batch(int curve, const HostVerifyKey& vk, const void* proofs, uint64_t proof_len, const void* pubs,
                       uint64_t pub_len, uint64_t count, uint64_t* bad, std::string* why) {
    const uint8_t* p = static_cast<const uint8_t*>(proofs);
    const uint8_t* q = static_cast<const uint8_t*>(pubs);
    return curve == 0 ? verify_batch_t<hp::Bn254Pairing>(vk, p, proof_len, q, pub_len, count, bad, why)
                      : verify_batch_t<hp::Bls12381Pairing>(vk, p, proof_len, q, pub_len, count, bad, why);
}

bool host_pairing_check(int curve, const void* g1s, const void* g2s, uint64_t n, std::string* why) {
    const uint8_t* a = static_cast<const uint8_t*>(g1s);
    const uint8_t* b = static_cast<const uint8_t*>(g2s);
    return curve == 0 ? hp::Pairing<hp::Bn254Pairing>::product_is_one(a, b, n, why)
                      : hp::Pairing<hp::Bls12381Pairing>::product_is_one(a, b, n, why);
}

template <class PC>
static const char* kzg_vk_load_t(const uint8_t* in, uint64_t len, uint8_t* out_g2, uint8_t* out_g1) {
    using PR = hp::Pairing<PC>;
    if (len != 5ull * PR::FPB) return "vk.bin has the wrong length (2 compressed G2 + 1 compressed G1)";
    typename PR::G2 q[2];
    typename PR::G1 g;
    for (int i = 0; i < 2; i++)
        if (const char* e = PR::g2_decompress(in + 2 * i * PR::FPB, &q[i])) return e;
    if (const char* e = PR::g1_decompress(in + 4 * PR::FPB, &g)) return e;
    // gnark's decoders check subgroup membership; G2 has a cofactor on both curves, G1 on BLS12-381: [r] Q == infinity
    using Fr = typename PC::Fr;
    uint64_t r[Fr::N];
    for (int i = 0; i < Fr::N; i++) r[i] = Fr::M(i);
    for (int i = 0; i < 2; i++)
        if (!PR::g2_mul(q[i], r, Fr::N).inf) return "vk.bin: a G2 point is not in the r-torsion subgroup";
    if (!PC::D_TWIST && !hp::HostVerifier<PC>::in_g1_subgroup({g.x, g.y, g.inf})) return "vk.bin: the G1 point is not in the r-torsion subgroup";
    PR::store_g2(q[0], out_g2);
    PR::store_g2(q[1], out_g2 + 4 * PR::FPB);
    if (g.inf) memset(out_g1, 0, 2 * PR::FPB);
    else { g.x.store(out_g1); g.y.store(out_g1 + PR::FPB); }
    return nullptr;
}

// kzg.VerifyingKey.ReadFrom on an embedded setup/<name>/vk.bin (setup/setup.go:174,190): nullptr = ok
const char* host_kzg_vk_load(int curve, const void* vk_bin, uint64_t len, void* out_g2, void* out_g1) {
    const uint8_t* in = static_cast<const uint8_t*>(vk_bin);
    return curve == 0 ? kzg_vk_load_t<hp::Bn254Pairing>(in, len, static_cast<uint8_t*>(out_g2), static_cast<uint8_t*>(out_g1))
                      : kzg_vk_load_t<hp::Bls12381Pairing>(in, len, static_cast<uint8_t*>(out_g2), static_cast<uint8_t*>(out_g1));
}

template <class PC>
static void g2_unsafe_t(const void* tau_mont, uint8_t* out) {
    using PR = hp::Pairing<PC>;
    using Fr = typename PC::Fr;
    const Fr tau = Fr::load(tau_mont).from_mont();
    const typename PR::G2 g = PR::g2_generator();
    PR::store_g2(g, out);
    PR::store_g2(PR::g2_mul(g, tau.v, Fr::N), out + 4 * PR::FPB);
}

// [1]_2, [tau]_2: the G2 half of unsafekzg.NewSRS (setup/setup.go:124), for the TestOnly setups
void host_g2_unsafe(int curve, const void* tau_mont, void* out) {
    if (curve == 0) g2_unsafe_t<hp::Bn254Pairing>(tau_mont, static_cast<uint8_t*>(out));
    else g2_unsafe_t<hp::Bls12381Pairing>(tau_mont, static_cast<uint8_t*>(out));
}

// Persisted keys (keyfile.hpp): utils.DeserializeCompiledCircuit's file, gnark's VerifyingKey / ProvingKey encodings
const char* host_gnark_file_parse(const void* file, uint64_t len, b2p_gnark_file* out) {
    return keyfile::parse_container(static_cast<const uint8_t*>(file), len, out);
}
const char* host_gnark_vk_parse(int curve, const void* b, uint64_t len, b2p_gnark_vk* out) {
    return keyfile::parse_vk(curve, b, len, out);
}
const char* host_gnark_pk_parse(int curve, const void* b, uint64_t len, b2p_gnark_pk* out) {
    return keyfile::parse_pk(curve, b, len, out);
}

}  // namespace b2p
