// Explicit instantiation: batch verification with the group arithmetic on the GPU (verify_batch.cuh), both curves.
#include "verify_batch.cuh"
namespace b2p {

template <class C, class PC>
static bool run_batch(const HostVerifyKey& vk, const uint8_t* proofs, uint64_t proof_len, const uint8_t* pubs,
                      uint64_t pub_len, uint64_t count, uint64_t* bad, std::string* why) {
    using V = hp::HostVerifier<PC>;
    typename V::Key k{vk.n, vk.nb_public, vk.k, vk.commit_idx, static_cast<const uint8_t*>(vk.vk_points),
                      static_cast<const uint8_t*>(vk.g1), static_cast<const uint8_t*>(vk.g2)};
    DeviceBatchVerifier<C, PC> dv;
    return dv.verify(k, proofs, proof_len, pubs, pub_len, count, bad, why);
}

bool device_verify_batch(int curve, const HostVerifyKey& vk, const void* proofs, uint64_t proof_len, const void* pubs,
                         uint64_t pub_len, uint64_t count, uint64_t* bad, std::string* why) {
    const uint8_t* p = static_cast<const uint8_t*>(proofs);
    const uint8_t* q = static_cast<const uint8_t*>(pubs);
    return curve == 0 ? run_batch<Bn254, hp::Bn254Pairing>(vk, p, proof_len, q, pub_len, count, bad, why)
                      : run_batch<Bls12381, hp::Bls12381Pairing>(vk, p, proof_len, q, pub_len, count, bad, why);
}

}  // namespace b2p
