// Explicit instantiation: NTT domains for both scalar fields, and the domain-sharded NTT built on them.
#define B2P_INSTANTIATE_NTT
#include "ntt_shard.cuh"
namespace b2p {
template struct NttDomain<FrBn254>;
template struct NttDomain<FrBls12381>;

template <class Fr>
static NttShardBase* make_shard(int curve, uint64_t n, uint32_t world, uint32_t rank) {
    NttShard<Fr>* s = new NttShard<Fr>();
    s->curve = curve;
    try { s->init(n, world, rank); } catch (...) { delete s; throw; }
    return s;
}
NttShardBase* new_ntt_shard(int curve, uint64_t n, uint32_t world, uint32_t rank) {
    if (curve == B2P_BN254) return make_shard<FrBn254>(curve, n, world, rank);
    if (curve == B2P_BLS12_381) return make_shard<FrBls12381>(curve, n, world, rank);
    throw Error(B2P_ERR_ARG, "unsupported curve id (B2P_BN254 = 0, B2P_BLS12_381 = 1)");
}
}
