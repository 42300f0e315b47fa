// Explicit instantiation: NTT domains for both scalar fields.
#define B2P_INSTANTIATE_NTT
#include "ntt.cuh"
namespace b2p {
template struct NttDomain<FrBn254>;
template struct NttDomain<FrBls12381>;
}
