// Explicit instantiation: MSM engine, Bls12381.
#define B2P_INSTANTIATE_MSM
#include "msm.cuh"
namespace b2p {
template struct MsmEngine<Bls12381>;
}
