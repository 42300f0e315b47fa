// Explicit instantiation: MSM engine, Bn254.
#define B2P_INSTANTIATE_MSM
#include "msm.cuh"
namespace b2p {
template struct MsmEngine<Bn254>;
}
