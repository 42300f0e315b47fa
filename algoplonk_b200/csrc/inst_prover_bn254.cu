// Explicit instantiation: SRS handle + circuit/prover + curve-erased ops, Bn254.
#define B2P_INSTANTIATE_PROVER
#include "shard_group.cuh"
namespace b2p {
template struct Srs<Bn254>;
template struct Circuit<Bn254>;
const CurveOps* curve_ops_bn254() {
    static const CurveOpsImpl<Bn254> ops;
    return &ops;
}
}
