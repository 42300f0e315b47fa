// Explicit instantiation: SRS handle + circuit/prover + curve-erased ops, Bls12381.
#define B2P_INSTANTIATE_PROVER
#include "shard_group.cuh"
namespace b2p {
template struct Srs<Bls12381>;
template struct Circuit<Bls12381>;
const CurveOps* curve_ops_bls12381() {
    static const CurveOpsImpl<Bls12381> ops;
    return &ops;
}
}
