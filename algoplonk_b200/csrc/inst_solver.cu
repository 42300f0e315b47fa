// Explicit instantiation: the witness solver over both scalar fields (solver.cuh).
#include "solver.cuh"
namespace b2p {

template <class Fr>
static SolverBase* make_solver(int curve, uint64_t n, uint32_t nb_public, uint64_t nb_variables, const uint32_t* input_ids,
                               uint32_t nb_inputs, const void* const cols[5], const uint32_t* xa, const uint32_t* xb,
                               const uint32_t* xc, const b2p_hint* hints, uint32_t n_hints, const uint8_t* unchecked) {
    Solver<Fr>* s = new Solver<Fr>();
    s->curve = curve;
    try { s->create(n, nb_public, nb_variables, input_ids, nb_inputs, cols, xa, xb, xc, hints, n_hints, unchecked); } catch (...) { delete s; throw; }
    return s;
}
SolverBase* new_solver(int curve, uint64_t n, uint32_t nb_public, uint64_t nb_variables, const uint32_t* input_ids,
                       uint32_t nb_inputs, const void* const cols[5], const uint32_t* xa, const uint32_t* xb,
                       const uint32_t* xc, const b2p_hint* hints, uint32_t n_hints, const uint8_t* unchecked) {
    if (curve == B2P_BN254) return make_solver<FrBn254>(curve, n, nb_public, nb_variables, input_ids, nb_inputs, cols, xa, xb, xc, hints, n_hints, unchecked);
    if (curve == B2P_BLS12_381) return make_solver<FrBls12381>(curve, n, nb_public, nb_variables, input_ids, nb_inputs, cols, xa, xb, xc, hints, n_hints, unchecked);
    throw Error(B2P_ERR_ARG, "unsupported curve id (B2P_BN254 = 0, B2P_BLS12_381 = 1)");
}

}  // namespace b2p
