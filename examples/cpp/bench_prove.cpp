// A compiled caller of the library with no Python in the loop: the benchmark's squaring chain built,
// compiled and proved through the C++ mirror of AlgoPlonk's API, timed three ways --
//   prove:        (*CompiledCircuit).Verify with the witness columns in PAGEABLE std::vectors (what a Go caller holds)
//   from_inputs:  the two circuit inputs only; the library solves (b2p_solver_*), L R O stay in HBM
//   reloaded:     the same after SaveKey + Compile(..., snapshot): the persisted-key path
// and prints one JSON line.  All three must give the same proof bytes.
//   bench_prove <BN254|BLS12_381> <log2 rows> <proofs> <snapshot path>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../algoplonk_b200/host/algoplonk.hpp"

namespace ap = algoplonk;

template <int CURVE>
static int run(ap::setup::Name name, int log2, int proofs, const char* snap) {
    using Fr = typename ap::ScalarField<CURVE>::Fr;
    ap::Builder<Fr> B;
    const uint64_t rows = 1ull << log2, m = rows - 2;             // 1 public row + m squarings + 1 equality row
    // y = x0^(2^m): the builder is eager, so the public value is known once the chain is built
    std::vector<Fr> chain(m + 1);
    chain[0] = ap::fr_from_u64<Fr>(2);
    for (uint64_t i = 0; i < m; i++) chain[i + 1] = chain[i] * chain[i];
    const uint32_t y = B.Public(chain[m]);
    uint32_t x = B.Secret(chain[0]);
    for (uint64_t i = 0; i < m; i++) x = B.Mul(x, x);
    B.AssertIsEqual(y, x);
    const std::vector<Fr> inputs = {chain[m], chain[0]};
    std::vector<Fr> blinding;
    for (int i = 1; i <= 9; i++) blinding.push_back(ap::fr_from_u64<Fr>(i));
    const Fr tau = ap::fr_from_u64<Fr>(0x1234567);

    auto seconds = [](auto&& fn) {
        const auto t0 = std::chrono::steady_clock::now();
        fn();
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    ap::CompiledCircuit<CURVE> cc;
    const double t_compile = seconds([&] { ap::Compile<CURVE>(cc, B.cs, name, &tau); });
    std::vector<uint8_t> first = cc.Verify(blinding).raw;          // warm-up + the self-check (plonk.Verify) once
    bool same = true;
    const double t_prove = seconds([&] { for (int i = 0; i < proofs; i++) same &= cc.ProveOnly(blinding).raw == first; });
    same &= cc.VerifyFromInputs(inputs, blinding).raw == first;    // creates the solver, verifies once
    const double t_inputs = seconds([&] { for (int i = 0; i < proofs; i++) same &= cc.VerifyFromInputs(inputs, blinding, false).raw == first; });
    cc.SaveKey(snap);
    ap::CompiledCircuit<CURVE> cc2;
    const double t_reload = seconds([&] { ap::Compile<CURVE>(cc2, B.cs, name, &tau, nullptr, 0, snap); });
    same &= cc2.VerifyFromInputs(inputs, blinding).raw == first;
    bool rejected = false;
    try {
        std::vector<Fr> bad = inputs;
        bad[1] = bad[1] + Fr::one();
        cc.VerifyFromInputs(bad, blinding);
    } catch (const ap::Error& e) { rejected = std::string(e.what()).find("not satisfied") != std::string::npos; }
    printf("{\"curve\": \"%s\", \"log2\": %d, \"proofs\": %d, \"same_bytes\": %s, \"bad_input_rejected\": %s, "
           "\"compile_s\": %.3f, \"reload_compile_s\": %.3f, \"prove_ms\": %.3f, \"from_inputs_ms\": %.3f, "
           "\"what\": \"one lane, blocking calls from a compiled caller; pageable host columns / circuit inputs only\"}\n",
           CURVE == B2P_BN254 ? "BN254" : "BLS12_381", log2, proofs, same ? "true" : "false", rejected ? "true" : "false",
           t_compile, t_reload, 1e3 * t_prove / proofs, 1e3 * t_inputs / proofs);
    return same && rejected ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: bench_prove <BN254|BLS12_381> <log2 rows> <proofs> <snapshot path>\n"); return 2; }
    try {
        const int log2 = atoi(argv[2]), proofs = atoi(argv[3]);
        if (log2 < 3 || log2 > 24 || proofs < 1) { fprintf(stderr, "bad size\n"); return 2; }
        if (std::string(argv[1]) == "BN254") return run<B2P_BN254>(ap::setup::Name::TestOnlyBN254, log2, proofs, argv[4]);
        return run<B2P_BLS12_381>(ap::setup::Name::TestOnlyBLS12381, log2, proofs, argv[4]);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
