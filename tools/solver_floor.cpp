// Lower bound of the witness solver for the benchmark circuit (what is left on the CPU of
// (*CompiledCircuit).Verify, algoplonk.go:81-89, once plonk.Prove runs on the GPU).  The squaring chain is 2^20 - 2
// DEPENDENT field multiplications x_{i+1} = x_i^2: no solver, gnark's included, can finish it faster than one core does
// this loop (gnark adds constraint decoding and wire bookkeeping on top).  64-bit-limb Montgomery product, the host field
// of the library (csrc/pairing_host.hpp).   g++ -O3 -std=c++17 -I algoplonk_b200/csrc tools/solver_floor.cpp -o /tmp/solver_floor
#include <chrono>
#include <cstdio>
#include "pairing_host.hpp"
using namespace b2p::hp;
int main() {
    typedef Fe<b2p::Bn254FrParams> Fr;
    Fr x = Fr::from_u64(2);
    const int n = (1 << 20) - 2;
    double best = 1e9;
    for (int rep = 0; rep < 5; rep++) {
        Fr y = x;
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n; i++) y = y.sqr();
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < best) best = ms;
        if (y.is_zero()) printf("unexpected\n");
    }
    printf("{\"dependent_squarings\": %d, \"ms\": %.2f, \"ns_per_squaring\": %.1f}\n", n, best, best * 1e6 / n);
    return 0;
}
