// Test-only program: identities inside csrc/pairing_host.hpp that the C ABI does not expose -- the Granger-Scott
// cyclotomic squaring equals the plain squaring on the cyclotomic subgroup (and differs off it), inverse = conjugate
// there.  Built and run by tests/test_verify_host.py; exit code = number of failed checks.
#define HD inline
#include <cstdio>
#include "../../algoplonk_b200/csrc/pairing_host.hpp"
using namespace b2p::hp;
template <class PC> int run() {
    using PR = Pairing<PC>; using Fp = typename PC::Fp; using E12 = typename PR::E12;
    E12 f;
    for (int i = 0; i < 6; i++) f.c[i] = {Fp::from_u64(1234567 * (i + 1) + 11), Fp::from_u64(7654321 * (i + 3) + 5)};
    E12 g = f.conj() * f.inverse();
    g = PR::frobenius(PR::frobenius(g)) * g;          // in the cyclotomic subgroup
    int bad = 0;
    for (int k = 0; k < 5; k++) {
        bad += !(g.cyclotomic_sqr() == g.sqr());
        g = g.sqr() * g;
    }
    bad += (f.cyclotomic_sqr() == f.sqr());           // and NOT valid outside it
    bad += !((g * g.conj()).is_one());                // inverse = conjugate in the subgroup
    printf("bad=%d\n", bad);
    return bad;
}
int main() { return run<Bn254Pairing>() + run<Bls12381Pairing>(); }
