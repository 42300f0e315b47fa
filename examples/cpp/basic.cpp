// examples/basic of the reference (examples/basic/logicsigVerifier/main.go:30-52: a*a + b*b == c*c, a and b
// public) through the C++ mirror of AlgoPlonk's API: Compile -> Verify -> MarshalProof, printed as hex.
//   basic <BN254|BLS12_381> <tau hex> <9 blinding scalars hex...>
// tests/test_gpu_host_mirror.py builds it with g++ and compares the output with the committed golden proof.
#include <cstdio>
#include <string>
#include <vector>
#include "../../algoplonk_b200/host/algoplonk.hpp"

namespace ap = algoplonk;

template <int CURVE>
static int run(ap::setup::Name name, char** argv) {
    using Fr = typename ap::ScalarField<CURVE>::Fr;
    ap::Builder<Fr> B;
    const uint32_t a = B.Public(ap::fr_from_u64<Fr>(3)), b = B.Public(ap::fr_from_u64<Fr>(4));
    const uint32_t c = B.Secret(ap::fr_from_u64<Fr>(5));
    const uint32_t aa = B.Mul(a, a), bb = B.Mul(b, b), cc = B.Mul(c, c);
    B.AssertIsEqual(B.Add(aa, bb), cc);

    const Fr tau = ap::fr_from_hex<Fr>(argv[2]);
    std::vector<Fr> blinding;
    for (int i = 0; i < 9; i++) blinding.push_back(ap::fr_from_hex<Fr>(argv[3 + i]));

    ap::CompiledCircuit<CURVE> circuit;
    ap::Compile<CURVE>(circuit, B.cs, name, &tau);
    const auto vp = circuit.Verify(blinding);
    for (uint8_t x : vp.MarshalProof()) printf("%02x", x);
    printf("\n");
    for (uint8_t x : vp.MarshalPublicInputs()) printf("%02x", x);
    printf("\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 12) { fprintf(stderr, "usage: basic <BN254|BLS12_381> <tau hex> <9 blinding hex>\n"); return 2; }
    try {
        if (std::string(argv[1]) == "BN254") return run<B2P_BN254>(ap::setup::Name::TestOnlyBN254, argv);
        return run<B2P_BLS12_381>(ap::setup::Name::TestOnlyBLS12381, argv);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
