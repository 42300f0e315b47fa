// Test-only shim: exposes the product's field / curve templates (host emulation
// path of ptx.cuh) through a C ABI so pytest can check them against Python ints
// on a machine without a GPU.
#include <cstring>
#include "../../algoplonk_b200/csrc/field.cuh"
using namespace b2p;

// mul_sub exists for the fields with two spare bits (the base fields of the EC formulas)
template <class F> static F mulsub(const F& a, const F& b, const F& c, const F& d) {
    if constexpr (F::Params::BITS + 2 <= 32 * F::N) return F::mul_sub(a, b, c, d);
    else return a * b - c * d;
}
template <class F> static void binop(int op, const uint32_t* a, const uint32_t* b, uint32_t* o) {
    F x, y, r;
    memcpy(x.v, a, sizeof x.v); memcpy(y.v, b, sizeof y.v);
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.neg(); break;
        case 4: r = x.inverse(); break;
        case 5: r = x.to_mont(); break;
        case 6: r = x.from_mont(); break;
        case 7: r = x.sqr(); break;
        case 8: r = F::reduce_to_mont(x); break;
        case 9: r = mulsub<F>(x, y, y, x + F::one()); break;   // x*y - y*(x+1) = -y / R
        case 10: r = mulsub<F>(x, x, y, y); break;
        case 11: { F c; for (int i = 0; i < F::N; i++) c.v[i] = 0; /* (p-1)/3, plain limbs as a Montgomery value */
                   F pm1 = F::modulus(); pm1.v[0] -= 1; uint64_t rem = 0; for (int i = F::N - 1; i >= 0; i--) { uint64_t cur = (rem << 32) | pm1.v[i]; c.v[i] = (uint32_t)(cur / 3); rem = cur % 3; }
                   r = mulsub<F>(x, y, y, c); break; }
        default: r = F::zero();
    }
    memcpy(o, r.v, sizeof r.v);
}
extern "C" void ht_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* o) {
    switch (field) {
        case 0: binop<FrBn254>(op, a, b, o); break;
        case 1: binop<FpBn254>(op, a, b, o); break;
        case 2: binop<FrBls12381>(op, a, b, o); break;
        case 3: binop<FpBls12381>(op, a, b, o); break;
    }
}
