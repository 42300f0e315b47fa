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

// ---- the curve formulas of ec.cuh (what k_msm_accumulate and the bucket reductions execute) -----------------
#include "../../algoplonk_b200/csrc/ec.cuh"
template <class Fp> static void ecop(int op, const uint32_t* p1, const uint32_t* p2, uint64_t k, uint32_t* o) {
    Affine<Fp> a, b;
    memcpy(&a, p1, sizeof a); memcpy(&b, p2, sizeof b);     // Montgomery coordinates, (0,0) = infinity
    XYZZ<Fp> acc = XYZZ<Fp>::from_affine(a);
    switch (op) {
        case 0: acc.add(XYZZ<Fp>::from_affine(b)); break;                       // general addition
        case 1: if (!b.is_inf()) acc.add_affine(b.x, b.y); break;              // mixed addition (the hot formula)
        case 2: acc = acc.dbl(); break;
        case 3: acc.add_affine_signed(b, true); break;                          // a - b
        case 4: acc.add_affine_signed(b, false); break;                         // a + b, infinity-aware
        case 5: acc = acc.mul_small(k); break;
        case 6: acc = XYZZ<Fp>::dbl_affine(a.x, a.y); break;
        case 7: acc = acc.dbl(); acc.add(XYZZ<Fp>::from_affine(a)); acc.add(acc.neg()); break;   // 3a - 3a
        case 8: { XYZZ<Fp> t = acc.dbl(); t.add(acc); acc = t; acc.add(XYZZ<Fp>::from_affine(b).dbl()); break; }  // 3a + 2b
    }
    Affine<Fp> r = acc.to_affine();
    memcpy(o, &r, sizeof r);
}
extern "C" void ht_ec_op(int curve, int op, const uint32_t* p1, const uint32_t* p2, uint64_t k, uint32_t* o) {
    if (curve == 0) ecop<FpBn254>(op, p1, p2, k, o);
    else ecop<FpBls12381>(op, p1, p2, k, o);
}

// ---- Fp2 (fp2.cuh): the coordinate field of G2, what MsmEngine<...G2> computes with ---------------------------
#include "../../algoplonk_b200/csrc/fp2.cuh"
template <class Fp> static void fp2op(int op, const uint32_t* a, const uint32_t* b, uint32_t* o) {
    Fp2<Fp> x, y, r;
    memcpy(x.v, a, sizeof x.v); memcpy(y.v, b, sizeof y.v);
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x.sqr(); break;
        case 2: r = x + y; break;
        case 3: r = x - y; break;
        case 4: r = x.neg(); break;
        case 5: r = x.inverse(); break;
        case 6: r = Fp2<Fp>::mul_sub(x, y, y, x + Fp2<Fp>::one()); break;      // x y - y (x + 1) = -y
        case 7: r = x.dbl(); break;
        default: r = Fp2<Fp>::zero();
    }
    memcpy(o, r.v, sizeof r.v);
}
extern "C" void ht_fp2_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t* o) {
    if (curve == 0) fp2op<FpBn254>(op, a, b, o);
    else fp2op<FpBls12381>(op, a, b, o);
}
