// Test-only shim: the reduced-radix field of field29.cuh (what the MSM accumulation kernel computes in) behind a C
// ABI, compiled for the host -- field29.cuh is plain C++ on 32/64-bit integers, so this is the very code the device runs.
#include <cstring>
#include "field29.cuh"
#include "../../algoplonk_b200/csrc/ec.cuh"
using namespace b2p;

// a, b, c, d: N32 32-bit words each (plain integers below 2^(32 N32)); o: the result as a plain integer (canonical)
template <class F> static void op29(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d,
                                    uint32_t* o) {
    typename F::Mem A, B, Cc, D, R;
    memcpy(A.v, a, sizeof A.v); memcpy(B.v, b, sizeof B.v); memcpy(Cc.v, c, sizeof Cc.v); memcpy(D.v, d, sizeof D.v);
    const F x = F::from_words(A), y = F::from_words(B), z = F::from_words(Cc), w = F::from_words(D);
    F r = F::zero();
    switch (op) {
        case 0: r = x * y; break;                                       // x y / R'
        case 1: r = x.sqr(); break;
        case 2: r = add(x, y); break;
        case 3: r = F::template sub<2>(x, y); break;                    // x + 2p - y
        case 4: r = F::template sub<8>(x, y); break;
        case 5: r = F::mul_add(x, y, z, w); break;                      // (x y + z w) / R'
        case 6: r = x.template neg<2>(); break;
        case 7: R = F::from_mem(A).to_mem(); memcpy(o, R.v, sizeof R.v); return;       // round trip through R'
        case 8: r = F::from_mem(A); break;                              // x R'/R
        case 9: {                                                        // lazy chain: ((x-y)+(z-w))^2 with 16p offsets
            const F s = add(F::template sub<4>(x, y), F::template sub<4>(z, w));
            r = s.sqr();
            break;
        }
        case 10: o[0] = x.is_zero_mod_p() ? 1u : 0u; return;
        case 11: r = add(add(x.dbl().dbl(), y.dbl()), add(z, w)); break;   // 4x + 2y + z + w < 8p: reduce_full's range
    }
    R = r.to_words();
    memcpy(o, R.v, sizeof R.v);
}
extern "C" void h29_field_op(int field, int op, const uint32_t* a, const uint32_t* b, const uint32_t* c,
                             const uint32_t* d, uint32_t* o) {
    switch (field) {
        case 0: op29<Fr29Bn254>(op, a, b, c, d, o); break;
        case 1: op29<Fp29Bn254>(op, a, b, c, d, o); break;
        case 2: op29<Fr29Bls12381>(op, a, b, c, d, o); break;
        case 3: op29<Fp29Bls12381>(op, a, b, c, d, o); break;
    }
}

// acc (XYZZ, memory format: 4 Montgomery-R coordinates) += (+-) point (affine, memory format), computed in the
// reduced-radix domain, result back in memory format: must equal XYZZ<Fp>::add_affine_signed coordinate for coordinate
template <class F> static void madd29(const uint32_t* acc, const uint32_t* pt, int neg, uint32_t* o) {
    typedef typename F::Mem Fp;
    XYZZ<Fp> a;
    Affine<Fp> p;
    memcpy(&a, acc, sizeof a); memcpy(&p, pt, sizeof p);
    XYZZ29<F> r;
    if (a.is_inf()) r = XYZZ29<F>::inf();
    else { r.X = F::from_mem(a.X); r.Y = F::from_mem(a.Y); r.ZZ = F::from_mem(a.ZZ); r.ZZZ = F::from_mem(a.ZZZ); }
    if (!p.is_inf()) {
        const F x = F::from_mem(p.x);
        F y = F::from_mem(p.y);
        if (neg) y = y.template neg<2>();
        r.add_affine(x, y);
    }
    XYZZ<Fp> out;
    if (r.is_inf()) out = XYZZ<Fp>::inf();
    else { out.X = r.X.to_mem(); out.Y = r.Y.to_mem(); out.ZZ = r.ZZ.to_mem(); out.ZZZ = r.ZZZ.to_mem(); }
    memcpy(o, &out, sizeof out);
}
extern "C" void h29_madd(int curve, const uint32_t* acc, const uint32_t* pt, int neg, uint32_t* o) {
    if (curve == 0) madd29<Fp29Bn254>(acc, pt, neg, o);
    else madd29<Fp29Bls12381>(acc, pt, neg, o);
}
// the reference: the same operation with field.cuh / ec.cuh
template <class Fp> static void madd32(const uint32_t* acc, const uint32_t* pt, int neg, uint32_t* o) {
    XYZZ<Fp> a;
    Affine<Fp> p;
    memcpy(&a, acc, sizeof a); memcpy(&p, pt, sizeof p);
    a.add_affine_signed(p, neg != 0);
    if (a.is_inf()) a = XYZZ<Fp>::inf();
    memcpy(o, &a, sizeof a);
}
extern "C" void h29_madd_ref(int curve, const uint32_t* acc, const uint32_t* pt, int neg, uint32_t* o) {
    if (curve == 0) madd32<FpBn254>(acc, pt, neg, o);
    else madd32<FpBls12381>(acc, pt, neg, o);
}
