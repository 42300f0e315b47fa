// Quadratic extension Fp2 = Fp[u] / (u^2 + 1) (both BN254 and BLS12-381 build their G2 over it): the coordinate field
// of G2 points, with the interface ec.cuh / msm.cuh expect of a coordinate field (N words in v[], zero / one, + - *,
// sqr, dbl, neg, mul_sub, inverse, is_zero) so that XYZZ<Fp2<Fp>> and MsmEngine run unchanged over G2.
// Memory layout = gnark-crypto's E2: A0 || A1, each an Fp in Montgomery form.
#pragma once
#include "field.cuh"

namespace b2p {

template <class Fp>
struct Fp2 {
    static constexpr int N = 2 * Fp::N;
    uint32_t v[N];          // a0 limbs, then a1 limbs (one contiguous array: ld_field / st_field move it as uint4s)

    HD Fp a0() const { Fp r; for (int i = 0; i < Fp::N; i++) r.v[i] = v[i]; return r; }
    HD Fp a1() const { Fp r; for (int i = 0; i < Fp::N; i++) r.v[i] = v[Fp::N + i]; return r; }
    HD static Fp2 make(const Fp& c0, const Fp& c1) {
        Fp2 r;
#pragma unroll
        for (int i = 0; i < Fp::N; i++) { r.v[i] = c0.v[i]; r.v[Fp::N + i] = c1.v[i]; }
        return r;
    }
    HD static Fp2 zero() { return make(Fp::zero(), Fp::zero()); }
    HD static Fp2 one() { return make(Fp::one(), Fp::zero()); }
    HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i];
        return acc == 0;
    }
    HD bool operator==(const Fp2& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i] ^ o.v[i];
        return acc == 0;
    }
    HD bool operator!=(const Fp2& o) const { return !(*this == o); }
    HD friend Fp2 operator+(const Fp2& a, const Fp2& b) { return make(a.a0() + b.a0(), a.a1() + b.a1()); }
    HD friend Fp2 operator-(const Fp2& a, const Fp2& b) { return make(a.a0() - b.a0(), a.a1() - b.a1()); }
    HD Fp2 neg() const { return make(a0().neg(), a1().neg()); }
    HD Fp2 dbl() const { return make(a0().dbl(), a1().dbl()); }
    // (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u: two fused a b - c d (one reduction each)
    HD friend Fp2 operator*(const Fp2& a, const Fp2& b) {
        const Fp x0 = a.a0(), x1 = a.a1(), y0 = b.a0(), y1 = b.a1();
        return make(Fp::mul_sub(x0, y0, x1, y1), Fp::mul_sub(x0, y1, x1.neg(), y0));
    }
    // (a0 + a1)(a0 - a1) + 2 a0 a1 u
    HD Fp2 sqr() const {
        const Fp x0 = a0(), x1 = a1();
        return make((x0 + x1) * (x0 - x1), (x0 * x1).dbl());
    }
    HD static Fp2 mul_sub(const Fp2& a, const Fp2& b, const Fp2& c, const Fp2& d) { return a * b - c * d; }
    // 1 / (a0 + a1 u) = (a0 - a1 u) / (a0^2 + a1^2); zero maps to zero
    HDN Fp2 inverse() const {
        const Fp x0 = a0(), x1 = a1();
        const Fp ninv = (x0.sqr() + x1.sqr()).inverse();
        return make(x0 * ninv, (x1 * ninv).neg());
    }
};

}  // namespace b2p
