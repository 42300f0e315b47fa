// Host-side pairing check for plonk.Verify (the step right after plonk.Prove in (*CompiledCircuit).Verify,
// /root/reference/algoplonk.go:93; gnark runs it on the CPU too).  Decides
//     prod_i e(P_i, Q_i) == 1        P_i in G1, Q_i in G2 (on the sextic twist over Fp2)
// for BN254 and BLS12-381, which is all a KZG / PLONK verifier needs: any non-degenerate bilinear map gives the
// same answer, so no convention has to match gnark's GT values.
//
// Construction (written for this library, 64-bit limbs, no device code):
//   Fp      Montgomery CIOS multiplication; the radix 2^(32 N32) of the device fields equals 2^(64 N64), so
//           gnark's in-memory elements are read with a memcpy
//   Fp2     Fp[u]/(u^2+1)            (both curves)
//   Fp12    Fp2[w]/(w^6 - xi), stored flat (coefficients of w^0..w^5), xi = 9+u (BN254) / 1+u (BLS12-381);
//           products, squarings and the inversion go through the tower view Fp6[w]/(w^2 - v), Fp6 = Fp2[v]/(v^3 - xi),
//           v = w^2 (Karatsuba: 18 / 12 products in Fp2); line functions multiply in sparsely on the flat form
//   Miller  ate pairing f_{T,Q}(P) with T = t-1: 6x^2 (BN254, 127 bits; no Frobenius correction lines needed
//           with this loop length) and |x| (BLS12-381); affine doubling/addition on the twist, line
//           coefficients (slope, slope*x_T - y_T) computed once per G2 point and cached (a verifying key has two)
//   twist   BN254 is a D-type twist  (x', y') -> (x' w^2, y' w^3):  l(P) = yP - slope*xP w + (slope*xT - yT) w^3
//           BLS12-381 is M-type      (x', y') -> (x'/w^2, y'/w^3):  l(P) w^3 = (slope*xT - yT) - slope*xP w^2 + yP w^3
//           (w^3 lies in Fp4, a proper subfield: the final exponentiation removes it)
//   final   easy part (p^6-1)(p^2+1); hard part by the curve's x-chain:
//           BN254      (p^4-p^2+1)/r = l0 + l1 p + l2 p^2 + p^3,  l0 = -2-18x-30x^2-36x^3,
//                      l1 = 1-12x-18x^2-36x^3, l2 = 1+6x^2
//           BLS12-381  3 (p^4-p^2+1)/r = (x-1)^2 (x+p)(x^2+p^2-1) + 3
//           (both are checked as integer identities in tests/test_verify_host.py; 3 is prime to r, so "== 1"
//           is unchanged)
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#if defined(__x86_64__)
#include <x86intrin.h>
#endif

#include "field_params.cuh"

namespace b2p {
namespace hp {

typedef unsigned __int128 u128;

// add / subtract with carry: the x86-64 intrinsics keep the chain in the flags register
#if defined(__x86_64__)
#define B2P_ADC(c, a, b, out) _addcarry_u64((c), (a), (b), reinterpret_cast<unsigned long long*>(out))
#define B2P_SBB(c, a, b, out) _subborrow_u64((c), (a), (b), reinterpret_cast<unsigned long long*>(out))
#else
static inline unsigned char b2p_adc(unsigned char c, uint64_t a, uint64_t b, uint64_t* out) {
    u128 s = (u128)a + b + c;
    *out = (uint64_t)s;
    return (unsigned char)(s >> 64);
}
static inline unsigned char b2p_sbb(unsigned char c, uint64_t a, uint64_t b, uint64_t* out) {
    u128 d = (u128)a - b - c;
    *out = (uint64_t)d;
    return (unsigned char)((d >> 64) & 1);
}
#define B2P_ADC(c, a, b, out) b2p_adc((c), (a), (b), (out))
#define B2P_SBB(c, a, b, out) b2p_sbb((c), (a), (b), (out))
#endif

// ---------------------------------------------------------------------------------------------------------
// prime field, N 64-bit limbs, Montgomery form
// ---------------------------------------------------------------------------------------------------------
template <class P32>
struct Fe {
    static constexpr int N = P32::N / 2;
    uint64_t v[N];

    static constexpr uint64_t M(int i) { return (uint64_t)P32::mod_[2 * i] | ((uint64_t)P32::mod_[2 * i + 1] << 32); }
    static constexpr uint64_t inv64() {   // -p^-1 mod 2^64 by Newton iteration from the 32-bit constant
        uint64_t p0 = M(0), y = 1;
        for (int i = 0; i < 6; i++) y = y * (2 - p0 * y);
        return ~y + 1;
    }
    static Fe zero() { Fe r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
    static Fe one() {
        Fe r;
        for (int i = 0; i < N; i++) r.v[i] = (uint64_t)P32::one_[2 * i] | ((uint64_t)P32::one_[2 * i + 1] << 32);
        return r;
    }
    static Fe r2() {
        Fe r;
        for (int i = 0; i < N; i++) r.v[i] = (uint64_t)P32::r2_[2 * i] | ((uint64_t)P32::r2_[2 * i + 1] << 32);
        return r;
    }
    static Fe from_u64(uint64_t x) { Fe r = zero(); r.v[0] = x; return mul(r, r2()); }
    // gnark fp.Element / fr.Element memory (little-endian limbs, Montgomery form)
    static Fe load(const void* p) { Fe r; memcpy(r.v, p, sizeof r.v); return r; }
    void store(void* p) const { memcpy(p, v, sizeof v); }
    bool is_zero() const { uint64_t a = 0; for (int i = 0; i < N; i++) a |= v[i]; return a == 0; }
    bool operator==(const Fe& o) const { uint64_t a = 0; for (int i = 0; i < N; i++) a |= v[i] ^ o.v[i]; return a == 0; }
    bool operator!=(const Fe& o) const { return !(*this == o); }
    static bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > M(i)) return true;
            if (a[i] < M(i)) return false;
        }
        return true;
    }
    // r = t - p if t >= p (or `force`), else t
    static inline void cond_sub(uint64_t* t, uint64_t force) {
        uint64_t d[N];
        unsigned char br = 0;
        for (int i = 0; i < N; i++) br = B2P_SBB(br, t[i], M(i), &d[i]);
        if (!br || force)
            for (int i = 0; i < N; i++) t[i] = d[i];
    }
    friend Fe operator+(const Fe& a, const Fe& b) {   // both moduli leave spare top bits: no carry out of limb N-1
        Fe r;
        unsigned char c = 0;
        for (int i = 0; i < N; i++) c = B2P_ADC(c, a.v[i], b.v[i], &r.v[i]);
        cond_sub(r.v, 0);
        return r;
    }
    friend Fe operator-(const Fe& a, const Fe& b) {
        Fe r;
        unsigned char br = 0;
        for (int i = 0; i < N; i++) br = B2P_SBB(br, a.v[i], b.v[i], &r.v[i]);
        if (br) {
            unsigned char c = 0;
            for (int i = 0; i < N; i++) c = B2P_ADC(c, r.v[i], M(i), &r.v[i]);
        }
        return r;
    }
    Fe neg() const { return is_zero() ? *this : zero() - *this; }
    Fe dbl() const { return *this + *this; }
    static Fe mul(const Fe& a, const Fe& b) {
        constexpr uint64_t INV = inv64();
        uint64_t t[N + 2];
        for (int i = 0; i < N + 2; i++) t[i] = 0;
        for (int i = 0; i < N; i++) {
            u128 c = 0;
            for (int j = 0; j < N; j++) {
                c += (u128)a.v[j] * b.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N] = (uint64_t)c;
            t[N + 1] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * INV;
            c = ((u128)m * M(0) + t[0]) >> 64;
            for (int j = 1; j < N; j++) {
                c += (u128)m * M(j) + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N - 1] = (uint64_t)c;
            t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        Fe r;
        cond_sub(t, t[N]);
        for (int i = 0; i < N; i++) r.v[i] = t[i];
        return r;
    }
    friend Fe operator*(const Fe& a, const Fe& b) { return mul(a, b); }
    Fe sqr() const { return mul(*this, *this); }
    Fe from_mont() const { Fe o = zero(); o.v[0] = 1; return mul(*this, o); }
    // exponent: little-endian 64-bit limbs
    Fe pow(const uint64_t* e, int limbs) const {
        Fe acc = one();
        for (int i = limbs * 64 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
        }
        return acc;
    }
    Fe inverse() const {   // Fermat; 0 -> 0
        uint64_t e[N];
        for (int i = 0; i < N; i++) e[i] = (uint64_t)P32::pm2_[2 * i] | ((uint64_t)P32::pm2_[2 * i + 1] << 32);
        return pow(e, N);
    }
    Fe mul_small(unsigned k) const {   // k * this by additions
        Fe acc = zero(), b = *this;
        while (k) {
            if (k & 1) acc = acc + b;
            b = b.dbl();
            k >>= 1;
        }
        return acc;
    }
};

// ---------------------------------------------------------------------------------------------------------
// Fp2 = Fp[u]/(u^2 + 1); XI0: the non-residue of the sextic extension is xi = XI0 + u
// ---------------------------------------------------------------------------------------------------------
template <class F, unsigned XI0>
struct Fp2T {
    F a, b;   // a + b u
    static Fp2T zero() { return {F::zero(), F::zero()}; }
    static Fp2T one() { return {F::one(), F::zero()}; }
    bool is_zero() const { return a.is_zero() && b.is_zero(); }
    bool operator==(const Fp2T& o) const { return a == o.a && b == o.b; }
    friend Fp2T operator+(const Fp2T& x, const Fp2T& y) { return {x.a + y.a, x.b + y.b}; }
    friend Fp2T operator-(const Fp2T& x, const Fp2T& y) { return {x.a - y.a, x.b - y.b}; }
    friend Fp2T operator*(const Fp2T& x, const Fp2T& y) {
        F t0 = x.a * y.a, t1 = x.b * y.b, t2 = (x.a + x.b) * (y.a + y.b);
        return {t0 - t1, t2 - t0 - t1};
    }
    Fp2T sqr() const { F t = a * b; return {(a + b) * (a - b), t.dbl()}; }
    Fp2T neg() const { return {a.neg(), b.neg()}; }
    Fp2T dbl() const { return {a.dbl(), b.dbl()}; }
    Fp2T conj() const { return {a, b.neg()}; }
    Fp2T scale(const F& k) const { return {a * k, b * k}; }
    Fp2T mul_xi() const { return {a.mul_small(XI0) - b, b.mul_small(XI0) + a}; }
    Fp2T inverse() const {
        F n = (a.sqr() + b.sqr()).inverse();
        return {a * n, (b * n).neg()};
    }
    Fp2T pow(const uint64_t* e, int limbs) const {
        Fp2T acc = one();
        for (int i = limbs * 64 - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * *this;
        }
        return acc;
    }
};

// Fp6 = Fp2[v]/(v^3 - xi): what the tower view of Fp12 needs
template <class E2>
struct Fp6T {
    E2 c[3];
    friend Fp6T operator*(const Fp6T& x, const Fp6T& y) {   // Karatsuba: 6 products in Fp2
        E2 v0 = x.c[0] * y.c[0], v1 = x.c[1] * y.c[1], v2 = x.c[2] * y.c[2];
        Fp6T r;
        r.c[0] = v0 + ((x.c[1] + x.c[2]) * (y.c[1] + y.c[2]) - v1 - v2).mul_xi();
        r.c[1] = (x.c[0] + x.c[1]) * (y.c[0] + y.c[1]) - v0 - v1 + v2.mul_xi();
        r.c[2] = (x.c[0] + x.c[2]) * (y.c[0] + y.c[2]) - v0 - v2 + v1;
        return r;
    }
    friend Fp6T operator+(const Fp6T& x, const Fp6T& y) { return {{x.c[0] + y.c[0], x.c[1] + y.c[1], x.c[2] + y.c[2]}}; }
    friend Fp6T operator-(const Fp6T& x, const Fp6T& y) { return {{x.c[0] - y.c[0], x.c[1] - y.c[1], x.c[2] - y.c[2]}}; }
    Fp6T mul_v() const { return {{c[2].mul_xi(), c[0], c[1]}}; }
    Fp6T neg() const { return {{c[0].neg(), c[1].neg(), c[2].neg()}}; }
    Fp6T inverse() const {
        E2 t0 = c[0].sqr() - (c[1] * c[2]).mul_xi();
        E2 t1 = c[2].sqr().mul_xi() - c[0] * c[1];
        E2 t2 = c[1].sqr() - c[0] * c[2];
        E2 d = (c[0] * t0 + (c[2] * t1 + c[1] * t2).mul_xi()).inverse();
        return {{t0 * d, t1 * d, t2 * d}};
    }
};

// Fp12 = Fp2[w]/(w^6 - xi), coefficients of w^0..w^5
template <class E2>
struct Fp12T {
    E2 c[6];
    static Fp12T one() {
        Fp12T r;
        r.c[0] = E2::one();
        for (int i = 1; i < 6; i++) r.c[i] = E2::zero();
        return r;
    }
    bool is_one() const {
        if (!(c[0] == E2::one())) return false;
        for (int i = 1; i < 6; i++) if (!c[i].is_zero()) return false;
        return true;
    }
    bool operator==(const Fp12T& o) const {
        for (int i = 0; i < 6; i++) if (!(c[i] == o.c[i])) return false;
        return true;
    }
    // the tower view Fp12 = Fp6[w]/(w^2 - v): A = (c0, c2, c4), B = (c1, c3, c5)
    using E6 = Fp6T<E2>;
    E6 even() const { return {{c[0], c[2], c[4]}}; }
    E6 odd() const { return {{c[1], c[3], c[5]}}; }
    static Fp12T from_tower(const E6& a, const E6& b) {
        Fp12T r;
        r.c[0] = a.c[0]; r.c[2] = a.c[1]; r.c[4] = a.c[2];
        r.c[1] = b.c[0]; r.c[3] = b.c[1]; r.c[5] = b.c[2];
        return r;
    }
    friend Fp12T operator*(const Fp12T& x, const Fp12T& y) {   // 3 products in Fp6
        E6 A = x.even(), B = x.odd(), C = y.even(), D = y.odd();
        E6 ac = A * C, bd = B * D;
        return from_tower(ac + bd.mul_v(), (A + B) * (C + D) - ac - bd);
    }
    Fp12T sqr() const {   // 2 products in Fp6
        E6 A = even(), B = odd();
        E6 ab = A * B;
        return from_tower((A + B) * (A + B.mul_v()) - ab - ab.mul_v(), ab + ab);
    }
    // this * (l0 w^p0 + l1 w^p1 + l2 w^p2)
    Fp12T mul_sparse(const int* pos, const E2* l) const {
        E2 lo[6], hi[6];
        for (int i = 0; i < 6; i++) lo[i] = hi[i] = E2::zero();
        for (int k = 0; k < 3; k++)
            for (int i = 0; i < 6; i++) {
                E2 t = c[i] * l[k];
                const int d = i + pos[k];
                if (d < 6) lo[d] = lo[d] + t;
                else hi[d - 6] = hi[d - 6] + t;
            }
        Fp12T r;
        for (int i = 0; i < 6; i++) r.c[i] = lo[i] + hi[i].mul_xi();
        return r;
    }
    // x -> x^(p^6): w^(p^6) = -w
    Fp12T conj() const {
        Fp12T r = *this;
        r.c[1] = r.c[1].neg(); r.c[3] = r.c[3].neg(); r.c[5] = r.c[5].neg();
        return r;
    }
    Fp12T inverse() const {   // (A + B w)^-1 = (A - B w) / (A^2 - v B^2)
        E6 A = even(), B = odd();
        E6 d = (A * A - (B * B).mul_v()).inverse();
        return from_tower(A * d, (B * d).neg());
    }
    // Squaring of an element of the cyclotomic subgroup (norm 1 over Fp6: everything after the easy part of the
    // final exponentiation).  Granger-Scott: with Fp12 = Fp4[.]^3 the square costs three squarings in
    // Fp4 = Fp2[y]/(y^2 - xi), i.e. 6 products in Fp2 instead of 12.  In the tower view (g0 g1 g2) + (g3 g4 g5) w
    // the three Fp4 elements are (g0, g4), (g3, g2), (g1, g5).
    Fp12T cyclotomic_sqr() const {
        const E2 &g0 = c[0], &g1 = c[2], &g2 = c[4], &g3 = c[1], &g4 = c[3], &g5 = c[5];
        auto fp4_sqr = [](const E2& a, const E2& b, E2* r0, E2* r1) {   // (a + b y)^2 = (a^2 + xi b^2) + 2ab y
            E2 ab = a * b;
            *r0 = (a + b) * (a + b.mul_xi()) - ab - ab.mul_xi();
            *r1 = ab.dbl();
        };
        E2 t0, t1, t2, t3, t4, t5;
        fp4_sqr(g0, g4, &t0, &t1);
        fp4_sqr(g3, g2, &t2, &t3);
        fp4_sqr(g1, g5, &t4, &t5);
        auto three_minus_two = [](const E2& t, const E2& z) { E2 d = t - z; return d.dbl() + t; };   // 3t - 2z
        auto three_plus_two = [](const E2& t, const E2& z) { E2 d = t + z; return d.dbl() + t; };    // 3t + 2z
        Fp12T r;
        r.c[0] = three_minus_two(t0, g0);             // g0'
        r.c[3] = three_plus_two(t1, g4);              // g4'
        r.c[1] = three_plus_two(t5.mul_xi(), g3);     // g3'
        r.c[4] = three_minus_two(t4, g2);             // g2'
        r.c[2] = three_minus_two(t2, g1);             // g1'
        r.c[5] = three_plus_two(t3, g5);              // g5'
        return r;
    }
    Fp12T pow_u64(uint64_t e, bool cyclotomic = false) const {
        if (e == 0) return one();
        int top = 63;
        while (!((e >> top) & 1)) top--;
        Fp12T acc = *this;
        for (int i = top - 1; i >= 0; i--) {
            acc = cyclotomic ? acc.cyclotomic_sqr() : acc.sqr();
            if ((e >> i) & 1) acc = acc * *this;
        }
        return acc;
    }
};

// ---------------------------------------------------------------------------------------------------------
// curve descriptions
// ---------------------------------------------------------------------------------------------------------
struct Bn254Pairing {
    using Fp = Fe<Bn254FpParams>;
    using Fr = Fe<Bn254FrParams>;
    using FrP = Bn254FrParams;
    static constexpr unsigned XI0 = 9;
    static constexpr bool D_TWIST = true;
    static constexpr uint64_t X = 4965661367192848881ull;   // curve parameter, positive
    static constexpr bool X_NEG = false;
    static u128 loop_count() { return (u128)6 * X * X; }    // t - 1
    static constexpr unsigned B = 3;
    // generator of G2 (canonical decimal strings would be long: big-endian hex, x.A0 x.A1 y.A0 y.A1)
    static const char* const* g2_hex() {
        static const char* const h[4] = {
            "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed",
            "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2",
            "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa",
            "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"};
        return h;
    }
};
struct Bls12381Pairing {
    using Fp = Fe<Bls12381FpParams>;
    using Fr = Fe<Bls12381FrParams>;
    using FrP = Bls12381FrParams;
    static constexpr unsigned XI0 = 1;
    static constexpr bool D_TWIST = false;
    static constexpr uint64_t X = 0xd201000000010000ull;    // |x|, x is negative
    static constexpr bool X_NEG = true;
    static u128 loop_count() { return (u128)X; }
    static constexpr unsigned B = 4;
    static const char* const* g2_hex() {
        static const char* const h[4] = {
            "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8",
            "13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e",
            "0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801",
            "0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"};
        return h;
    }
};

template <class F>
inline F fe_from_hex(const char* h) {   // canonical big-endian hex -> Montgomery
    F r = F::zero();
    const size_t len = strlen(h);
    for (size_t i = 0; i < len; i++) {
        const char ch = h[len - 1 - i];
        const uint64_t d = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
        r.v[i / 16] |= d << (4 * (i % 16));
    }
    return F::mul(r, F::r2());
}

template <class PC>
struct Pairing {
    using Fp = typename PC::Fp;
    using E2 = Fp2T<Fp, PC::XI0>;
    using E12 = Fp12T<E2>;
    static constexpr int FPB = Fp::N * 8;   // bytes of an Fp element in memory

    struct G1 { Fp x, y; bool inf; };
    struct G2 { E2 x, y; bool inf; };
    struct Line { E2 slope, c; };           // c = slope * x_T - y_T
    struct Prepared { bool inf; std::vector<Line> lines; };

    // gnark G1Affine / G2Affine memory: X | Y (G2: X.A0 X.A1 Y.A0 Y.A1), Montgomery; all-zero = infinity
    static G1 load_g1(const uint8_t* p) {
        G1 r{Fp::load(p), Fp::load(p + FPB), false};
        r.inf = r.x.is_zero() && r.y.is_zero();
        return r;
    }
    static G2 load_g2(const uint8_t* p) {
        G2 r{{Fp::load(p), Fp::load(p + FPB)}, {Fp::load(p + 2 * FPB), Fp::load(p + 3 * FPB)}, false};
        r.inf = r.x.is_zero() && r.y.is_zero();
        return r;
    }
    static void store_g2(const G2& q, uint8_t* p) {
        if (q.inf) { memset(p, 0, 4 * FPB); return; }
        q.x.a.store(p); q.x.b.store(p + FPB); q.y.a.store(p + 2 * FPB); q.y.b.store(p + 3 * FPB);
    }
    static E2 twist_b() {   // D-type: b / xi, M-type: b * xi
        E2 b = {Fp::from_u64(PC::B), Fp::zero()};
        E2 xi = {Fp::from_u64(PC::XI0), Fp::one()};
        return PC::D_TWIST ? b * xi.inverse() : b * xi;
    }
    static bool g1_on_curve(const G1& p) {
        if (p.inf) return true;
        return p.y.sqr() == p.x.sqr() * p.x + Fp::from_u64(PC::B);
    }
    static bool g2_on_curve(const G2& q) {
        if (q.inf) return true;
        return q.y.sqr() == q.x.sqr() * q.x + twist_b();
    }
    static G2 g2_generator() {
        const char* const* h = PC::g2_hex();
        return {{fe_from_hex<Fp>(h[0]), fe_from_hex<Fp>(h[1])}, {fe_from_hex<Fp>(h[2]), fe_from_hex<Fp>(h[3])}, false};
    }
    // affine group law on the twist; `slope` receives the slope of the chord / tangent used
    static G2 g2_double(const G2& t, E2* slope) {
        if (t.inf || t.y.is_zero()) { if (slope) *slope = E2::zero(); return {E2::zero(), E2::zero(), true}; }
        E2 x2 = t.x.sqr();
        E2 l = (x2.dbl() + x2) * t.y.dbl().inverse();
        E2 x3 = l.sqr() - t.x.dbl();
        E2 y3 = l * (t.x - x3) - t.y;
        if (slope) *slope = l;
        return {x3, y3, false};
    }
    static G2 g2_add(const G2& t, const G2& q, E2* slope, bool* vertical) {
        if (vertical) *vertical = false;
        if (t.inf) { if (vertical) *vertical = true; return q; }
        if (q.inf) { if (vertical) *vertical = true; return t; }
        if (t.x == q.x) {
            if (t.y == q.y) return g2_double(t, slope);
            if (vertical) *vertical = true;
            return {E2::zero(), E2::zero(), true};
        }
        E2 l = (q.y - t.y) * (q.x - t.x).inverse();
        E2 x3 = l.sqr() - t.x - q.x;
        E2 y3 = l * (t.x - x3) - t.y;
        if (slope) *slope = l;
        return {x3, y3, false};
    }
    // k * Q, k: canonical little-endian limbs
    static G2 g2_mul(const G2& q, const uint64_t* k, int limbs) {
        G2 acc{E2::zero(), E2::zero(), true};
        for (int i = limbs * 64 - 1; i >= 0; i--) {
            acc = g2_double(acc, nullptr);
            if ((k[i >> 6] >> (i & 63)) & 1) acc = g2_add(acc, q, nullptr, nullptr);
        }
        return acc;
    }

    // ---- gnark's compressed encodings (kzg.VerifyingKey.ReadFrom on setup/<name>/vk.bin, setup/setup.go:174,190) ----
    // Both base fields are 3 mod 4: sqrt(a) = a^((p+1)/4) when it exists.
    static bool fp_sqrt(const Fp& a, Fp* out) {
        uint64_t e[Fp::N];
        uint64_t carry = 1;
        for (int i = 0; i < Fp::N; i++) {           // p + 1
            u128 s = (u128)Fp::M(i) + carry;
            e[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
        for (int i = 0; i < Fp::N; i++) e[i] = (e[i] >> 2) | (i + 1 < Fp::N ? e[i + 1] << 62 : 0);
        Fp s = a.pow(e, Fp::N);
        if (!(s.sqr() == a)) return false;
        *out = s;
        return true;
    }
    // through the norm: x0^2 = (a0 +- sqrt(a0^2 + a1^2)) / 2, x1 = a1 / (2 x0)
    static bool e2_sqrt(const E2& a, E2* out) {
        if (a.b.is_zero()) {
            Fp s;
            if (fp_sqrt(a.a, &s)) { *out = {s, Fp::zero()}; return true; }
            if (fp_sqrt(a.a.neg(), &s)) { *out = {Fp::zero(), s}; return true; }   // -1 is a non-residue
            return false;
        }
        Fp n;
        if (!fp_sqrt(a.a.sqr() + a.b.sqr(), &n)) return false;
        const Fp half = Fp::from_u64(2).inverse();
        const Fp cands[2] = {(a.a + n) * half, (a.a - n) * half};
        for (const Fp& cand : cands) {
            Fp x0;
            if (!fp_sqrt(cand, &x0) || x0.is_zero()) continue;
            E2 x = {x0, a.b * x0.dbl().inverse()};
            if (x.sqr() == a) { *out = x; return true; }
        }
        return false;
    }
    static bool lex_largest(const Fp& mont) {       // canonical value > (p - 1) / 2
        const Fp c = mont.from_mont();
        for (int i = Fp::N - 1; i >= 0; i--) {
            const uint64_t h = (Fp::M(i) >> 1) | (i + 1 < Fp::N ? Fp::M(i + 1) << 63 : 0);   // (p - 1) / 2, p odd
            if (c.v[i] > h) return true;
            if (c.v[i] < h) return false;
        }
        return false;
    }
    static bool be_to_fp(const uint8_t* in, uint8_t first_mask, Fp* out) {   // big-endian, canonical
        Fp raw = Fp::zero();
        for (int b = 0; b < FPB; b++) {
            uint8_t byte = in[FPB - 1 - b];
            if (b == FPB - 1) byte &= first_mask;
            raw.v[b >> 3] |= (uint64_t)byte << (8 * (b & 7));
        }
        if (Fp::geq_mod(raw.v)) return false;
        *out = Fp::mul(raw, Fp::r2());
        return true;
    }
    struct Flags { uint8_t mask, small, large, inf; int shift; };
    static Flags flags() {   // BN254: 2 flag bits (10 smallest y, 11 largest, 01 infinity); BLS12-381: 3 (100 / 101 / 110)
        return PC::D_TWIST ? Flags{0x3F, 2, 3, 1, 6} : Flags{0x1F, 4, 5, 6, 5};
    }
    // X.A1 || X.A0 with the flags on the first byte; y is "largest" by A1, or by A0 when A1 = 0.  nullptr = ok
    static const char* g2_decompress(const uint8_t* in, G2* out) {
        const Flags f = flags();
        const uint8_t flag = in[0] >> f.shift;
        if (flag == f.inf) { *out = {E2::zero(), E2::zero(), true}; return nullptr; }
        if (flag != f.small && flag != f.large) return "compressed G2: invalid flag";
        E2 x;
        if (!be_to_fp(in, f.mask, &x.b) || !be_to_fp(in + FPB, 0xFF, &x.a)) return "compressed G2: coordinate not reduced";
        E2 y;
        if (!e2_sqrt(x.sqr() * x + twist_b(), &y)) return "compressed G2: x is not on the twist";
        const bool largest = y.b.is_zero() ? lex_largest(y.a) : lex_largest(y.b);
        if (largest != (flag == f.large)) y = y.neg();
        *out = {x, y, false};
        return nullptr;
    }
    static const char* g1_decompress(const uint8_t* in, G1* out) {
        const Flags f = flags();
        const uint8_t flag = in[0] >> f.shift;
        if (flag == f.inf) { *out = {Fp::zero(), Fp::zero(), true}; return nullptr; }
        if (flag != f.small && flag != f.large) return "compressed G1: invalid flag";
        Fp x, y;
        if (!be_to_fp(in, f.mask, &x)) return "compressed G1: coordinate not reduced";
        if (!fp_sqrt(x.sqr() * x + Fp::from_u64(PC::B), &y)) return "compressed G1: x is not on the curve";
        if (lex_largest(y) != (flag == f.large)) y = y.neg();
        *out = {x, y, false};
        return nullptr;
    }

    // line coefficients of the Miller loop of Q, in the order the loop consumes them
    static Prepared prepare(const G2& q) {
        Prepared pr;
        pr.inf = q.inf;
        if (q.inf) return pr;
        const u128 T = PC::loop_count();
        int top = 127;
        while (!((T >> top) & 1)) top--;
        G2 t = q;
        for (int i = top - 1; i >= 0; i--) {
            E2 l;
            G2 t2 = g2_double(t, &l);
            pr.lines.push_back({l, l * t.x - t.y});
            t = t2;
            if ((T >> i) & 1) {
                bool vert;
                G2 t3 = g2_add(t, q, &l, &vert);
                // a vertical chord cannot occur for a point of order r (T < r); keep the loop total anyway
                pr.lines.push_back(vert ? Line{E2::zero(), E2::zero()} : Line{l, l * t.x - t.y});
                t = t3;
            }
        }
        return pr;
    }
    static std::shared_ptr<const Prepared> prepared_cached(const uint8_t* g2_bytes) {
        static std::mutex mu;
        static std::map<std::string, std::shared_ptr<const Prepared>> cache;
        std::string key(reinterpret_cast<const char*>(g2_bytes), 4 * FPB);
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = cache.find(key);
            if (it != cache.end()) return it->second;
        }
        auto pr = std::make_shared<const Prepared>(prepare(load_g2(g2_bytes)));
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 64) cache.clear();
        cache[key] = pr;
        return pr;
    }

    static E12 line_mul(const E12& f, const Line& ln, const G1& p) {
        if (ln.slope.is_zero() && ln.c.is_zero()) return f;
        E2 yp = {p.y, Fp::zero()};
        E2 sx = ln.slope.scale(p.x).neg();
        if (PC::D_TWIST) {
            const int pos[3] = {0, 1, 3};
            const E2 l[3] = {yp, sx, ln.c};
            return f.mul_sparse(pos, l);
        }
        const int pos[3] = {0, 2, 3};
        const E2 l[3] = {ln.c, sx, yp};
        return f.mul_sparse(pos, l);
    }

    // prod_i f_{T,Q_i}(P_i)
    static E12 miller(const std::vector<G1>& ps, const std::vector<std::shared_ptr<const Prepared>>& qs) {
        const u128 T = PC::loop_count();
        int top = 127;
        while (!((T >> top) & 1)) top--;
        std::vector<size_t> active;
        for (size_t k = 0; k < ps.size(); k++)
            if (!ps[k].inf && !qs[k]->inf) active.push_back(k);
        E12 f = E12::one();
        size_t li = 0;
        for (int i = top - 1; i >= 0; i--) {
            f = f.sqr();
            for (size_t k : active) f = line_mul(f, qs[k]->lines[li], ps[k]);
            li++;
            if ((T >> i) & 1) {
                for (size_t k : active) f = line_mul(f, qs[k]->lines[li], ps[k]);
                li++;
            }
        }
        return f;
    }

    // Frobenius constants gamma_i = xi^(i (p-1)/6)
    struct Frob { E2 g[6]; };
    static const Frob& frob_consts() {
        static const Frob fc = [] {
            uint64_t e[Fp::N];
            for (int i = 0; i < Fp::N; i++) e[i] = Fp::M(i);
            e[0] -= 1;   // p is odd: no borrow
            uint64_t rem = 0;
            for (int i = Fp::N - 1; i >= 0; i--) {   // (p-1)/6
                u128 cur = ((u128)rem << 64) | e[i];
                e[i] = (uint64_t)(cur / 6);
                rem = (uint64_t)(cur % 6);
            }
            E2 xi = {Fp::from_u64(PC::XI0), Fp::one()};
            Frob f;
            f.g[0] = E2::one();
            f.g[1] = xi.pow(e, Fp::N);
            for (int i = 2; i < 6; i++) f.g[i] = f.g[i - 1] * f.g[1];
            return f;
        }();
        return fc;
    }
    static E12 frobenius(const E12& x) {
        const Frob& fc = frob_consts();
        E12 r;
        r.c[0] = x.c[0].conj();
        for (int i = 1; i < 6; i++) r.c[i] = x.c[i].conj() * fc.g[i];
        return r;
    }
    // f^x for f in the cyclotomic subgroup (inverse = conjugate)
    static E12 exp_x(const E12& f) {
        E12 r = f.pow_u64(PC::X, true);
        return PC::X_NEG ? r.conj() : r;
    }
    static E12 pow_small(const E12& f, unsigned e) { return f.pow_u64(e, true); }   // cyclotomic inputs only

    static E12 final_exponentiation(const E12& f0) {
        // easy part: f^((p^6 - 1)(p^2 + 1))
        E12 f = f0.conj() * f0.inverse();
        f = frobenius(frobenius(f)) * f;
        if (PC::D_TWIST) {   // BN254
            E12 fx = exp_x(f), fx2 = exp_x(fx), fx3 = exp_x(fx2);
            E12 fx3_36 = pow_small(fx3, 36);
            E12 a = ((f.sqr() * pow_small(fx, 18)) * pow_small(fx2, 30) * fx3_36).conj();          // l0
            E12 b = f * (pow_small(fx, 12) * pow_small(fx2, 18) * fx3_36).conj();                  // l1
            E12 c = f * pow_small(fx2, 6);                                                         // l2
            return a * frobenius(b) * frobenius(frobenius(c)) * frobenius(frobenius(frobenius(f)));
        }
        // BLS12-381: (x-1)^2 (x+p) (x^2+p^2-1) + 3
        E12 y0 = exp_x(f) * f.conj();                  // f^(x-1)
        E12 y1 = exp_x(y0) * y0.conj();                // ^(x-1)
        E12 y2 = exp_x(y1) * frobenius(y1);            // ^(x+p)
        E12 y3 = exp_x(exp_x(y2)) * frobenius(frobenius(y2)) * y2.conj();   // ^(x^2+p^2-1)
        return y3 * f.sqr() * f;
    }

    // prod e(P_i, Q_i) == 1 ; points in gnark memory layout
    static bool product_is_one(const uint8_t* g1s, const uint8_t* g2s, size_t n, std::string* why) {
        std::vector<G1> ps;
        std::vector<std::shared_ptr<const Prepared>> qs;
        for (size_t i = 0; i < n; i++) {
            G1 p = load_g1(g1s + i * 2 * FPB);
            if (!g1_on_curve(p)) { if (why) *why = "G1 point not on the curve"; return false; }
            G2 q = load_g2(g2s + i * 4 * FPB);
            if (!g2_on_curve(q)) { if (why) *why = "G2 point not on the twist"; return false; }
            ps.push_back(p);
            qs.push_back(prepared_cached(g2s + i * 4 * FPB));
        }
        return final_exponentiation(miller(ps, qs)).is_one();
    }
};

}  // namespace hp
}  // namespace b2p
