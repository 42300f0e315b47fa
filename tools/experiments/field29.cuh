// Reduced-radix prime-field arithmetic for the multiplier-bound kernels: L limbs of W bits in 32-bit registers
// (BN254 Fp/Fr, BLS12-381 Fr: 9 x 29; BLS12-381 Fp: 14 x 28), Montgomery constant R' = 2^(L W).
//
// Why a second representation next to field.cuh's 8/12 x 32-bit limbs (which stays the MEMORY format -- it is
// gnark-crypto's): with W-bit limbs a partial product is below 2^(2W) and a 64-bit column accumulator takes 2L of them
// without overflowing, so every partial product is one  IMAD.WIDE.U32 Rd(64) = Ra * Rb + Rc(64)  with no carry flag in or
// out.  Measured on B200 (tools/microbench.cu, profiles/microbench_r2_*.json): that instruction issues every 2.4
// cycles per SM sub-partition, the carry-chained IMAD.WIDE.U32.X the 32-bit-limb multiplier consists of every 4.0
// (and ptxas' alternative, IMAD + IMAD.HI, costs 2 + 2).  81 + 81 carry-free products per 254-bit multiplication
// against 64 + 64 chained ones: ~400 multiplier-pipe cycles instead of ~512.  Carries are resolved once per row with
// shifts and adds on the otherwise idle ALU pipe.
//
// Everything here is plain C++ on uint32_t / uint64_t: the same code runs on the host (tests/test_host.py drives it
// against Python integers).
//
// Invariants ("normalized"): limbs 0..L-2 are below 2^W, the top limb holds the rest; the VALUE may exceed p (lazy
// reduction).  mul / sqr accept operands whose limbs are below 2^(W+1) and whose values satisfy a * b < R' * p, and return
// a normalized value below 2p.  sub<K>(a, b) = a + K p - b needs b normalized and b <= K p.
#pragma once
#include <cstdint>
#include "../../algoplonk_b200/csrc/field.cuh"
#include "field29_params.cuh"

namespace b2p {

template <class P29, class F32>
struct F29 {
    static constexpr int L = P29::L;
    static constexpr int W = P29::W;
    static constexpr int N32 = P29::N32;
    static constexpr uint32_t MASK = (1u << W) - 1;
    using Params = P29;
    using Mem = F32;                 // the memory-format field this one converts from / to
    uint32_t v[L];

    HD static F29 zero() {
        F29 r;
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = 0;
        return r;
    }
    HD static F29 one() {            // 1 in R'-Montgomery form
        F29 r;
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = P29::one(i);
        return r;
    }
    HD bool limbs_all_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < L; i++) acc |= v[i];
        return acc == 0;
    }

    // ---- radix change (the integer is unchanged) ---------------------------------------------------------
    // w: N32 32-bit words, little endian
    HD static F29 unpack(const uint32_t* w) {
        F29 r;
#pragma unroll
        for (int i = 0; i < L; i++) {
            const int bit = i * W, lo = bit >> 5, sh = bit & 31;
            uint64_t x = lo < N32 ? w[lo] : 0u;
            if (lo + 1 < N32) x |= (uint64_t)w[lo + 1] << 32;
            const uint32_t t = (uint32_t)(x >> sh);
            r.v[i] = i == L - 1 ? t : (t & MASK);
        }
        return r;
    }
    // normalized value below 2^(32 N32) -> words
    HD void pack(uint32_t* w) const {
#pragma unroll
        for (int k = 0; k < N32; k++) {
            // bits [32k, 32k+32) of sum_i v[i] 2^(W i)
            const int i0 = (32 * k) / W, off = 32 * k - i0 * W;      // limb holding the first bit, offset inside it
            uint64_t x = (uint64_t)v[i0] >> off;
            if (i0 + 1 < L) x |= (uint64_t)v[i0 + 1] << (W - off);
            if (i0 + 2 < L && 2 * W - off < 32) x |= (uint64_t)v[i0 + 2] << (2 * W - off);
            w[k] = (uint32_t)x;
        }
    }

    // ---- carry resolution ------------------------------------------------------------------------------------
    HD void normalize() {
#pragma unroll
        for (int i = 0; i < L - 1; i++) {
            v[i + 1] += v[i] >> W;
            v[i] &= MASK;
        }
    }
    HD friend F29 add(const F29& a, const F29& b) {
        F29 r;
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = a.v[i] + b.v[i];
        r.normalize();
        return r;
    }
    HD F29 dbl() const { return add(*this, *this); }
    // a + K p - b with b normalized and b <= K p: every limb of K p below the top one was lifted by 2^W (params kpK)
    template <int K>
    HD static uint32_t kp(int i) {
        static_assert(K == 2 || K == 4 || K == 8 || K == 16, "K p constants exist for K = 2, 4, 8, 16");
        return K == 2 ? P29::kp2(i) : K == 4 ? P29::kp4(i) : K == 8 ? P29::kp8(i) : P29::kp16(i);
    }
    template <int K>
    HD static F29 sub(const F29& a, const F29& b) {
        F29 r;
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = a.v[i] + kp<K>(i) - b.v[i];
        r.normalize();
        return r;
    }
    template <int K>
    HD F29 neg() const {             // K p - x
        F29 r;
#pragma unroll
        for (int i = 0; i < L; i++) r.v[i] = kp<K>(i) - v[i];
        r.normalize();
        return r;
    }

    // ---- Montgomery multiplication, R' = 2^(L W) ---------------------------------------------------------------
    // Operand scanning over a rolling window of L 64-bit columns: row i adds a * b_i and m_i * p (m_i clears the low
    // limb), then the window moves down one limb.  Column bound: L products below 2^(2W+2) (both operands lazy, limbs
    // < 2^(W+1)) plus L below 2^(2W) plus carries: 9 (2^60 + 2^58) < 2^64 for 9 x 29, 14 (2^58 + 2^56) for 14 x 28.
    HD static void mont_rows(uint64_t* t, const uint32_t* a, const uint32_t* b) {
#pragma unroll
        for (int i = 0; i < L; i++) {
#pragma unroll
            for (int j = 0; j < L; j++) t[j] += (uint64_t)a[j] * b[i];
            redc_step(t);
        }
    }
    HD static void redc_step(uint64_t* t) {
        const uint32_t m = ((uint32_t)t[0] * P29::INV) & MASK;
#pragma unroll
        for (int j = 0; j < L; j++) t[j] += (uint64_t)m * P29::mod(j);
        const uint64_t carry = t[0] >> W;
#pragma unroll
        for (int j = 0; j < L - 1; j++) t[j] = t[j + 1];
        t[0] += carry;
        t[L - 1] = 0;
    }
    HD static F29 columns_to_limbs(const uint64_t* t) {
        F29 r;
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < L - 1; j++) {
            const uint64_t x = t[j] + c;
            r.v[j] = (uint32_t)x & MASK;
            c = x >> W;
        }
        r.v[L - 1] = (uint32_t)(t[L - 1] + c);
        return r;
    }
    HD friend F29 operator*(const F29& a, const F29& b) {
        uint64_t t[L];
#pragma unroll
        for (int j = 0; j < L; j++) t[j] = 0;
        mont_rows(t, a.v, b.v);
        return columns_to_limbs(t);
    }
    // a^2: the off-diagonal products once with a doubled operand (limbs < 2^(W+1) stay within the column bound
    // when *this is normalized; for a lazy *this the caller uses operator*)
    HD F29 sqr() const {
        uint32_t d[L];
#pragma unroll
        for (int j = 0; j < L; j++) d[j] = v[j] << 1;
        uint64_t t[2 * L];
#pragma unroll
        for (int j = 0; j < 2 * L; j++) t[j] = 0;
#pragma unroll
        for (int i = 0; i < L; i++) {
            t[2 * i] += (uint64_t)v[i] * v[i];
#pragma unroll
            for (int j = i + 1; j < L; j++) t[i + j] += (uint64_t)v[i] * d[j];
        }
        // reduction: L rows of m p over the 2L columns
#pragma unroll
        for (int i = 0; i < L; i++) {
            const uint32_t m = ((uint32_t)t[i] * P29::INV) & MASK;
#pragma unroll
            for (int j = 0; j < L; j++) t[i + j] += (uint64_t)m * P29::mod(j);
            t[i + 1] += t[i] >> W;
        }
        return columns_to_limbs(t + L);
    }
    // a b + c d with one reduction (the two products share their columns); value bound a b + c d < R' p
    HD static F29 mul_add(const F29& a, const F29& b, const F29& c, const F29& d) {
        uint64_t t[L];
#pragma unroll
        for (int j = 0; j < L; j++) t[j] = 0;
#pragma unroll
        for (int i = 0; i < L; i++) {
#pragma unroll
            for (int j = 0; j < L; j++) t[j] += (uint64_t)a.v[j] * b.v[i];
#pragma unroll
            for (int j = 0; j < L; j++) t[j] += (uint64_t)c.v[j] * d.v[i];
            redc_step(t);
        }
        return columns_to_limbs(t);
    }

    // ---- exact reductions / tests (rare paths and the final store) --------------------------------------------
    template <int K>
    HD static uint32_t mulk(int i) {             // K p, plain limbs
        return K == 1 ? P29::mul1(i) : K == 2 ? P29::mul2(i) : K == 4 ? P29::mul4(i) : K == 8 ? P29::mul8(i) : P29::mul16(i);
    }
    // if (*this >= K p) *this -= K p   (normalized limbs)
    template <int K>
    HD void cond_sub() {
        bool ge = true;
        for (int i = L - 1; i >= 0; i--) {
            const uint32_t q = mulk<K>(i);
            if (v[i] != q) { ge = v[i] > q; break; }
        }
        if (!ge) return;
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < L; i++) {
            const uint32_t x = v[i] - mulk<K>(i) - borrow;       // mod 2^32
            if (i < L - 1) {
                borrow = x >> 31;                                // the true difference is in (-2^W, 2^W)
                v[i] = x & MASK;
            } else {
                v[i] = x;
            }
        }
    }
    // canonical representative in [0, p); *this normalized and below 32 p
    HD F29 reduce_full() const {
        F29 r = *this;
        r.template cond_sub<16>();
        r.template cond_sub<8>();
        r.template cond_sub<4>();
        r.template cond_sub<2>();
        r.template cond_sub<1>();
        return r;
    }
    // value == 0 mod p ?  Exact.  A multiple k p has low limb k p_0 mod 2^W, so (v_0 / p_0 mod 2^W) = k is tiny:
    // one multiplication rules the common case out; the full reduction runs only then.
    HD bool is_zero_mod_p() const {
        const uint32_t k = (v[0] * P29::PINV_POS) & MASK;
        if (k >= 32) return false;
        return reduce_full().limbs_all_zero();
    }

    // ---- memory format <-> registers ------------------------------------------------------------------------
    // x R (memory, canonical)  ->  x R' (registers)
    HD static F29 from_mem(const F32& x) {
        F29 c;
#pragma unroll
        for (int i = 0; i < L; i++) c.v[i] = P29::to29(i);
        return unpack(x.v) * c;
    }
    // x R' stored as a plain integer in the memory layout (what the MSM table holds): radix change only
    HD static F29 from_words(const F32& x) { return unpack(x.v); }
    // x R' (registers, any lazy value the multiplier accepts)  ->  x R (memory, canonical)
    HD F32 to_mem() const {
        F29 c;
#pragma unroll
        for (int i = 0; i < L; i++) c.v[i] = P29::from29(i);
        const F29 r = (*this * c).reduce_full();
        F32 o;
        r.pack(o.v);
        return o;
    }
    // x R' canonical, as a plain integer in the memory layout
    HD F32 to_words() const {
        const F29 r = reduce_full();
        F32 o;
        r.pack(o.v);
        return o;
    }
};

using Fp29Bn254 = F29<Bn254FpParams29, FpBn254>;
using Fr29Bn254 = F29<Bn254FrParams29, FrBn254>;
using Fr29Bls12381 = F29<Bls12381FrParams29, FrBls12381>;
using Fp29Bls12381 = F29<Bls12381FpParams29, FpBls12381>;

// ---------------------------------------------------------------------------------------------------------------
// XYZZ accumulator over F29 (madd-2008-s), lazy reduction.  Value bounds (p = modulus; every product below 2p):
//   X < 8p, Y < 2p, ZZ, ZZZ < 2p;  Pv = x ZZ - X < 10p,  Rv = y ZZZ - Y < 4p,  D = Q - X3 < 10p;
//   largest product Pv^2 < 100 p^2 < R' p  (R' > 69 p for all four fields).
// The affine operand (x, y) is canonical (below p).
// ---------------------------------------------------------------------------------------------------------------
template <class F>
struct XYZZ29 {
    F X, Y, ZZ, ZZZ;
    HD static XYZZ29 inf() { return XYZZ29{F::zero(), F::zero(), F::zero(), F::zero()}; }
    HD bool is_inf() const { return ZZ.limbs_all_zero(); }       // ZZ is set to exact zero for infinity, never lazily

    HD static XYZZ29 dbl_affine(const F& x, const F& y) {        // mdbl-2008-s-1, operands canonical
        XYZZ29 r;
        const F U = y.dbl();                     // < 2p
        const F V = U.sqr();
        const F Wv = U * V;
        const F S = x * V;
        const F xx = x.sqr();
        const F M = add(xx.dbl(), xx);           // < 6p
        const F X3 = F::template sub<4>(M.sqr(), S.dbl());       // < 6p
        r.X = X3;
        r.Y = F::mul_add(M, F::template sub<8>(S, X3), Wv, y.template neg<2>());     // M (S - X3) - W y
        r.ZZ = V;
        r.ZZZ = Wv;
        return r;
    }
    // this += (x, y), the point not at infinity
    HD void add_affine(const F& x, const F& y) {
        if (is_inf()) {
            X = x; Y = y; ZZ = F::one(); ZZZ = F::one();
            return;
        }
        const F Pv = F::template sub<8>(x * ZZ, X);              // < 10p
        const F Rv = F::template sub<2>(y * ZZZ, Y);             // < 4p
        if (Pv.is_zero_mod_p()) {
            if (Rv.is_zero_mod_p()) *this = dbl_affine(x, y);
            else *this = inf();
            return;
        }
        const F PP = Pv.sqr();                   // limbs are normalized; the value bound is what matters: 100 p^2 < R' p
        const F PPP = Pv * PP;
        const F Q = X * PP;
        const F X3 = F::template sub<4>(F::template sub<2>(Rv.sqr(), PPP), Q.dbl());   // < 2p + 2p + 4p
        Y = F::mul_add(Rv, F::template sub<8>(Q, X3), Y.template neg<2>(), PPP);      // Rv (Q - X3) - Y PPP
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }
};

}  // namespace b2p
