// Prime-field arithmetic in Montgomery form, 32-bit limbs, register resident.
// One template serves the four fields of the hot path; the in-memory layout is
// gnark-crypto's fr.Element / fp.Element (little-endian limbs, R = 2^(32N)), so
// buffers handed over the C-ABI by the Go shim are used as they are
// ("pass unsafe.Pointer(&pk.Kzg.G1[0])").
#pragma once
#include <cstdint>
#include "ptx.cuh"
#include "field_params.cuh"

namespace b2p {

template <class P>
struct Field {
    static constexpr int N = P::N;
    using Params = P;
    uint32_t v[N];

    // ---- constants ------------------------------------------------------
    HD static Field zero() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    HD static Field one() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::one(i);
        return r;
    }
    HD static Field r2() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
        return r;
    }
    HD static Field modulus() {
        Field r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::mod(i);
        return r;
    }

    HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i];
        return acc == 0;
    }
    HD bool operator==(const Field& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= v[i] ^ o.v[i];
        return acc == 0;
    }
    HD bool operator!=(const Field& o) const { return !(*this == o); }

    // ---- add / sub ------------------------------------------------------
    // r = a - p if a >= p else a   (a < 2p)
    HD static void final_sub(uint32_t* a) {
        uint32_t t[N];
        t[0] = ptx::sub_cc(a[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(a[i], P::mod(i));
        uint32_t borrow = ptx::subc(0, 0);   // 0 or 0xffffffff
#pragma unroll
        for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
    }

    HD friend Field operator+(const Field& a, const Field& b) {
        Field r;
        r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
        r.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);   // moduli leave >= 1 spare bit: no carry out
        final_sub(r.v);
        return r;
    }

    HD friend Field operator-(const Field& a, const Field& b) {
        Field r;
        r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = ptx::subc(0, 0);
        // add p back when the subtraction borrowed (masked add keeps it branch-free)
        uint32_t t0 = ptx::add_cc(r.v[0], P::mod(0) & borrow);
        r.v[0] = t0;
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], P::mod(i) & borrow);
        r.v[N - 1] = ptx::addc(r.v[N - 1], P::mod(N - 1) & borrow);
        return r;
    }

    HD Field neg() const { return is_zero() ? *this : (modulus_raw_sub(*this)); }

    HD static Field modulus_raw_sub(const Field& a) {   // p - a, a != 0
        Field r;
        r.v[0] = ptx::sub_cc(P::mod(0), a.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::subc_cc(P::mod(i), a.v[i]);
        r.v[N - 1] = ptx::subc(P::mod(N - 1), a.v[N - 1]);
        return r;
    }

    HD Field dbl() const { return *this + *this; }

    // K * x for a small compile-time K as a chain of modular additions: the coset shifts (5, 7) and their
    // squares multiply on the ALU pipe instead of occupying the IMAD pipe with a full Montgomery product
    template <uint32_t K>
    HD Field mul_small() const {
        static_assert(K >= 1, "mul_small needs K >= 1");
        int top = 31;
        while (!((K >> top) & 1u)) top--;
        Field acc = *this;
#pragma unroll
        for (int b = top - 1; b >= 0; b--) {
            acc = acc.dbl();
            if ((K >> b) & 1u) acc = acc + *this;
        }
        return acc;
    }

    // ---- Montgomery multiplication --------------------------------------
    // Coarsely-integrated operand scanning with the products of even and odd
    // limbs of `a` kept in two accumulators, so that every (lo,hi) pair lands on
    // adjacent registers of one carry chain: ptxas fuses each
    // mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32(.X).
    //   acc[j] of `even` is column j, acc[j] of `odd` is column j+1.
    // Provenance: the even/odd two-accumulator scheme and the helper names below (mul_n, cmad_n, madc_n_rshift,
    // mad_n_redc) follow the widely used `mont_t` GPU multiplier of Supranational's sppark (Apache-2.0), as
    // published with the ZPrize MSM entries; re-typed here for this template (any N, host emulation of the carry
    // flag, the separated product/reduction variants further down are this repo's).  Nothing of it comes from the
    // reference repository, which holds no field arithmetic of its own (gnark-crypto is un-vendored).

    // acc[0..N) = a[0], a[2], ... times bi (disjoint 64-bit products)
    HD static void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            acc[j] = ptx::mul_lo(a[j], bi);
            acc[j + 1] = ptx::mul_hi(a[j], bi);
        }
    }
    // acc += a[0], a[2], ... times bi; leaves the carry out in CC
    HD static void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
        acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
        acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
            acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
        }
    }
    // same, modulus limbs as the multiplicand (compile-time immediates)
    template <int OFF>
    HD static void cmad_mod(uint32_t* acc, uint32_t mi) {
        acc[0] = ptx::mad_lo_cc(P::mod(OFF), mi, acc[0]);
        acc[1] = ptx::madc_hi_cc(P::mod(OFF), mi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = ptx::madc_lo_cc(P::mod(OFF + j), mi, acc[j]);
            acc[j + 1] = ptx::madc_hi_cc(P::mod(OFF + j), mi, acc[j + 1]);
        }
    }
    // odd[j] = a[j]*bi + odd[j+2] with the incoming carry (CC), i.e. add and shift right by two limbs
    HD static void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            odd[j] = ptx::madc_lo_cc(a[j], bi, odd[j + 2]);
            odd[j + 1] = ptx::madc_hi_cc(a[j], bi, odd[j + 3]);
        }
        odd[N - 2] = ptx::madc_lo_cc(a[N - 2], bi, 0);
        odd[N - 1] = ptx::madc_hi(a[N - 2], bi, 0);
    }
    template <bool FIRST>
    HD static void mad_n_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
        if (FIRST) {
            mul_n(odd, a + 1, bi);
            mul_n(even, a, bi);
        } else {
            even[0] = ptx::add_cc(even[0], odd[1]);
            madc_n_rshift(odd, a + 1, bi);
            cmad_n(even, a, bi);
            odd[N - 1] = ptx::addc(odd[N - 1], 0);
        }
        uint32_t mi = even[0] * P::INV;
        cmad_mod<1>(odd, mi);
        cmad_mod<0>(even, mi);
        odd[N - 1] = ptx::addc(odd[N - 1], 0);
    }

#if !defined(__CUDA_ARCH__) && !defined(B2P_HOST_EMULATE_DEVICE_MUL)
    // Host build: the same Montgomery product on 64-bit limbs with a 128-bit accumulator
    // (the carry-flag emulation of ptx.cuh is exact but ~20x slower; the prover's host side
    // inverts a few scalars and converts the proof points to affine with this).
    // -DB2P_HOST_EMULATE_DEVICE_MUL keeps the emulated device code path (tests/test_host.py).
    static uint64_t inv64() {
        // P::INV = -p^-1 mod 2^32; one Newton step lifts p^-1 to 64 bits
        const uint64_t p0 = (uint64_t)P::mod(0) | ((uint64_t)P::mod(1) << 32);
        uint64_t y = (uint64_t)(uint32_t)(0u - P::INV);
        y *= 2 - p0 * y;
        return 0 - y;
    }
    static Field mul_host(const Field& a, const Field& b) {
        constexpr int M = N / 2;
        typedef unsigned __int128 u128;
        uint64_t A[M], B[M], Pm[M], t[M + 2];
        for (int i = 0; i < M; i++) {
            A[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
            B[i] = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
            Pm[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
        }
        for (int i = 0; i < M + 2; i++) t[i] = 0;
        static const uint64_t ninv = inv64();
        for (int i = 0; i < M; i++) {
            u128 c = 0;
            for (int j = 0; j < M; j++) {
                c += (u128)A[j] * B[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M] = (uint64_t)c;
            t[M + 1] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * ninv;
            c = (u128)m * Pm[0] + t[0];
            c >>= 64;
            for (int j = 1; j < M; j++) {
                c += (u128)m * Pm[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M - 1] = (uint64_t)c;
            t[M] = t[M + 1] + (uint64_t)(c >> 64);
        }
        // t < 2p (operands < p, or one operand any N-limb value as in reduce_to_mont): one conditional subtraction
        uint64_t d[M];
        unsigned char borrow = 0;
        for (int i = 0; i < M; i++) {
            u128 s = (u128)t[i] - Pm[i] - borrow;
            d[i] = (uint64_t)s;
            borrow = (unsigned char)((s >> 64) & 1);
        }
        const bool ge = t[M] != 0 || !borrow;
        Field r;
        for (int i = 0; i < M; i++) {
            const uint64_t x = ge ? d[i] : t[i];
            r.v[2 * i] = (uint32_t)x;
            r.v[2 * i + 1] = (uint32_t)(x >> 32);
        }
        return r;
    }
#endif

    HD friend Field operator*(const Field& a, const Field& b) {
#if !defined(__CUDA_ARCH__) && !defined(B2P_HOST_EMULATE_DEVICE_MUL)
        return mul_host(a, b);
#else
        uint32_t even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            if (i == 0) mad_n_redc<true>(even, odd, a.v, b.v[0]);
            else        mad_n_redc<false>(even, odd, a.v, b.v[i]);
            mad_n_redc<false>(odd, even, a.v, b.v[i + 1]);
        }
        // merge: result column j = even[j] + odd[j+1]
        Field r;
        r.v[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(even[i], odd[i + 1]);
        r.v[N - 1] = ptx::addc(even[N - 1], 0);
        final_sub(r.v);
        return r;
#endif
    }

    // ---- separated product / reduction: squaring, a*b - c*d ------------------
    // A 2N-limb value is kept as two accumulators of aligned (lo, hi) pairs, like the product above:
    //   value = sum_k E[k] 2^(32k) + sum_k O[k] 2^(32(k+1)).
    // acc[col .. col+2*cnt) += x[0], x[2], ... (cnt limbs, stride 2) times y, carry into acc[col + 2*cnt]
    template <int CNT>
    HD static void wide_mad_chain(uint32_t* acc, const uint32_t* x, uint32_t y) {
        if (CNT <= 0) return;
        acc[0] = ptx::mad_lo_cc(x[0], y, acc[0]);
        acc[1] = ptx::madc_hi_cc(x[0], y, acc[1]);
#pragma unroll
        for (int t = 1; t < CNT; t++) {
            acc[2 * t] = ptx::madc_lo_cc(x[2 * t], y, acc[2 * t]);
            acc[2 * t + 1] = ptx::madc_hi_cc(x[2 * t], y, acc[2 * t + 1]);
        }
        acc[2 * CNT] = ptx::addc(acc[2 * CNT], 0);
    }
    // flat[0..2N) = E + (O << 32)
    HD static void wide_merge(uint32_t* flat, const uint32_t* E, const uint32_t* O) {
        flat[0] = E[0];
        flat[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; k++) flat[k] = ptx::addc_cc(E[k], O[k - 1]);
        flat[2 * N - 1] = ptx::addc(E[2 * N - 1], O[2 * N - 2]);
    }
    // (E, O) = a * b + c * d, E and O zero on entry.  The rows of the two products are interleaved: the limb
    // that receives the carry out of a chain (acc[2 * CNT] in wide_mad_chain) must hold nothing but earlier
    // carries, or the addition could overflow it -- adding the second product after the first one is complete
    // would hit full limbs there (a 2^-32 event per row: about one wrong point addition per 30 proofs).
    HD static void wide_mul2(uint32_t* E, uint32_t* O, const uint32_t* a, const uint32_t* b, const uint32_t* c,
                             const uint32_t* d) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            // x[j] * y[i] lands on column i + j: even columns -> E, odd columns -> O (column k is O[k-1])
            if (i % 2 == 0) {
                wide_mad_chain<N / 2>(E + i, a, b[i]);            // j = 0, 2, ...
                wide_mad_chain<N / 2>(E + i, c, d[i]);
                wide_mad_chain<N / 2>(O + i, a + 1, b[i]);        // j = 1, 3, ...: column i + j - 1 in O
                wide_mad_chain<N / 2>(O + i, c + 1, d[i]);
            } else {
                wide_mad_chain<N / 2>(O + i - 1, a, b[i]);        // j even, i odd: odd column i + j
                wide_mad_chain<N / 2>(O + i - 1, c, d[i]);
                wide_mad_chain<N / 2>(E + i + 1, a + 1, b[i]);    // j odd: even column i + j
                wide_mad_chain<N / 2>(E + i + 1, c + 1, d[i]);
            }
        }
    }
    // flat = a^2: the off-diagonal products once, doubled, plus the diagonal
    HD static void wide_sqr(uint32_t* flat, const uint32_t* a) {
        uint32_t E[2 * N + 2], O[2 * N + 2];
#pragma unroll
        for (int k = 0; k < 2 * N + 2; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            // j = i+1, i+3, ...: odd column 2i+1, ... -> O[2i ...];  j = i+2, i+4, ...: even column 2i+2 ... -> E
            wide_mad_chain_rt(O + 2 * i, a + i + 1, a[i], (N - i) / 2);
            wide_mad_chain_rt(E + 2 * i + 2, a + i + 2, a[i], (N - i - 1) / 2);
        }
        uint32_t S[2 * N];
        wide_merge(S, E, O);
        // flat = 2 S + sum_i a_i^2 2^(64 i): operands first, then one uninterrupted carry chain
        uint32_t D[2 * N], S2[2 * N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            D[2 * i] = ptx::mul_lo(a[i], a[i]);
            D[2 * i + 1] = ptx::mul_hi(a[i], a[i]);
        }
        S2[0] = S[0] << 1;
#pragma unroll
        for (int k = 1; k < 2 * N; k++) S2[k] = (S[k] << 1) | (S[k - 1] >> 31);
        flat[0] = ptx::add_cc(D[0], S2[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) flat[k] = ptx::addc_cc(D[k], S2[k]);
        flat[2 * N - 1] = ptx::addc(D[2 * N - 1], S2[2 * N - 1]);
    }
    // same chain with a trip count that is a compile-time constant after unrolling
    HD static void wide_mad_chain_rt(uint32_t* acc, const uint32_t* x, uint32_t y, int cnt) {
        if (cnt <= 0) return;
        acc[0] = ptx::mad_lo_cc(x[0], y, acc[0]);
        acc[1] = ptx::madc_hi_cc(x[0], y, acc[1]);
#pragma unroll
        for (int t = 1; t < N / 2; t++) {
            if (t < cnt) {
                acc[2 * t] = ptx::madc_lo_cc(x[2 * t], y, acc[2 * t]);
                acc[2 * t + 1] = ptx::madc_hi_cc(x[2 * t], y, acc[2 * t + 1]);
            }
        }
        acc[2 * cnt] = ptx::addc(acc[2 * cnt], 0);
    }
    // Montgomery reduction of a 2N-limb value T < 2 p R:  T / R mod p  =  mul_by_1(T_lo) + T_hi, reduced.
    // mul_by_1 is the product loop above with the a * b_i rows left out (N rows of m_i * p).
    template <bool FIRST>
    HD static void redc_row(uint32_t* even, uint32_t* odd) {
        if (!FIRST) {
            even[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j++) odd[j] = ptx::addc_cc(odd[j + 2], 0);
            odd[N - 2] = ptx::addc(0, 0);
            odd[N - 1] = 0;
        }
        uint32_t mi = even[0] * P::INV;
        cmad_mod<1>(odd, mi);
        cmad_mod<0>(even, mi);
        odd[N - 1] = ptx::addc(odd[N - 1], 0);
    }
    template <int SUBS>
    HD static Field wide_redc(const uint32_t* flat) {
        uint32_t even[N], odd[N];
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            even[j] = flat[j]; even[j + 1] = 0;
            odd[j] = flat[j + 1]; odd[j + 1] = 0;
        }
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            if (i == 0) redc_row<true>(even, odd);
            else        redc_row<false>(even, odd);
            redc_row<false>(odd, even);
        }
        // (T_lo + M p) / R  <=  p, then + T_hi
        Field r;
        r.v[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(even[i], odd[i + 1]);
        r.v[N - 1] = ptx::addc(even[N - 1], 0);
        r.v[0] = ptx::add_cc(r.v[0], flat[N]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], flat[N + i]);
        r.v[N - 1] = ptx::addc(r.v[N - 1], flat[2 * N - 1]);
#pragma unroll
        for (int k = 0; k < SUBS; k++) final_sub(r.v);
        return r;
    }

    HD Field sqr() const {
#if !defined(__CUDA_ARCH__) && !defined(B2P_HOST_EMULATE_DEVICE_MUL)
        return mul_host(*this, *this);
#else
        uint32_t T[2 * N];
        wide_sqr(T, v);
        return wide_redc<1>(T);
#endif
    }
    // a * b - c * d with ONE Montgomery reduction
    HD static Field mul_sub(const Field& a, const Field& b, const Field& c, const Field& d) {
        static_assert(P::BITS + 2 <= 32 * N, "mul_sub needs 3p < 2^(32N)");
#if !defined(__CUDA_ARCH__) && !defined(B2P_HOST_EMULATE_DEVICE_MUL)
        return mul_host(a, b) - mul_host(c, d);
#else
        // a*b - c*d == a*b + (p - c)*d (mod p): both products are accumulated into ONE pair of accumulators
        uint32_t nc[N];
        nc[0] = ptx::sub_cc(P::mod(0), c.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) nc[i] = ptx::subc_cc(P::mod(i), c.v[i]);
        nc[N - 1] = ptx::subc(P::mod(N - 1), c.v[N - 1]);
        uint32_t E[2 * N + 2], O[2 * N + 2];
#pragma unroll
        for (int k = 0; k < 2 * N + 2; k++) { E[k] = 0; O[k] = 0; }
        wide_mul2(E, O, a.v, b.v, nc, d.v);
        uint32_t T[2 * N];
        wide_merge(T, E, O);
        return wide_redc<2>(T);      // T < 2 p^2 < 2 p R: the result is below 3 p before the subtractions
#endif
    }

    HD Field& operator+=(const Field& o) { *this = *this + o; return *this; }
    HD Field& operator-=(const Field& o) { *this = *this - o; return *this; }
    HD Field& operator*=(const Field& o) { *this = *this * o; return *this; }

    // ---- conversions ----------------------------------------------------
    HD Field to_mont() const { return (*this) * r2(); }   // *this must be canonical (< p)
    // Any N-limb value (e.g. a 256-bit hash, possibly >= p) -> Montgomery form of (x mod p).
    // The unreduced value must be the operand that is scanned limb by limb (the right one):
    // the running sum then stays below 2p; as the left operand it can overflow N limbs.
    HD static Field reduce_to_mont(const Field& raw) { return r2() * raw; }
    HD Field from_mont() const {
        Field o = zero();
        o.v[0] = 1;
        return (*this) * o;
    }
    HD static Field from_u32(uint32_t x) {
        Field o = zero();
        o.v[0] = x;
        return o.to_mont();
    }

    // ---- exponentiation / inversion -------------------------------------
    HDN Field pow_u64(uint64_t e) const {
        Field acc = one(), base = *this;
        while (e) {
            if (e & 1) acc = acc * base;
            base = base.sqr();
            e >>= 1;
        }
        return acc;
    }
    // a^(p-2); zero maps to zero (gnark's Inverse convention)
    HDN Field inverse() const {
        Field acc = one();
        for (int i = P::BITS - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((P::pm2(i >> 5) >> (i & 31)) & 1) acc = acc * (*this);
        }
        return acc;
    }
};

using FrBn254 = Field<Bn254FrParams>;
using FpBn254 = Field<Bn254FpParams>;
using FrBls12381 = Field<Bls12381FrParams>;
using FpBls12381 = Field<Bls12381FpParams>;

}  // namespace b2p
