// G1 arithmetic for short-Weierstrass curves with a = 0 (BN254, BLS12-381).
// Bucket accumulators use extended Jacobian ("XYZZ") coordinates:
//   x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0.
// Affine points use gnark's in-memory G1Affine layout: X || Y Montgomery limbs,
// (0,0) == infinity.
#pragma once
#include "field.cuh"

namespace b2p {

template <class Fp>
struct Affine {
    Fp x, y;
    HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    HD static Affine inf() { return Affine{Fp::zero(), Fp::zero()}; }
};

template <class Fp>
struct XYZZ {
    Fp X, Y, ZZ, ZZZ;

    HD static XYZZ inf() { return XYZZ{Fp::zero(), Fp::zero(), Fp::zero(), Fp::zero()}; }
    HD bool is_inf() const { return ZZ.is_zero(); }

    HD static XYZZ from_affine(const Affine<Fp>& p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, Fp::one(), Fp::one()};
    }

    // 2 * (affine point)   -- mdbl-2008-s-1
    HD static XYZZ dbl_affine(const Fp& x, const Fp& y) {
        XYZZ r;
        Fp U = y.dbl();
        Fp V = U.sqr();
        Fp W = U * V;
        Fp S = x * V;
        Fp xx = x.sqr();
        Fp M = xx.dbl() + xx;
        r.X = M.sqr() - S.dbl();
        r.Y = Fp::mul_sub(M, S - r.X, W, y);
        r.ZZ = V;
        r.ZZZ = W;
        return r;
    }

    // this += (x, y) affine, (x,y) not infinity   -- madd-2008-s
    HD void add_affine(const Fp& x, const Fp& y) {
        if (is_inf()) {
            X = x; Y = y; ZZ = Fp::one(); ZZZ = Fp::one();
            return;
        }
        Fp Pv = x * ZZ - X;
        Fp Rv = y * ZZZ - Y;
        if (Pv.is_zero()) {
            if (Rv.is_zero()) *this = dbl_affine(x, y);
            else *this = inf();
            return;
        }
        Fp PP = Pv.sqr();
        Fp PPP = Pv * PP;
        Fp Q = X * PP;
        Fp X3 = Rv.sqr() - PPP - Q.dbl();
        Y = Fp::mul_sub(Rv, Q - X3, Y, PPP);       // two products, one Montgomery reduction
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }
    // signed variant: neg != 0 adds -(x,y)
    HD void add_affine_signed(const Affine<Fp>& p, bool neg) {
        if (p.is_inf()) return;
        Fp y = neg ? p.y.neg() : p.y;
        add_affine(p.x, y);
    }

    // dbl-2008-s-1
    HD XYZZ dbl() const {
        if (is_inf()) return *this;
        XYZZ r;
        Fp U = Y.dbl();
        Fp V = U.sqr();
        Fp W = U * V;
        Fp S = X * V;
        Fp xx = X.sqr();
        Fp M = xx.dbl() + xx;
        r.X = M.sqr() - S.dbl();
        r.Y = Fp::mul_sub(M, S - r.X, W, Y);
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }

    // this += o   -- add-2008-s
    HD void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        Fp U1 = X * o.ZZ;
        Fp U2 = o.X * ZZ;
        Fp S1 = Y * o.ZZZ;
        Fp S2 = o.Y * ZZZ;
        Fp Pv = U2 - U1;
        Fp Rv = S2 - S1;
        if (Pv.is_zero()) {
            if (Rv.is_zero()) *this = dbl();
            else *this = inf();
            return;
        }
        Fp PP = Pv.sqr();
        Fp PPP = Pv * PP;
        Fp Q = U1 * PP;
        Fp X3 = Rv.sqr() - PPP - Q.dbl();
        Y = Fp::mul_sub(Rv, Q - X3, S1, PPP);
        X = X3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }

    HD XYZZ neg() const { return XYZZ{X, Y.neg(), ZZ, ZZZ}; }

    // k * this for a small non-negative k (double-and-add, MSB first)
    HDN XYZZ mul_small(uint64_t k) const {
        XYZZ acc = inf();
        for (int i = 63; i >= 0; i--) {
            acc = acc.dbl();
            if ((k >> i) & 1) acc.add(*this);
        }
        return acc;
    }

    // to affine: one field inversion.  x = X * ZZ^2 * I^2, y = Y * I with I = 1/ZZZ
    HDN Affine<Fp> to_affine() const {
        if (is_inf()) return Affine<Fp>::inf();
        Fp I = ZZZ.inverse();
        Fp I2 = I.sqr();
        Affine<Fp> a;
        a.x = X * ZZ.sqr() * I2;
        a.y = Y * I;
        return a;
    }
};

}  // namespace b2p
