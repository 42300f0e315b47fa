// plonk.Verify on the host: the self-check (*CompiledCircuit).Verify runs right after plonk.Prove
// (/root/reference/algoplonk.go:93; testutils/testutils.go:51 runs the same call).  gnark's verifier is
// third-party code that is not in /root/reference; the acceptance condition restated here is the one the
// reference's own generated verifiers implement (verifier/templateLogicSigBN254.go:126-397,
// templateLogicSigBLS12_381.go:144-420), on the marshalled proof of helper.go:27-88:
//   1. challenges gamma, beta, alpha, zeta from the SHA-256 transcript (vk commitments, public inputs, L R O,
//      BSB22 commitments, Z, H0 H1 H2)
//   2. PI(zeta) from the public inputs and the BSB22 commitment hashes; alpha^2 L_1(zeta)
//   3. the constant term of the linearised polynomial and its commitment [Lin]
//   4. fold challenge v, folded digest / claimed value of the batch opening at zeta
//   5. batching of the two openings (zeta and omega zeta) with u = H(digest, W_zeta, Z, W_{omega zeta}, zeta, v)
//   6. e(digest, [1]_2) e(-(W_zeta + u W_{omega zeta}), [tau]_2) == 1           (pairing_host.hpp)
// Everything runs in 64-bit-limb host arithmetic; the handful of scalar multiplications share one Straus pass
// per linear combination.  Infinity commitments are hashed the way the prover hashes them (prover.cuh
// point_marshal: gnark's 0x40 flag on BLS12-381, zero bytes on BN254).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "pairing_host.hpp"
#include "sha256.hpp"

namespace b2p {
namespace hp {

template <class PC>
struct HostVerifier {
    using Fp = typename PC::Fp;
    using Fr = typename PC::Fr;
    using PR = Pairing<PC>;
    static constexpr int NB = Fp::N * 8;    // bytes per coordinate
    static constexpr int PB = 2 * NB;       // marshalled point
    static constexpr bool BLS = !PC::D_TWIST;

    struct Aff { Fp x, y; bool inf; };
    struct Ext {                            // extended Jacobian: x = X/ZZ, y = Y/ZZZ
        Fp X, Y, ZZ, ZZZ;
        static Ext inf() { return {Fp::zero(), Fp::zero(), Fp::zero(), Fp::zero()}; }
        bool is_inf() const { return ZZ.is_zero(); }
        static Ext from(const Aff& a) { return a.inf ? inf() : Ext{a.x, a.y, Fp::one(), Fp::one()}; }
        Ext dbl() const {
            if (is_inf() || Y.is_zero()) return inf();
            Fp U = Y.dbl(), V = U.sqr(), W = U * V, S = X * V, X2 = X.sqr(), M = X2.dbl() + X2;
            Fp X3 = M.sqr() - S.dbl();
            return {X3, M * (S - X3) - W * Y, V * ZZ, W * ZZZ};
        }
        Ext add(const Ext& o) const {
            if (is_inf()) return o;
            if (o.is_inf()) return *this;
            Fp U1 = X * o.ZZ, U2 = o.X * ZZ, S1 = Y * o.ZZZ, S2 = o.Y * ZZZ, P = U2 - U1, R = S2 - S1;
            if (P.is_zero()) return R.is_zero() ? dbl() : inf();
            Fp PP = P.sqr(), PPP = P * PP, Q = U1 * PP;
            Fp X3 = R.sqr() - PPP - Q.dbl();
            return {X3, R * (Q - X3) - S1 * PPP, ZZ * o.ZZ * PP, ZZZ * o.ZZZ * PPP};
        }
        Ext add_affine(const Aff& o) const {   // mixed addition: o has ZZ = ZZZ = 1
            if (o.inf) return *this;
            if (is_inf()) return from(o);
            Fp U2 = o.x * ZZ, S2 = o.y * ZZZ, P = U2 - X, R = S2 - Y;
            if (P.is_zero()) return R.is_zero() ? from(o).dbl() : inf();
            Fp PP = P.sqr(), PPP = P * PP, Q = X * PP;
            Fp X3 = R.sqr() - PPP - Q.dbl();
            return {X3, R * (Q - X3) - Y * PPP, ZZ * PP, ZZZ * PPP};
        }
        Aff to_affine() const {
            if (is_inf()) return {Fp::zero(), Fp::zero(), true};
            // ZZ = Z^2, ZZZ = Z^3:  1/ZZ = Z^4 / Z^6 = ZZ^2 / ZZZ^2
            Fp zi = ZZZ.inverse();
            Fp zzi = zi.sqr() * ZZ.sqr();
            return {X * zzi, Y * zi, false};
        }
    };
    static Aff neg(const Aff& a) { return a.inf ? a : Aff{a.x, a.y.neg(), false}; }

    // sum_i s_i P_i, scalars in Montgomery form; Straus with 4-bit windows and shared doublings
    // `plus`: points with coefficient one, added at the end (no table, no window digits)
    static Aff lincomb(const std::vector<Aff>& pts, const std::vector<Fr>& sc_mont, const std::vector<Aff>& plus = {}) {
        const size_t cnt = pts.size();
        std::vector<Ext> tbl(cnt * 16);
        std::vector<Fr> sc(cnt);
        for (size_t i = 0; i < cnt; i++) {
            sc[i] = sc_mont[i].from_mont();
            Ext* t = &tbl[i * 16];
            t[0] = Ext::inf();
            t[1] = Ext::from(pts[i]);
            for (int j = 2; j < 16; j++) t[j] = t[j - 1].add_affine(pts[i]);
        }
        Ext acc = Ext::inf();
        for (int w = Fr::N * 16 - 1; w >= 0; w--) {
            for (int k = 0; k < 4; k++) acc = acc.dbl();
            for (size_t i = 0; i < cnt; i++) {
                const unsigned d = (unsigned)(sc[i].v[w >> 4] >> (4 * (w & 15))) & 15u;
                if (d) acc = acc.add(tbl[i * 16 + d]);
            }
        }
        for (const Aff& p : plus) acc = acc.add_affine(p);
        return acc.to_affine();
    }

    // ---- encodings --------------------------------------------------------------------------------------
    template <class F>
    static bool from_be(const uint8_t* in, int nbytes, F* out, bool reduce) {   // canonical big-endian -> Montgomery
        F raw = F::zero();
        for (int b = 0; b < nbytes; b++) raw.v[b >> 3] |= (uint64_t)in[nbytes - 1 - b] << (8 * (b & 7));
        if (!reduce && F::geq_mod(raw.v)) return false;
        *out = F::mul(raw, F::r2());
        return true;
    }
    template <class F>
    static void to_be(const F& mont, uint8_t* out) {
        F c = mont.from_mont();
        constexpr int nbytes = F::N * 8;
        for (int b = 0; b < nbytes; b++) out[nbytes - 1 - b] = (uint8_t)(c.v[b >> 3] >> (8 * (b & 7)));
    }
    static Fr fr_mod(const uint8_t* be32) { Fr r; from_be(be32, 32, &r, true); return r; }
    // X || Y big-endian.  Infinity: all zero (BN254, MarshalSolidity, helper.go:16-17) or, on BLS12-381, 0x40 followed
    // by zeros -- what G1Affine.RawBytes() writes (helper.go:35; verifier/verifier.go:95-99); all zero is accepted
    // there too.  false: not reduced, not on the curve, or (BLS12-381) not in the r-torsion subgroup.
    static bool parse_point(const uint8_t* in, Aff* out, bool check_subgroup = true) {
        bool zero = true;
        for (int i = 1; i < PB; i++) zero &= in[i] == 0;
        if (zero && (in[0] == 0 || (BLS && in[0] == 0x40))) { *out = {Fp::zero(), Fp::zero(), true}; return true; }
        Aff a{Fp::zero(), Fp::zero(), false};
        if (!from_be(in, NB, &a.x, false) || !from_be(in + NB, NB, &a.y, false)) return false;
        if (!(a.y.sqr() == a.x.sqr() * a.x + Fp::from_u64(PC::B))) return false;
        if (BLS && check_subgroup && !in_g1_subgroup(a)) return false;   // BN254's G1 has cofactor 1: the curve equation suffices
        *out = a;
        return true;
    }
    // r-torsion test for BLS12-381 G1 (cofactor != 1), as gnark's decoders perform it: [r] P == infinity.
    // (Plain double-and-add over the 255-bit group order: ~0.1 ms per point, a dozen points per proof.)
    static bool in_g1_subgroup(const Aff& a) {
        if (a.inf) return true;
        Ext acc = Ext::inf();
        for (int i = Fr::N * 64 - 1; i >= 0; i--) {
            acc = acc.dbl();
            if ((Fr::M(i >> 6) >> (i & 63)) & 1) acc = acc.add_affine(a);
        }
        return acc.is_inf();
    }
    static void marshal(const Aff& a, uint8_t* out, bool transcript_flag) {
        if (a.inf) {
            memset(out, 0, PB);
            if (transcript_flag && BLS) out[0] = 0x40;
            return;
        }
        to_be(a.x, out);
        to_be(a.y, out + NB);
    }
    static Aff load_aff(const uint8_t* mem) {   // gnark G1Affine memory
        Aff a{Fp::load(mem), Fp::load(mem + NB), false};
        a.inf = a.x.is_zero() && a.y.is_zero();
        return a;
    }
    static void store_aff(const Aff& a, uint8_t* mem) {
        if (a.inf) { memset(mem, 0, PB); return; }
        a.x.store(mem); a.y.store(mem + NB);
    }

    // hash_to_field, DST "BSB22-Plonk" (templateLogicSigBN254.go:386-397)
    static Fr hash_fr(const uint8_t* point_bytes, size_t len) {
        static const uint8_t dst_prime[12] = {'B', 'S', 'B', '2', '2', '-', 'P', 'l', 'o', 'n', 'k', 0x0b};
        uint8_t b0[32], b1[32], b2[32], z[64] = {0};
        Sha256 h;
        h.update(z, 64); h.update(point_bytes, len);
        const uint8_t lib[3] = {0x00, 0x30, 0x00};
        h.update(lib, 3); h.update(dst_prime, 12); h.final(b0);
        h.reset(); h.update(b0, 32); uint8_t one = 1; h.update(&one, 1); h.update(dst_prime, 12); h.final(b1);
        uint8_t x[32];
        for (int i = 0; i < 32; i++) x[i] = b0[i] ^ b1[i];
        h.reset(); h.update(x, 32); uint8_t two = 2; h.update(&two, 1); h.update(dst_prime, 12); h.final(b2);
        uint8_t lo[32] = {0};
        memcpy(lo + 16, b2, 16);
        uint64_t e128[1] = {128};
        return fr_mod(b1) * Fr::from_u64(2).pow(e128, 1) + fr_mod(lo);
    }

    static Fr pow_u64(const Fr& b, uint64_t e) { return b.pow(&e, 1); }
    static Fr root_of_unity(uint64_t n) {   // generator of the size-n subgroup (gnark fft.NewDomain)
        Fr w;
        for (int i = 0; i < Fr::N; i++)
            w.v[i] = (uint64_t)PC::FrP::root_[2 * i] | ((uint64_t)PC::FrP::root_[2 * i + 1] << 32);
        int logn = 0;
        while ((1ull << logn) < n) logn++;
        for (int i = logn; i < PC::FrP::TWO_ADICITY; i++) w = w.sqr();
        return w;
    }

    struct Key {
        uint64_t n;
        uint32_t nb_public, k;
        const uint64_t* commit_idx;   // CommitmentConstraintIndexes
        const uint8_t* vk_points;     // S1 S2 S3 Ql Qr Qm Qo Qk Qcp*  (G1Affine memory)
        const uint8_t* g1;            // Kzg.G1
        const uint8_t* g2;            // Kzg.G2[0], Kzg.G2[1]  (G2Affine memory)
    };
    static uint64_t proof_size(uint32_t k) { return 9ull * PB + 6 * 32 + (uint64_t)k * (32 + PB); }

    // One proof on its way through the checks.  The three point combinations of the verifier ([Lin], the folded digest,
    // the pair for the pairing check) are REQUESTS (pts / sc / plus): reduce() answers them with lincomb on the host,
    // the device batch verifier (verify_batch.cuh) answers the requests of a whole batch with one kernel per stage.
    struct Staged {
        Aff LRO[3], H[3], Z, Wz, Wzw;
        Fr l_z, r_z, o_z, s1_z, s2_z, z_zw, zeta, omega, lin_z, claims, v;
        std::vector<Fr> qcp_z;
        std::vector<Aff> bsb, vkp;
        std::vector<uint8_t> vk_fs, lro_fs;
        Aff G1, lin, digest;
        // the current request:  result = sum sc[i] pts[i] + sum plus[j]   (stage 3 has a second one for rhs, negated)
        std::vector<Aff> pts, plus, pts2, plus2;
        std::vector<Fr> sc, sc2;
        std::vector<Aff> to_check;      // points whose r-torsion test was deferred (stage1's defer_subgroup)
    };
    // Steps 1-3: parse, challenges, PI(zeta), the scalars of [Lin].  false = rejected, *why says at which check
    // defer_subgroup: the r-torsion tests of the proof's points (BLS12-381) are left to the caller, which finds the
    // points in st.to_check -- the device batch verifier runs them as [r] P on the GPU with the stage-1 combinations
    static bool stage1(const Key& vk, const uint8_t* proof, uint64_t proof_len, const uint8_t* pub, uint64_t pub_len,
                       Staged& st, std::string* why, bool defer_subgroup = false) {
        auto fail = [&](const char* m) { if (why) *why = m; return false; };
        const uint32_t k = vk.k;
        if (vk.n < 2 || (vk.n & (vk.n - 1))) return fail("domain size is not a power of two");
        if (vk.n > (1ull << PC::FrP::TWO_ADICITY)) return fail("domain size exceeds the scalar field's 2-adicity");
        if (proof_len != proof_size(k)) return fail("proof has the wrong length");
        if (pub_len != 32ull * vk.nb_public) return fail("public inputs have the wrong length");

        // -- proof fields (helper.go:27-88) -------------------------------------------------------------
        const uint8_t* p = proof;
        Aff (&LRO)[3] = st.LRO, (&H)[3] = st.H;
        Aff &Z = st.Z, &Wz = st.Wz, &Wzw = st.Wzw;
        Fr &l_z = st.l_z, &r_z = st.r_z, &o_z = st.o_z, &s1_z = st.s1_z, &s2_z = st.s2_z, &z_zw = st.z_zw;
        std::vector<Fr>& qcp_z = st.qcp_z;
        std::vector<Aff>& bsb = st.bsb;
        qcp_z.assign(k, Fr::zero());
        bsb.assign(k, Aff{Fp::zero(), Fp::zero(), true});
        st.to_check.clear();
        auto point = [&](Aff* a) {
            bool ok = parse_point(p, a, !defer_subgroup);
            p += PB;
            if (ok && BLS && defer_subgroup && !a->inf) st.to_check.push_back(*a);
            return ok;
        };
        auto scalar = [&](Fr* f) { bool ok = from_be(p, 32, f, false); p += 32; return ok; };
        bool ok = true;
        for (int i = 0; i < 3; i++) ok &= point(&LRO[i]);
        for (int i = 0; i < 3; i++) ok &= point(&H[i]);
        if (!ok) return fail("a wire or quotient commitment is not a point of the curve");
        ok &= scalar(&l_z); ok &= scalar(&r_z); ok &= scalar(&o_z); ok &= scalar(&s1_z); ok &= scalar(&s2_z);
        if (!ok) return fail("an evaluation is not reduced mod r");
        if (!point(&Z)) return fail("[Z] is not a point of the curve");
        if (!scalar(&z_zw)) return fail("z(omega zeta) is not reduced mod r");
        if (!point(&Wz) || !point(&Wzw)) return fail("an opening proof is not a point of the curve");
        for (uint32_t c = 0; c < k; c++) if (!scalar(&qcp_z[c])) return fail("qcp(zeta) is not reduced mod r");
        for (uint32_t c = 0; c < k; c++) if (!point(&bsb[c])) return fail("a BSB22 commitment is not a point of the curve");
        std::vector<Fr> pubv(vk.nb_public);
        for (uint32_t i = 0; i < vk.nb_public; i++)
            if (!from_be(pub + 32 * i, 32, &pubv[i], false)) return fail("a public input is not reduced mod r");

        // -- verifying key ----------------------------------------------------------------------------------
        std::vector<Aff>& vkp = st.vkp;
        std::vector<uint8_t>& vk_fs = st.vk_fs;
        vkp.resize(8 + k);
        vk_fs.resize((size_t)(8 + k) * PB);
        for (uint32_t i = 0; i < 8 + k; i++) {
            vkp[i] = load_aff(vk.vk_points + (size_t)i * PB);
            marshal(vkp[i], &vk_fs[(size_t)i * PB], true);
        }
        const Aff &S3 = vkp[2], &Ql = vkp[3], &Qr = vkp[4], &Qm = vkp[5], &Qo = vkp[6], &Qk = vkp[7];
        st.G1 = load_aff(vk.g1);

        // -- challenges (templateLogicSigBN254.go:131-140) ---------------------------------------------------
        uint8_t gamma_pre[32], beta_pre[32], alpha_pre[32], zeta_pre[32], pb[PB];
        std::vector<uint8_t>& lro_fs = st.lro_fs;
        std::vector<uint8_t> bsb_fs((size_t)k * PB);
        lro_fs.resize(3 * PB);
        for (int j = 0; j < 3; j++) marshal(LRO[j], &lro_fs[j * PB], true);
        for (uint32_t c = 0; c < k; c++) marshal(bsb[c], &bsb_fs[(size_t)c * PB], true);
        Sha256 hs;
        hs.update("gamma"); hs.update(vk_fs); hs.update(pub, pub_len); hs.update(lro_fs); hs.final(gamma_pre);
        hs.reset(); hs.update("beta"); hs.update(gamma_pre, 32); hs.final(beta_pre);
        hs.reset(); hs.update("alpha"); hs.update(beta_pre, 32); hs.update(bsb_fs);
        marshal(Z, pb, true); hs.update(pb, PB); hs.final(alpha_pre);
        hs.reset(); hs.update("zeta"); hs.update(alpha_pre, 32);
        for (int j = 0; j < 3; j++) { marshal(H[j], pb, true); hs.update(pb, PB); }
        hs.final(zeta_pre);
        const Fr gamma = fr_mod(gamma_pre), beta = fr_mod(beta_pre), alpha = fr_mod(alpha_pre), zeta = fr_mod(zeta_pre);
        st.zeta = zeta;

        // -- PI(zeta), alpha^2 L_1(zeta) (:142-201) ----------------------------------------------------------
        const Fr one = Fr::one();
        const Fr omega = root_of_unity(vk.n);
        st.omega = omega;
        const Fr zn = pow_u64(zeta, vk.n);
        const Fr zh = zn - one;
        const Fr zh_n = zh * Fr::from_u64(vk.n).inverse();
        Fr PI = Fr::zero();
        {
            // L_i(zeta) = omega^i (zeta^n - 1) / (n (zeta - omega^i)), one inversion for all public inputs
            std::vector<Fr> den(vk.nb_public), pre(vk.nb_public + 1);
            Fr w = one;
            pre[0] = one;
            for (uint32_t i = 0; i < vk.nb_public; i++) {
                den[i] = zeta - w;
                pre[i + 1] = pre[i] * den[i];
                w = w * omega;
            }
            Fr inv = pre[vk.nb_public].inverse();
            std::vector<Fr> li(vk.nb_public);
            for (uint32_t i = vk.nb_public; i > 0; i--) {
                li[i - 1] = inv * pre[i - 1];
                inv = inv * den[i - 1];
            }
            w = one;
            for (uint32_t i = 0; i < vk.nb_public; i++) {
                PI = PI + li[i] * zh_n * w * pubv[i];
                w = w * omega;
            }
            for (uint32_t c = 0; c < k; c++) {
                const Fr wp = pow_u64(omega, vk.nb_public + vk.commit_idx[c]);
                const Fr lag = (zeta - wp).inverse() * wp * zh_n;
                PI = PI + hash_fr(&bsb_fs[(size_t)c * PB], PB) * lag;
            }
        }
        const Fr a2l = (zeta - one).inverse() * zh_n * alpha * alpha;

        // -- constant term of the linearised polynomial (:203-218) --------------------------------------------
        const Fr t1 = l_z + beta * s1_z + gamma, t2 = r_z + beta * s2_z + gamma;
        st.lin_z = (t1 * t2 * (o_z + gamma) * alpha * z_zw + PI - a2l).neg();

        // -- [Lin] (:220-278) ---------------------------------------------------------------------------------
        const Fr u = Fr::from_u64(PC::FrP::SHIFT_SMALL), u2 = u * u;
        const Fr s1p = alpha * beta * z_zw * t1 * t2;
        const Fr bz = beta * zeta;
        const Fr s2p = a2l - alpha * (l_z + bz + gamma) * (r_z + bz * u + gamma) * (o_z + bz * u2 + gamma);
        const Fr zn2 = pow_u64(zeta, vk.n + 2);
        const Fr mzh = zh.neg();
        std::vector<Aff>& pts = st.pts;
        std::vector<Fr>& sc = st.sc;
        pts = {Ql, Qr, Qm, Qo};                                // Qk enters with coefficient one
        sc = {l_z, r_z, l_z * r_z, o_z};
        for (uint32_t c = 0; c < k; c++) { pts.push_back(bsb[c]); sc.push_back(qcp_z[c]); }
        pts.push_back(S3); sc.push_back(s1p);
        pts.push_back(Z); sc.push_back(s2p);
        pts.push_back(H[0]); sc.push_back(mzh);
        pts.push_back(H[1]); sc.push_back(mzh * zn2);
        pts.push_back(H[2]); sc.push_back(mzh * zn2 * zn2);
        st.plus = {Qk};
        return true;
    }
    // Step 4: the fold challenge and the request for the folded digest at zeta (:280-321); consumes [Lin]
    static void stage2(const Key& vk, Staged& st, const Aff& lin) {
        const uint32_t k = vk.k;
        st.lin = lin;
        uint8_t v_pre[32], pb[PB], b32[32];
        Sha256 hs;
        hs.update("gamma");
        to_be(st.zeta, b32); hs.update(b32, 32);
        marshal(lin, pb, false); hs.update(pb, PB);
        hs.update(st.lro_fs);
        hs.update(st.vk_fs.data(), 2 * PB);
        hs.update(st.vk_fs.data() + 8 * PB, (size_t)k * PB);
        to_be(st.lin_z, b32); hs.update(b32, 32);
        const Fr evs[5] = {st.l_z, st.r_z, st.o_z, st.s1_z, st.s2_z};
        for (int i = 0; i < 5; i++) { to_be(evs[i], b32); hs.update(b32, 32); }
        for (uint32_t c = 0; c < k; c++) { to_be(st.qcp_z[c], b32); hs.update(b32, 32); }
        to_be(st.z_zw, b32); hs.update(b32, 32);
        hs.final(v_pre);
        const Fr v = fr_mod(v_pre);
        st.v = v;
        std::vector<Aff> pts = {lin, st.LRO[0], st.LRO[1], st.LRO[2], st.vkp[0], st.vkp[1]};
        std::vector<Fr> vals = {st.lin_z, st.l_z, st.r_z, st.o_z, st.s1_z, st.s2_z};
        for (uint32_t c = 0; c < k; c++) { pts.push_back(st.vkp[8 + c]); vals.push_back(st.qcp_z[c]); }
        std::vector<Fr> sc;
        Fr claims = Fr::zero(), acc = Fr::one();
        for (size_t i = 0; i < pts.size(); i++) {
            sc.push_back(acc);
            claims = claims + vals[i] * acc;
            acc = acc * v;
        }
        st.claims = claims;
        st.pts.assign(pts.begin() + 1, pts.end());
        st.sc.assign(sc.begin() + 1, sc.end());
        st.plus = {lin};                                       // [Lin] enters with coefficient v^0 = 1
    }
    // Step 5: both openings in one pairing check (:323-356); consumes the digest.  Requests: lhs = sum sc pts + plus,
    // rhs = -(sum sc2 pts2 + plus2)
    static void stage3(Staged& st, const Aff& digest) {
        st.digest = digest;
        uint8_t u_pre[32], pb[PB], b32[32];
        Sha256 hs;
        marshal(digest, pb, false); hs.update(pb, PB);
        marshal(st.Wz, pb, false); hs.update(pb, PB);
        marshal(st.Z, pb, true); hs.update(pb, PB);
        marshal(st.Wzw, pb, false); hs.update(pb, PB);
        to_be(st.zeta, b32); hs.update(b32, 32);
        to_be(st.v, b32); hs.update(b32, 32);
        hs.final(u_pre);
        const Fr ub = fr_mod(u_pre);
        const Fr claims = st.claims + st.z_zw * ub;
        st.pts = {st.Z, st.G1, st.Wz, st.Wzw};
        st.sc = {ub, claims.neg(), st.zeta, ub * st.zeta * st.omega};
        st.plus = {digest};
        st.pts2 = {st.Wzw};
        st.sc2 = {ub};
        st.plus2 = {st.Wz};
    }

    // Steps 1-5: everything but the pairing.  On success the proof is valid iff e(lhs, G2[0]) e(rhs, G2[1]) == 1.
    // false = rejected before the pairing, *why says at which check
    static bool reduce(const Key& vk, const uint8_t* proof, uint64_t proof_len, const uint8_t* pub, uint64_t pub_len,
                       Aff* lhs_out, Aff* rhs_out, std::string* why) {
        Staged st;
        if (!stage1(vk, proof, proof_len, pub, pub_len, st, why)) return false;
        stage2(vk, st, lincomb(st.pts, st.sc, st.plus));
        stage3(st, lincomb(st.pts, st.sc, st.plus));
        *lhs_out = lincomb(st.pts, st.sc, st.plus);
        *rhs_out = neg(lincomb(st.pts2, st.sc2, st.plus2));
        return true;
    }

    // The folding weights of a batch: rho_0 = 1, rho_i = 128 bits hashed from the whole batch
    static std::vector<Fr> batch_weights(const uint8_t* proofs, uint64_t proof_len, const uint8_t* pubs, uint64_t pub_len,
                                         uint64_t count) {
        uint8_t seed[32];
        Sha256 hs;
        hs.update("b2p-batch-verify");
        hs.update(proofs, count * proof_len);
        hs.update(pubs, count * pub_len);
        hs.final(seed);
        std::vector<Fr> rho(count);
        if (count) rho[0] = Fr::one();
        for (uint64_t i = 1; i < count; i++) {
            uint8_t d[32], ctr[8], lo[32] = {0};
            for (int b = 0; b < 8; b++) ctr[b] = (uint8_t)(i >> (8 * b));
            hs.reset(); hs.update(seed, 32); hs.update(ctr, 8); hs.final(d);
            memcpy(lo + 16, d, 16);
            rho[i] = fr_mod(lo);
        }
        return rho;
    }

    static bool pair_is_one(const Key& vk, const Aff& lhs, const Aff& rhs, std::string* why) {
        uint8_t g1s[2 * PB];
        store_aff(lhs, g1s);
        store_aff(rhs, g1s + PB);
        std::string pwhy;
        if (PR::product_is_one(g1s, vk.g2, 2, &pwhy)) return true;
        if (why) *why = pwhy.empty() ? "pairing check failed" : pwhy;
        return false;
    }

    // true = accepted; false = rejected, *why says at which check
    static bool verify(const Key& vk, const uint8_t* proof, uint64_t proof_len, const uint8_t* pub, uint64_t pub_len,
                       std::string* why) {
        Aff lhs, rhs;
        return reduce(vk, proof, proof_len, pub, pub_len, &lhs, &rhs, why) && pair_is_one(vk, lhs, rhs, why);
    }

    // `count` proofs of ONE circuit (same key), proof i at proofs + i*proof_len, its public inputs at pubs + i*pub_len.
    // Each proof is reduced to its pair (lhs_i, rhs_i); the pairs are folded with 128-bit coefficients rho_i drawn
    // from a hash of the whole batch (rho_0 = 1), and ONE pairing check decides
    //     e(sum rho_i lhs_i, G2[0]) e(sum rho_i rhs_i, G2[1]) == 1,
    // which holds for an invalid batch with probability 2^-128 (the same folding kzg.BatchVerifyMultiPoints
    // applies to the two openings of one proof).  *bad: index of the first proof rejected before the pairing, or
    // `count` when only the folded check failed (some proof is invalid: verify them one by one to find it).
    static bool verify_batch(const Key& vk, const uint8_t* proofs, uint64_t proof_len, const uint8_t* pubs,
                             uint64_t pub_len, uint64_t count, uint64_t* bad, std::string* why) {
        if (bad) *bad = count;
        if (count == 0) return true;
        std::vector<Aff> lhs(count), rhs(count);
        // the per-proof reductions are independent: spread them over host threads (B2P_VERIFY_THREADS, default
        // min(cores, 8); small batches stay on the calling thread)
        unsigned T = 1;
        if (count >= 4) {
            const char* e = getenv("B2P_VERIFY_THREADS");
            const int asked = e ? atoi(e) : 0;          // garbage / negative / zero: the default
            const unsigned want = asked > 0 ? (unsigned)asked : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
            T = (unsigned)std::min<uint64_t>(std::min(64u, want), count);      // never more than 64 threads
        }
        std::vector<uint64_t> first_bad(T, count);
        std::vector<std::string> whys(T);
        // nothing may escape a worker thread (an exception there is std::terminate, across a C ABI that promises
        // never to abort): a failure inside reduce() -- bad_alloc -- is recorded like a rejected proof
        auto work = [&](unsigned t) {
            uint64_t i = t;
            try {
                for (; i < count; i += T)
                    if (!reduce(vk, proofs + i * proof_len, proof_len, pubs + i * pub_len, pub_len, &lhs[i], &rhs[i], &whys[t])) {
                        first_bad[t] = i;
                        return;
                    }
            } catch (...) {
                first_bad[t] = i < count ? i : count - 1;
                try { whys[t] = "internal error while reducing the proof (out of memory?)"; } catch (...) {}
            }
        };
        if (T == 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            unsigned started = 1;                 // thread 0 is the caller
            try {
                pool.reserve(T - 1);
                for (unsigned t = 1; t < T; t++) { pool.emplace_back(work, t); started = t + 1; }
            } catch (...) {
                // thread creation failed (std::system_error): the shares of the threads that never started are done here
            }
            work(0);
            for (unsigned t = started; t < T; t++) work(t);
            for (auto& th : pool) th.join();
        }
        unsigned best = 0;                        // every thread stops at its own first failure: the smallest wins
        for (unsigned t = 1; t < T; t++)
            if (first_bad[t] < first_bad[best]) best = t;
        if (first_bad[best] < count) {
            if (bad) *bad = first_bad[best];
            if (why) *why = whys[best];
            return false;
        }
        const std::vector<Fr> rho = batch_weights(proofs, proof_len, pubs, pub_len, count);
        return pair_is_one(vk, lincomb(lhs, rho), lincomb(rhs, rho), why);
    }
};

}  // namespace hp
}  // namespace b2p
