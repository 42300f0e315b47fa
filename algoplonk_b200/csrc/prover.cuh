// PLONK prover orchestration: the replacement for gnark v0.15.0
// backend/plonk/{bn254,bls12-381}.Prove called from the reference at
// algoplonk.go:89.  The five rounds follow SURVEY Appendix A; the schedule is
// GPU-first (whole 4n coset resident, selectors pre-evaluated at circuit load,
// every polynomial lives in HBM from the moment L,R,O arrive until the 9 proof
// points leave).  The host only hashes (SHA-256) and does O(1) scalar algebra.
#pragma once
#include <chrono>
#include <memory>
#include <vector>
#include "common.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "poly.cuh"
#include "sha256.hpp"
#include "../../include/b200plonk.h"
#include "iface.hpp"

namespace b2p {

// ---------------------------------------------------------------------------
// host-side byte conversions (gnark Marshal()/Bytes(): big-endian canonical)
// ---------------------------------------------------------------------------
template <class F>
inline void field_to_be(const F& mont, uint8_t* out) {
    F c = mont.from_mont();
    constexpr int NB = F::N * 4;
    for (int i = 0; i < F::N; i++)
        for (int b = 0; b < 4; b++) out[NB - 1 - (4 * i + b)] = (uint8_t)(c.v[i] >> (8 * b));
}
// 32 big-endian bytes -> Fr (reduced mod r, Montgomery form)
template <class Fr>
inline Fr fr_from_be32_mod(const uint8_t* in) {
    Fr raw;
    for (int i = 0; i < 8; i++)
        raw.v[i] = ((uint32_t)in[31 - 4 * i]) | ((uint32_t)in[30 - 4 * i] << 8) | ((uint32_t)in[29 - 4 * i] << 16) |
                   ((uint32_t)in[28 - 4 * i] << 24);
    return Fr::reduce_to_mont(raw);
}
template <class C>
inline void point_marshal(const Affine<typename C::Fp>& p, uint8_t* out, bool gnark_inf_flag) {
    constexpr int NB = C::Fp::N * 4;
    if (p.is_inf()) {
        memset(out, 0, 2 * NB);
        if (gnark_inf_flag && C::ID == B2P_BLS12_381) out[0] = 0x40;   // mUncompressedInfinity
        return;
    }
    field_to_be(p.x, out);
    field_to_be(p.y, out + NB);
}

// inverse of point_marshal: uncompressed X || Y big-endian (gnark G1Affine.Marshal()); the flag bits gnark
// keeps in the top of the first byte are masked, an infinity flag or all-zero coordinates give infinity
template <class C>
inline Affine<typename C::Fp> point_unmarshal(const uint8_t* in) {
    using Fp = typename C::Fp;
    constexpr int NB = Fp::N * 4;
    const uint8_t mask = C::ID == B2P_BLS12_381 ? 0xE0 : 0xC0;
    if ((in[0] & mask) == 0x40) return Affine<Fp>::inf();
    Fp c[2];
    for (int k = 0; k < 2; k++) {
        for (int i = 0; i < Fp::N; i++) c[k].v[i] = 0;
        for (int b = 0; b < NB; b++) {
            uint8_t byte = in[k * NB + NB - 1 - b];
            if (k == 0 && b == NB - 1) byte &= (uint8_t)~mask;
            c[k].v[b >> 2] |= (uint32_t)byte << (8 * (b & 3));
        }
    }
    if (c[0].is_zero() && c[1].is_zero()) return Affine<Fp>::inf();
    return Affine<Fp>{c[0].to_mont(), c[1].to_mont()};
}

// sum_i s_i * P_i for a handful of points, on the host (Straus, 4-bit windows, shared doublings).
// Used for the commitment to the linearised polynomial, which is a combination of commitments
// that are already known (the verifier computes the same combination,
// templateLogicSigBN254.go:256-278); runs while the device is busy with an MSM.
template <class C>
inline Affine<typename C::Fp> host_msm_small(const Affine<typename C::Fp>* pts, const typename C::Fr* scalars_mont,
                                             int cnt) {
    using Fp = typename C::Fp;
    using Fr = typename C::Fr;
    std::vector<XYZZ<Fp>> tbl((size_t)cnt * 16);
    std::vector<Fr> sc(cnt);
    for (int i = 0; i < cnt; i++) {
        sc[i] = scalars_mont[i].from_mont();
        XYZZ<Fp>* t = &tbl[(size_t)i * 16];
        t[0] = XYZZ<Fp>::inf();
        t[1] = XYZZ<Fp>::from_affine(pts[i]);
        for (int j = 2; j < 16; j++) { t[j] = t[j - 1]; t[j].add(t[1]); }
    }
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    for (int w = Fr::N * 8 - 1; w >= 0; w--) {
        for (int k = 0; k < 4; k++) acc = acc.dbl();
        for (int i = 0; i < cnt; i++) {
            const uint32_t d = (sc[i].v[w >> 3] >> (4 * (w & 7))) & 15u;
            if (d) acc.add(tbl[(size_t)i * 16 + d]);
        }
    }
    return acc.to_affine();
}

// hash_to_field with DST "BSB22-Plonk" (templateLogicSigBN254.go:386-397)
template <class C>
inline typename C::Fr hash_fr(const uint8_t* point_bytes, size_t len) {
    using Fr = typename C::Fr;
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
    // (b1 * 2^128 + b2[:16]) mod r
    uint8_t lo[32] = {0};
    memcpy(lo + 16, b2, 16);
    Fr hi = fr_from_be32_mod<Fr>(b1);
    Fr two128 = Fr::from_u32(2).pow_u64(128);
    return hi * two128 + fr_from_be32_mod<Fr>(lo);
}

// [tau^(first+j)] G1 for j < n  (unsafekzg.NewSRS; first > 0: one rank's shard of it)
template <class C>
__global__ void k_srs_from_tau(Affine<typename C::Fp>* __restrict__ out, uint64_t first, uint64_t stride, uint64_t n,
                               typename C::Fr tau, Affine<typename C::Fp> g) {
    using Fr = typename C::Fr;
    using Fp = typename C::Fp;
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    Fr s = tau.pow_u64(first + j * stride).from_mont();
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    for (int b = Fr::Params::BITS - 1; b >= 0; b--) {
        acc = acc.dbl();
        if ((s.v[b >> 5] >> (b & 31)) & 1) acc.add_affine(g.x, g.y);
    }
    Affine<Fp> a = acc.to_affine();
    st_field(&out[j].x, a.x);
    st_field(&out[j].y, a.y);
}

// Compressed G1 stream -> affine Montgomery points: what srs.Pk.ReadFrom does to the embedded pk.bin files
// (setup/setup.go:173,189; format pinned by setup/trusted_setup_test.go:53-59,132,290-303).  One thread per
// point: x from the big-endian bytes under the flag bits, y = (x^3 + b)^((p+1)/4) (both base fields are
// 3 mod 4), the flag picks the lexicographically smallest / largest root.  err: 0 ok, 1 bad flag,
// 2 x >= p, 3 x^3 + b is not a square (x not on the curve).
template <class C>
__global__ void k_g1_decompress(const uint8_t* __restrict__ in, Affine<typename C::Fp>* __restrict__ out, uint64_t n,
                                typename C::Fp exp_p1_4 /* (p+1)/4, plain limbs */, uint32_t curve_b,
                                uint32_t* __restrict__ err) {
    using Fp = typename C::Fp;
    constexpr int NB = Fp::N * 4;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + i * NB;
    const bool bls = C::ID == B2P_BLS12_381;
    const uint32_t flag = bls ? (b[0] >> 5) : (b[0] >> 6);
    const uint8_t mask = bls ? 0x1F : 0x3F;
    const uint32_t f_small = bls ? 4u : 2u, f_large = bls ? 5u : 3u, f_inf = bls ? 6u : 1u;
    if (flag == f_inf) {
        st_field(&out[i].x, Fp::zero());
        st_field(&out[i].y, Fp::zero());
        return;
    }
    if (flag != f_small && flag != f_large) { atomicMax(err, 1u); return; }
    Fp x;
#pragma unroll
    for (int l = 0; l < Fp::N; l++) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int idx = NB - 1 - (4 * l + k);           // byte of weight 8 * (4l + k)
            uint32_t byte = b[idx];
            if (idx == 0) byte &= mask;
            w |= byte << (8 * k);
        }
        x.v[l] = w;
    }
    // canonical? (x < p)
    bool lt = false, decided = false;
    for (int l = Fp::N - 1; l >= 0 && !decided; l--) {
        const uint32_t pm = Fp::Params::mod(l);
        if (x.v[l] != pm) { lt = x.v[l] < pm; decided = true; }
    }
    if (!lt) { atomicMax(err, 2u); return; }
    x = x.to_mont();
    const Fp rhs = x.sqr() * x + Fp::from_u32(curve_b);
    Fp y = Fp::one();
    for (int bit = Fp::Params::BITS - 1; bit >= 0; bit--) {
        y = y.sqr();
        if ((exp_p1_4.v[bit >> 5] >> (bit & 31)) & 1) y = y * rhs;
    }
    if (y.sqr() != rhs) { atomicMax(err, 3u); return; }
    // lexicographically largest root?  y > p - y as integers
    const Fp yc = y.from_mont();
    const Fp ny = y.neg().from_mont();
    bool larger = false;
    decided = false;
    for (int l = Fp::N - 1; l >= 0 && !decided; l--)
        if (yc.v[l] != ny.v[l]) { larger = yc.v[l] > ny.v[l]; decided = true; }
    if (larger != (flag == f_large)) y = y.neg();
    st_field(&out[i].x, x);
    st_field(&out[i].y, y);
}

// ---------------------------------------------------------------------------------------------------------
// kzg.ToLagrangeG1 (setup/setup.go:124,138): the Lagrange-basis SRS  [L_j(tau)]_1 = (1/n) sum_i w^(-ij) [tau^i]_1,
// i.e. the inverse DFT of the canonical points over the group -- gnark runs it as an FFT whose butterflies multiply a
// point by a twiddle (minutes on a CPU at 2^20).  Same dataflow here: decimation in time over XYZZ points in HBM, one
// launch per stage, one butterfly per thread:  B' = [w^-e] B (double-and-add over the twiddle's bits),  A + B', A - B';
// the 1/n is folded into the last stage (A and B' are multiplied by 1/n and w^-e / n there).  The prover does not need
// this table (Lagrange commitments are iNTT + canonical MSM, msm.cuh); it exists for callers that want gnark's own
// pk.KzgLagrange -- the Go shim's fallback to plonk.Prove -- without paying for it on the CPU.
// ---------------------------------------------------------------------------------------------------------
template <class C>
__device__ XYZZ<typename C::Fp> ec_scalar_mul(const XYZZ<typename C::Fp>& P, const typename C::Fr& k_canonical) {
    using Fp = typename C::Fp;
    using Fr = typename C::Fr;
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    int top = Fr::Params::BITS - 1;
    while (top >= 0 && !((k_canonical.v[top >> 5] >> (top & 31)) & 1)) top--;
    for (int b = top; b >= 0; b--) {
        acc = acc.dbl();
        if ((k_canonical.v[b >> 5] >> (b & 31)) & 1) acc.add(P);
    }
    return acc;
}
// buf[i] = table[brev(i)] lifted to XYZZ (the bit reversal of the DIT input is the gather)
template <class C>
__global__ void k_ecntt_load(XYZZ<typename C::Fp>* __restrict__ buf, const Affine<typename C::Fp>* __restrict__ pts, int logn) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << logn)) return;
    const uint32_t j = logn ? brev32(i, logn) : 0u;
    Affine<typename C::Fp> p;
    p.x = ld_field(&pts[j].x);
    p.y = ld_field(&pts[j].y);
    st_xyzz(buf + i, XYZZ<typename C::Fp>::from_affine(p));
}
// stage s (half-block 2^s) of the inverse transform; tw_inv[k] = w^-k, k < n/2 (Montgomery form)
template <class C>
__global__ void __launch_bounds__(128) k_ecntt_stage(XYZZ<typename C::Fp>* __restrict__ buf, const typename C::Fr* __restrict__ tw_inv,
                                                     int logn, int s, typename C::Fr n_inv, int last) {
    using Fr = typename C::Fr;
    using Ext = XYZZ<typename C::Fp>;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (1u << (logn - 1))) return;
    const uint32_t half = 1u << s, k = t & (half - 1);
    const uint32_t i = ((t >> s) << (s + 1)) | k, j = i + half;
    Ext A = ld_xyzz(buf + i), B = ld_xyzz(buf + j);
    Fr w = ld_field(tw_inv + ((uint64_t)k << (logn - 1 - s)));
    if (last) {
        A = ec_scalar_mul<C>(A, n_inv.from_mont());
        w = w * n_inv;
    }
    if (k != 0 || last) B = ec_scalar_mul<C>(B, w.from_mont());       // w = 1 for k = 0: nothing to multiply
    Ext S = A, D = A;
    S.add(B);
    D.add(B.neg());
    st_xyzz(buf + i, S);
    st_xyzz(buf + j, D);
}
template <class C>
__global__ void __launch_bounds__(128) k_ecntt_store(Affine<typename C::Fp>* __restrict__ out, const XYZZ<typename C::Fp>* __restrict__ buf,
                                                     uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine<typename C::Fp> a = ld_xyzz(buf + i).to_affine();
    st_field(&out[i].x, a.x);
    st_field(&out[i].y, a.y);
}

template <class C> struct CurveConsts;
template <> struct CurveConsts<Bn254> {
    static constexpr uint32_t B = 3;      // y^2 = x^3 + 3
    static Affine<FpBn254> generator() {
        Affine<FpBn254> g;
        g.x = FpBn254::from_u32(1);
        g.y = FpBn254::from_u32(2);
        return g;
    }
};
template <> struct CurveConsts<Bls12381> {
    static constexpr uint32_t B = 4;      // y^2 = x^3 + 4
    static Affine<FpBls12381> generator() {
        // canonical little-endian 32-bit limbs of the standard generator
        static const uint32_t gx[12] = {0xdb22c6bbu, 0xfb3af00au, 0xf97a1aefu, 0x6c55e83fu, 0x171bac58u, 0xa14e3a3fu,
                                        0x9774b905u, 0xc3688c4fu, 0x4fa9ac0fu, 0x2695638cu, 0x3197d794u, 0x17f1d3a7u};
        static const uint32_t gy[12] = {0x46c5e7e1u, 0x0caa2329u, 0xa2888ae4u, 0xd03cc744u, 0x2c04b3edu, 0x00db18cbu,
                                        0xd5d00af6u, 0xfcf5e095u, 0x741d8ae4u, 0xa09e30edu, 0xe3aaa0f1u, 0x08b3f481u};
        Affine<FpBls12381> g;
        for (int i = 0; i < 12; i++) { g.x.v[i] = gx[i]; g.y.v[i] = gy[i]; }
        g.x = g.x.to_mont();
        g.y = g.y.to_mont();
        return g;
    }
};

// Where the commitments of a proving key go when they are not computed on its own table: a shard group
// (shard_group.cuh) spreads every kzg.Commit over the GPUs of the box.
struct CommitRouter {
    virtual ~CommitRouter() {}
    virtual void begin_proof() = 0;
    virtual void commit(const void* d_scalars, uint64_t n, int slot, cudaStream_t st) = 0;
    virtual void fetch(int first, int cnt, void* host_affine_out, cudaStream_t st) = 0;
    // the proof's size-4n transforms spread over the same ranks (shard_group.cuh); `which`: 0..3 = el er eo ez
    virtual bool shards_ntt() const = 0;
    virtual void ntt_forward(const void* d_coeffs, uint64_t len, int which, cudaStream_t st) = 0;
    virtual void ntt_forward_wait(cudaStream_t st) = 0;
    virtual void ntt_inverse(uint64_t out_len, cudaStream_t st) = 0;
};

// ---------------------------------------------------------------------------
// SRS handle
// ---------------------------------------------------------------------------
template <class C>
struct Srs : SrsBase {
    using Fr = typename C::Fr;
    using Fp = typename C::Fp;
    using Aff = Affine<Fp>;
    using Ext = XYZZ<Fp>;
    cudaStream_t stream = nullptr;
    MsmEngine<C> msm;
    DevBuf<Fr> scratch;         // scalar staging for b2p_msm_g1
    Profiler* prof = nullptr;
    static constexpr int OUT_SLOTS = MSM_SLOTS;
    // b2p_srs_set_commit_hook: every commitment on this handle is delegated (a point-set-sharded MSM over
    // several GPUs, algoplonk_b200/sharded_prover.py); results wait in hook_out until fetch()
    b2p_commit_fn hook = nullptr;
    void* hook_ctx = nullptr;
    Aff hook_out[MSM_SLOTS];
    void set_commit_hook(b2p_commit_fn fn, void* ctx) override { hook = fn; hook_ctx = ctx; }
    // b2p_shard_group_attach: the native multi-GPU route (peer memory + flags, no host hop per commitment)
    CommitRouter* router = nullptr;

    Srs() { curve = C::ID; B2P_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)); }
    ~Srs() override { if (stream) cudaStreamDestroy(stream); }

    void finish_init() {
        B2P_CUDA(cudaStreamSynchronize(stream));
    }
    void load(const void* pts, uint64_t n) override {
        B2P_REQUIRE(n >= 1, "empty SRS");
        msm.load(pts, n, env_force_c(), stream);
        finish_init();
    }
    // pk.bin payload (count compressed points, header already stripped by the caller)
    void load_compressed(const uint8_t* bytes, uint64_t n) override {
        B2P_REQUIRE(n >= 1, "empty SRS");
        constexpr int NB = Fp::N * 4;
        MsmPlan pl = msm_plan(n, Fr::Params::BITS, env_force_c());
        B2P_REQUIRE((uint64_t)pl.W * n < (1ull << 31), "SRS too large for 31-bit table indices");
        DevBuf<Aff> tbl((size_t)pl.W * n);
        DevBuf<uint8_t> raw(n * NB);
        DevBuf<uint32_t> err(1);
        B2P_CUDA(cudaMemcpyAsync(raw.p, bytes, n * NB, cudaMemcpyHostToDevice, stream));
        B2P_CUDA(cudaMemsetAsync(err.p, 0, sizeof(uint32_t), stream));
        // (p + 1) / 4 as plain limbs
        Fp e;
        uint64_t carry = 1;
        for (int l = 0; l < Fp::N; l++) {
            const uint64_t v = (uint64_t)Fp::Params::mod(l) + carry;
            e.v[l] = (uint32_t)v;
            carry = v >> 32;
        }
        for (int l = 0; l < Fp::N; l++)
            e.v[l] = (e.v[l] >> 2) | (l + 1 < Fp::N ? e.v[l + 1] << 30 : 0u);
        B2P_LAUNCH((k_g1_decompress<C>), div_up(n, 128), 128, 0, stream, raw.p, tbl.p, n, e, CurveConsts<C>::B, err.p);
        uint32_t herr = 0;
        B2P_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof herr, cudaMemcpyDeviceToHost, stream));
        B2P_CUDA(cudaStreamSynchronize(stream));
        B2P_REQUIRE(herr != 1, "compressed SRS: invalid point flag");
        B2P_REQUIRE(herr != 2, "compressed SRS: coordinate not reduced");
        B2P_REQUIRE(herr != 3, "compressed SRS: point not on the curve");
        msm.load_device(std::move(tbl), n, pl, stream);
        finish_init();
    }
    void get_points(uint64_t first, uint64_t count, void* out) const override {
        B2P_REQUIRE(count <= msm.npoints && first <= msm.npoints - count, "range exceeds SRS size");
        B2P_CUDA(cudaMemcpy(out, msm.table.p + first, count * sizeof(Aff), cudaMemcpyDeviceToHost));
    }
    // kzg.ToLagrangeG1(srs.Pk.G1[:n]) (setup/setup.go:124,138) -> n G1Affine on the host
    void to_lagrange(uint64_t n, void* out) override {
        B2P_REQUIRE(n >= 1 && (n & (n - 1)) == 0 && n <= msm.npoints, "ToLagrangeG1: n must be a power of two <= SRS size");
        int logn = 0;
        while ((1ull << logn) < n) logn++;
        B2P_REQUIRE(logn <= Fr::Params::TWO_ADICITY && logn <= 31, "ToLagrangeG1: n exceeds the field's 2-adicity");
        NttDomain<Fr> d;
        d.init(logn, false, stream);
        DevBuf<Ext> buf(n);
        DevBuf<Aff> res(n);
        B2P_LAUNCH((k_ecntt_load<C>), div_up(n, 128), 128, 0, stream, buf.p, msm.table.p, logn);
        for (int s = 0; s < logn; s++)
            B2P_LAUNCH((k_ecntt_stage<C>), div_up(n / 2, 128), 128, 0, stream, buf.p, d.tw_inv.p, logn, s, d.n_inv,
                       (int)(s == logn - 1));
        B2P_LAUNCH((k_ecntt_store<C>), div_up(n, 128), 128, 0, stream, res.p, buf.p, n);
        B2P_CUDA(cudaMemcpyAsync(out, res.p, n * sizeof(Aff), cudaMemcpyDeviceToHost, stream));
        B2P_CUDA(cudaStreamSynchronize(stream));
    }
    uint64_t size() const override { return msm.npoints; }
    void* stream_handle() override { return (void*)stream; }
    void msm_params(int* c, int* windows, uint64_t* buckets) const override {
        if (c) *c = msm.plan.c;
        if (windows) *windows = msm.plan.W;
        if (buckets) *buckets = msm.plan.nbuckets;
    }
    // b2p_msm_g1: host scalars (Montgomery) -> affine result on the host
    void msm_g1(int basis, const void* scalars, uint64_t n, void* out, bool device_scalars) override {
        B2P_REQUIRE(!router, "this SRS handle's commitments are routed to a shard group: only b2p_prove may use it");
        B2P_REQUIRE(n <= msm.npoints, "more scalars than SRS points");
        if (scratch.n < n + 1) scratch.alloc(n + 1);
        if (n) B2P_CUDA(cudaMemcpyAsync(scratch.p, scalars, n * sizeof(Fr),
                                        device_scalars ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
        if (basis == B2P_BASIS_LAGRANGE) {
            // MSM(Lagrange SRS, v) == commit(iNTT(v)) on the canonical SRS
            B2P_REQUIRE(n >= 1 && (n & (n - 1)) == 0, "Lagrange basis needs a power-of-two length");
            int logn = 0;
            while ((1ull << logn) < n) logn++;
            NttDomain<Fr> d;
            d.init(logn, false, stream);
            d.inverse_natural(scratch.p, stream);
            B2P_CUDA(cudaStreamSynchronize(stream));   // domain tables die with `d`
        } else {
            B2P_REQUIRE(basis == B2P_BASIS_CANONICAL, "unknown basis");
        }
        commit_async(scratch.p, n, 0);
        Aff a;
        fetch(0, 1, &a);
        memcpy(out, &a, sizeof a);
    }
    // [tau^(first + j*stride)]_1, j < n: the whole SRS (first 0, stride 1), a block of it, or a cyclic shard
    void generate_unsafe(const void* tau_p, uint64_t first, uint64_t stride, uint64_t n) override {
        Fr tau;
        memcpy(&tau, tau_p, sizeof tau);
        B2P_REQUIRE(n >= 1, "empty SRS");
        MsmPlan pl = msm_plan(n, Fr::Params::BITS, env_force_c());
        B2P_REQUIRE((uint64_t)pl.W * n < (1ull << 31), "SRS too large for 31-bit table indices");
        DevBuf<Aff> tbl((size_t)pl.W * n);
        B2P_REQUIRE(stride >= 1 && stride <= (1ull << 20) && first < (1ull << 40) && n < (1ull << 40), "SRS range too large");
        B2P_LAUNCH((k_srs_from_tau<C>), div_up(n, 128), 128, 0, stream, tbl.p, first, stride, n, tau,
                   CurveConsts<C>::generator());
        msm.load_device(std::move(tbl), n, pl, stream);
        finish_init();
    }
    static int env_force_c() {
        const char* e = getenv("B2P_MSM_C");
        return e ? atoi(e) : 0;
    }
    // queue an MSM of n device scalars (Montgomery form); result lands in slot
    void commit_async(const Fr* d_scalars, uint64_t n, int slot) {
        if (router) {
            int id = prof ? prof->begin(B2P_STAT_MSM_MS, stream) : -1;
            router->commit(d_scalars, n, slot, stream);
            if (prof) prof->end(id, stream);
            return;
        }
        if (hook) {
            B2P_REQUIRE(slot >= 0 && slot < OUT_SLOTS, "result slot out of range");
            B2P_CUDA(cudaStreamSynchronize(stream));      // the scalars are final before the hook reads them
            if (hook(hook_ctx, d_scalars, n, &hook_out[slot]) != 0)
                throw Error(B2P_ERR_INTERNAL, "the commit hook reported a failure");
            return;
        }
        int id = prof ? prof->begin(B2P_STAT_MSM_MS, stream) : -1;
        msm.prof = prof;
        msm.run_async(d_scalars, n, true, stream, slot);
        if (prof) prof->end(id, stream);
    }
    // bring slots [first, first+cnt) to the host (synchronises) and convert them to affine there:
    // one field inversion per point takes ~10 us on a host core, against ~160 us for a
    // single-thread kernel on the device.
    void fetch(int first, int cnt, Aff* host_out) {
        Ext h[OUT_SLOTS];
        B2P_REQUIRE(first >= 0 && cnt >= 0 && first + cnt <= OUT_SLOTS, "result slot out of range");
        if (router) {
            router->fetch(first, cnt, host_out, stream);
            return;
        }
        if (hook) {
            for (int i = 0; i < cnt; i++) host_out[i] = hook_out[first + i];
            return;
        }
        int id = prof ? prof->begin(B2P_STAT_MSM_MS, stream) : -1;
        msm.finish_async(first, cnt, stream);          // the latency-bound end of the reductions, once for all slots
        if (prof) prof->end(id, stream);
        B2P_CUDA(cudaMemcpyAsync(h, msm.result.p + first, cnt * sizeof(Ext), cudaMemcpyDeviceToHost, stream));
        B2P_CUDA(cudaStreamSynchronize(stream));
        for (int i = 0; i < cnt; i++) host_out[i] = h[i].to_affine();
    }
};

// ---------------------------------------------------------------------------
// circuit handle + prover
// ---------------------------------------------------------------------------
template <class C>
struct Circuit : CircuitBase {
    using Fr = typename C::Fr;
    using Fp = typename C::Fp;
    using Aff = Affine<Fp>;
    static constexpr int PB = 2 * Fp::N * 4;   // marshalled point bytes

    Srs<C>* srs;
    cudaStream_t st;
    Fr* h_pub = nullptr;                       // pinned: public inputs of a device-resident witness
    cudaStream_t copy_st = nullptr;            // host -> device uploads of the wire columns
    cudaEvent_t ev_ready = nullptr, ev_col[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t n, m;
    int logn, logm, log_rho;
    uint32_t nb_public, k;
    std::vector<uint64_t> commit_idx;
    NttDomain<Fr> d0, d1;
    Profiler prof;

    // resident circuit data
    DevBuf<Fr> lag_qk, lag_S;
    DevBuf<Fr> c_ql, c_qr, c_qm, c_qo, c_qk, c_s1, c_s2, c_s3;
    DevBuf<Fr> e_ql, e_qr, e_qm, e_qo, e_qk, e_s1, e_s2, e_s3, e_x, e_l1;
    std::vector<DevBuf<Fr>> c_qcp, e_qcp;
    Fr zh_inv[8];
    Fr u, u2;
    std::vector<uint8_t> vk_bytes;
    std::vector<Aff> vk_points;        // S1 S2 S3 Ql Qr Qm Qo Qk Qcp*
    std::vector<Aff> vk_lin_points;    // the same commitments as the transcript binds them (used for [Lin])
    bool have_vk_points = false;

    // per-proof workspace
    DevBuf<Fr> L, R, O, cl, cr, co, cz, Zf, Zg, fscratch;
    DevBuf<Fr> el, er, eo, ez, eqk, h;
    std::vector<DevBuf<Fr>> c_pi2, e_pi2;
    DevBuf<Fr> lin, folded, quot, T, powz, powzi, powzw, powzwi;
    DevBuf<Fr> dot_partial, dot_out, small;
    DevBuf<uint32_t> small_idx;
    static constexpr int DOT_BLOCKS = 148 * 2;

    void set_profiling(bool on) override { prof.on = on; }
    // the evaluation buffers a shard group lets the other ranks write into / read from: el er eo ez h
    void shard_buffers(void** out5, uint64_t* n_out) override {
        out5[0] = el.p; out5[1] = er.p; out5[2] = eo.p; out5[3] = ez.p; out5[4] = h.p;
        *n_out = n;
    }
    void vk_commitments(void* out) override {
        if (!have_vk_points) {
            auto keep = vk_bytes;
            auto keep_pts = vk_lin_points;
            compute_vk();
            vk_bytes = keep;   // the caller-supplied transcript bytes stay authoritative
            vk_lin_points = keep_pts;
        }
        memcpy(out, vk_points.data(), vk_points.size() * sizeof(Aff));
    }

    size_t coeff_cap() const { return n + 4; }

    // Lagrange (natural, device, n) -> canonical (natural) in place
    void to_canonical(Fr* d) { d0.inverse_natural(d, st); }
    // canonical coefficients (len <= m) -> evaluations on the big coset (bit-reversed) into dst (m)
    void to_coset(Fr* dst, const Fr* coeffs, uint64_t len) {
        int id = prof.begin(B2P_STAT_NTT_MS, st);
        d1.coset_forward_dif_from(dst, coeffs, len, st);
        prof.end(id, st);
    }
    void upload_column_canonical(DevBuf<Fr>& c, const void* host_lagrange) {
        c.alloc(n);
        B2P_CUDA(cudaMemcpyAsync(c.p, host_lagrange, n * sizeof(Fr), cudaMemcpyHostToDevice, st));
        to_canonical(c.p);
    }

    void load(SrsBase* srs_base, uint64_t n_, uint32_t nb_public_, const void* ql, const void* qr, const void* qm,
              const void* qo, const void* qk, const int64_t* perm, uint32_t k_, const void* const* qcp,
              const uint64_t* cidx, const void* vkb, uint64_t vkb_len) override {
        curve = C::ID;
        B2P_REQUIRE(srs_base && srs_base->curve == C::ID, "SRS and circuit are on different curves");
        srs = static_cast<Srs<C>*>(srs_base);
        st = srs->stream;
        n = n_;
        nb_public = nb_public_;
        k = k_;
        B2P_REQUIRE(n >= 2 && (n & (n - 1)) == 0, "domain size must be a power of two >= 2");
        B2P_REQUIRE(srs->msm.npoints >= n + 3, "SRS holds fewer than n+3 points (setup.go:113-114)");
        B2P_REQUIRE(k <= MAX_QCP, "too many BSB22 commitments");
        B2P_REQUIRE(nb_public <= n, "more public inputs than rows");
        logn = 0;
        while ((1ull << logn) < n) logn++;
        log_rho = n < 6 ? 3 : 2;                     // gnark: 8n when sizeSystem < 6, else 4n
        logm = logn + log_rho;
        m = 1ull << logm;
        B2P_REQUIRE(logm <= Fr::Params::TWO_ADICITY, "domain too large for the scalar field");
        for (uint32_t c = 0; c < k; c++) {
            B2P_REQUIRE(nb_public + cidx[c] < n, "commitment constraint index out of range");
            commit_idx.push_back(cidx[c]);
        }
        d0.init(logn, false, st);
        d1.init(logm, true, st);
        u = d1.shift;
        u2 = u * u;

        // selectors: canonical + coset evaluations
        lag_qk.alloc(n);
        B2P_CUDA(cudaMemcpyAsync(lag_qk.p, qk, n * sizeof(Fr), cudaMemcpyHostToDevice, st));
        upload_column_canonical(c_ql, ql);
        upload_column_canonical(c_qr, qr);
        upload_column_canonical(c_qm, qm);
        upload_column_canonical(c_qo, qo);
        upload_column_canonical(c_qk, qk);
        c_qcp.resize(k);
        e_qcp.resize(k);
        for (uint32_t c = 0; c < k; c++) upload_column_canonical(c_qcp[c], qcp[c]);

        // permutation -> S1,S2,S3
        {
            DevBuf<int64_t> dperm(3 * n);
            B2P_CUDA(cudaMemcpyAsync(dperm.p, perm, 3 * n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
            std::vector<int64_t> chk(perm, perm + 3 * n);
            for (auto v : chk) B2P_REQUIRE(v >= 0 && (uint64_t)v < 3 * n, "permutation entry out of range");
            lag_S.alloc(3 * n);
            B2P_LAUNCH((k_perm_to_lagrange<Fr>), div_up(3 * n, 256), 256, 0, st, lag_S.p, dperm.p, d0.tw.p, n, u, u2);
            B2P_CUDA(cudaStreamSynchronize(st));
        }
        c_s1.alloc(n); c_s2.alloc(n); c_s3.alloc(n);
        Fr* cs[3] = {c_s1.p, c_s2.p, c_s3.p};
        for (int j = 0; j < 3; j++) {
            B2P_CUDA(cudaMemcpyAsync(cs[j], lag_S.p + j * n, n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
            to_canonical(cs[j]);
        }
        auto mk = [&](DevBuf<Fr>& e, const DevBuf<Fr>& c) { e.alloc(m); to_coset(e.p, c.p, n); };
        mk(e_ql, c_ql); mk(e_qr, c_qr); mk(e_qm, c_qm); mk(e_qo, c_qo); mk(e_qk, c_qk);
        mk(e_s1, c_s1); mk(e_s2, c_s2); mk(e_s3, c_s3);
        for (uint32_t c = 0; c < k; c++) mk(e_qcp[c], c_qcp[c]);

        // coset points, vanishing polynomial, L_1
        e_x.alloc(m);
        B2P_LAUNCH((k_coset_points<Fr>), div_up(m, 256), 256, 0, st, e_x.p, d1.tw.p, logm, d1.shift);
        const uint32_t rho = 1u << log_rho;
        ZhVals<Fr> zh;
        {
            Fr gn = d1.shift.pow_u64(n);                 // g^n
            Fr wr = d1.omega.pow_u64(n);                 // primitive rho-th root
            Fr cur = gn;
            for (uint32_t i = 0; i < rho; i++) {
                zh.v[i] = cur - Fr::one();
                zh_inv[i] = zh.v[i].inverse();
                cur = cur * wr;
            }
        }
        e_l1.alloc(m);
        B2P_LAUNCH((k_l1_denoms<Fr>), div_up(m, 256), 256, 0, st, e_l1.p, e_x.p, m, fr_from_u64(n));
        B2P_LAUNCH((k_batch_inverse<Fr>), div_up(div_up(m, BINV_CHUNK), 128), 128, 0, st, e_l1.p, m);
        B2P_LAUNCH((k_l1_finish<Fr>), div_up(m, 256), 256, 0, st, e_l1.p, m, logm, log_rho, zh);

        alloc_workspace();

        if (vkb) {
            B2P_REQUIRE(vkb_len == (size_t)(8 + k) * PB, "vk_transcript has the wrong length");
            vk_bytes.assign((const uint8_t*)vkb, (const uint8_t*)vkb + vkb_len);
            vk_lin_points.resize(8 + k);
            for (uint32_t i = 0; i < 8 + k; i++) vk_lin_points[i] = point_unmarshal<C>(&vk_bytes[(size_t)i * PB]);
        } else {
            compute_vk();
        }
        B2P_CUDA(cudaStreamSynchronize(st));
    }

    // B2P_NO_PI_DIRECT=1 forces the general completeQk path (tests compare the two)
    static bool env_no_pi_direct() {
        const char* e = getenv("B2P_NO_PI_DIRECT");
        return e && atoi(e) != 0;
    }
    static Fr fr_from_u64(uint64_t x) {
        Fr r = Fr::zero();
        r.v[0] = (uint32_t)x;
        r.v[1] = (uint32_t)(x >> 32);
        return r.to_mont();
    }

    ~Circuit() override {
        if (h_pub) cudaFreeHost(h_pub);
        if (copy_st) cudaStreamDestroy(copy_st);
        if (ev_ready) cudaEventDestroy(ev_ready);
        for (auto& e : ev_col) if (e) cudaEventDestroy(e);
    }
    void alloc_workspace() {
        B2P_CUDA(cudaMallocHost(&h_pub, (nb_public ? nb_public : 1) * sizeof(Fr)));
        B2P_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
        B2P_CUDA(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
        for (auto& e : ev_col) B2P_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        L.alloc(n); R.alloc(n); O.alloc(n);
        cl.alloc(coeff_cap()); cr.alloc(coeff_cap()); co.alloc(coeff_cap()); cz.alloc(coeff_cap());
        Zf.alloc(n); Zg.alloc(n);
        fscratch.alloc(fscan_scratch_elems(n + 4) + 8);
        el.alloc(m); er.alloc(m); eo.alloc(m); ez.alloc(m); eqk.alloc(m); h.alloc(m);
        c_pi2.resize(k); e_pi2.resize(k);
        for (uint32_t c = 0; c < k; c++) { c_pi2[c].alloc(n); e_pi2[c].alloc(m); }
        lin.alloc(coeff_cap()); folded.alloc(coeff_cap()); quot.alloc(coeff_cap()); T.alloc(coeff_cap());
        powz.alloc(coeff_cap()); powzi.alloc(coeff_cap()); powzw.alloc(coeff_cap()); powzwi.alloc(coeff_cap());
        dot_partial.alloc((size_t)MAX_DOT * DOT_BLOCKS);
        dot_out.alloc(MAX_DOT);
        small.alloc(64);
        small_idx.alloc(64);
    }

    // plonk.Setup's commitments (setup.go:149): S1 S2 S3 Ql Qr Qm Qo Qk Qcp*
    void compute_vk() {
        const Fr* cols[8] = {c_s1.p, c_s2.p, c_s3.p, c_ql.p, c_qr.p, c_qm.p, c_qo.p, c_qk.p};
        vk_points.resize(8 + k);
        for (int i = 0; i < 8; i++) srs->commit_async(cols[i], n, i);
        srs->fetch(0, 8, vk_points.data());
        for (uint32_t c = 0; c < k; c++) {
            srs->commit_async(c_qcp[c].p, n, 0);
            srs->fetch(0, 1, &vk_points[8 + c]);
        }
        vk_bytes.resize((size_t)(8 + k) * PB);
        for (size_t i = 0; i < vk_points.size(); i++) point_marshal<C>(vk_points[i], &vk_bytes[i * PB], true);
        have_vk_points = true;
        vk_lin_points = vk_points;
    }

    // ---- evaluation helpers ---------------------------------------------
    // dot products of polys with pow table; results to host (synchronises)
    void eval_many(const std::vector<std::pair<const Fr*, uint64_t>>& polys, const Fr* pow, Fr* host_out) {
        DotArgs<Fr> a;
        a.npoly = (int)polys.size();
        B2P_REQUIRE(a.npoly <= MAX_DOT, "too many polynomials in one evaluation batch");
        for (int i = 0; i < a.npoly; i++) { a.poly[i] = polys[i].first; a.len[i] = polys[i].second; }
        a.pow = pow;
        a.partial = dot_partial.p;
        dim3 grid(DOT_BLOCKS, a.npoly);
        B2P_LAUNCH((k_dot_pow<Fr>), grid, DOT_THREADS, 0, st, a);
        B2P_LAUNCH((k_dot_finish<Fr>), a.npoly, DOT_THREADS, 0, st, dot_partial.p, DOT_BLOCKS, dot_out.p);
        B2P_CUDA(cudaMemcpyAsync(host_out, dot_out.p, a.npoly * sizeof(Fr), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
    }
    void pow_table(Fr* out, const Fr& base, uint64_t len) {
        B2P_LAUNCH((k_pow_table<Fr>), div_up(div_up(len, 16), 128), 128, 0, st, out, base, len, Fr::one());
    }
    // quot = (p(X) - p(z)) / (X - z), len-1 coefficients; pow = z^j, powi = z^-j tables
    void divide_linear(const Fr* p, uint64_t len, const Fr* pow, const Fr* powi) {
        B2P_LAUNCH((k_mul_pointwise<Fr>), div_up(len, 256), 256, 0, st, T.p, p, pow, len);
        field_scan_inclusive<Fr, OpAdd>(T.p, len, fscratch.p, st);
        B2P_LAUNCH((k_div_finish<Fr>), div_up(len, 256), 256, 0, st, quot.p, T.p, powi, len);
    }

    // ---- the prover -------------------------------------------------------
    void* stream_handle() override { return (void*)st; }

    void prove(const void* hL, const void* hR, const void* hO, const void* const* h_pi2, const void* h_bsb22,
               const void* h_blinding, void* out_raw, bool device_inputs) override {
        auto t0 = std::chrono::steady_clock::now();
        for (auto& s : stats) s = 0;
        const unsigned long long launches0 = g_launch_count;
        // srs->prof points into this circuit: reset on EVERY exit (a throw below must not leave it dangling)
        struct ProfGuard {
            Srs<C>* s;
            ~ProfGuard() { s->prof = nullptr; s->msm.prof = nullptr; }
        } prof_guard{srs};
        srs->prof = prof.on ? &prof : nullptr;
        if (srs->router) srs->router->begin_proof();
        B2P_CUDA(cudaMemsetAsync(srs->msm.adds_total.p, 0, sizeof(unsigned long long), st));
        // wire columns: host buffers (the cgo path) or buffers already resident in HBM
        const cudaMemcpyKind in_kind = device_inputs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        // public inputs on the host (transcript, quotient): with device-resident columns they come back through a
        // pinned buffer; the copy is complete at the first fetch() below, before anything reads it
        if (device_inputs && nb_public)
            B2P_CUDA(cudaMemcpyAsync(h_pub, hL, nb_public * sizeof(Fr), cudaMemcpyDeviceToHost, st));
        const Fr* hLf = device_inputs ? h_pub : static_cast<const Fr*>(hL);
        const Aff* bsb = static_cast<const Aff*>(h_bsb22);
        B2P_REQUIRE(k == 0 || (h_pi2 && h_bsb22), "BSB22 inputs missing");

        // -- upload -----------------------------------------------------------
        // Host columns go up on a copy stream, one event per column: the transform and commitment of L
        // run while R and O are still on the wire (pinned host buffers; pageable ones are staged by the
        // driver and simply serialise).
        B2P_CUDA(cudaMemcpyAsync(small.p, h_blinding, 9 * sizeof(Fr), cudaMemcpyHostToDevice, st));
        const void* hcols[3] = {hL, hR, hO};
        Fr* dcols[3] = {L.p, R.p, O.p};
        cudaStream_t up = device_inputs ? st : copy_st;
        if (!device_inputs) {
            B2P_CUDA(cudaEventRecord(ev_ready, st));            // the previous proof is done with L, R, O, pi2
            B2P_CUDA(cudaStreamWaitEvent(copy_st, ev_ready, 0));
        }
        for (int j = 0; j < 3; j++) {
            B2P_CUDA(cudaMemcpyAsync(dcols[j], hcols[j], n * sizeof(Fr), in_kind, up));
            if (!device_inputs) B2P_CUDA(cudaEventRecord(ev_col[j], copy_st));
        }
        stats[B2P_STAT_H2D_BYTES] += (device_inputs ? 0.0 : 3.0 * n * sizeof(Fr)) + 9 * sizeof(Fr);
        for (uint32_t c = 0; c < k; c++) {
            B2P_CUDA(cudaMemcpyAsync(c_pi2[c].p, h_pi2[c], n * sizeof(Fr), in_kind, up));
            if (!device_inputs) stats[B2P_STAT_H2D_BYTES] += (double)n * sizeof(Fr);
        }
        if (!device_inputs) B2P_CUDA(cudaEventRecord(ev_col[3], copy_st));   // pi2 columns

        // -- round 1: l, r, o -------------------------------------------------
        Fr* wires_c[3] = {cl.p, cr.p, co.p};
        const Fr* wires_l[3] = {L.p, R.p, O.p};
        for (int j = 0; j < 3; j++) {
            if (!device_inputs) B2P_CUDA(cudaStreamWaitEvent(st, ev_col[j], 0));
            B2P_CUDA(cudaMemcpyAsync(wires_c[j], wires_l[j], n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
            B2P_CUDA(cudaMemsetAsync(wires_c[j] + n, 0, (coeff_cap() - n) * sizeof(Fr), st));
            int id = prof.begin(B2P_STAT_NTT_MS, st);
            to_canonical(wires_c[j]);
            prof.end(id, st);
            B2P_LAUNCH((k_blind<Fr>), 1, 32, 0, st, wires_c[j], n, 2, small.p + 2 * j);
            srs->commit_async(wires_c[j], n + 2, j);
        }
        if (!device_inputs) B2P_CUDA(cudaStreamWaitEvent(st, ev_col[3], 0));
        Aff pts[10];   // LRO[3], Z, H[3], batched H, zshift H, [Lin]
        srs->fetch(0, 3, pts);

        // -- transcript: gamma, beta -------------------------------------------
        uint8_t gamma_pre[32], beta_pre[32], alpha_pre[32], zeta_pre[32], pb[PB];
        std::vector<uint8_t> lro_bytes(3 * PB);
        for (int j = 0; j < 3; j++) point_marshal<C>(pts[j], &lro_bytes[j * PB], true);
        {
            Sha256 hsh;
            hsh.update("gamma");
            hsh.update(vk_bytes);
            for (uint32_t i = 0; i < nb_public; i++) {
                uint8_t b32[32];
                field_to_be(hLf[i], b32);
                hsh.update(b32, 32);
            }
            hsh.update(lro_bytes);
            hsh.final(gamma_pre);
            hsh.reset();
            hsh.update("beta");
            hsh.update(gamma_pre, 32);
            hsh.final(beta_pre);
        }
        const Fr gamma = fr_from_be32_mod<Fr>(gamma_pre), beta = fr_from_be32_mod<Fr>(beta_pre);

        // -- round 2: grand product Z -------------------------------------------
        B2P_LAUNCH((k_z_terms<Fr>), div_up(n, 256), 256, 0, st, Zf.p, Zg.p, L.p, R.p, O.p, lag_S.p, d0.tw.p, n, beta, gamma, u, u2);
        field_scan_inclusive<Fr, OpMul>(Zf.p, n, fscratch.p, st);
        field_scan_inclusive<Fr, OpMul>(Zg.p, n, fscratch.p, st);
        {
            Fr gall;   // product of all denominators; a zero denominator (probability ~2^-230) zeroes Z
            B2P_CUDA(cudaMemcpyAsync(&gall, Zg.p + (n - 1), sizeof(Fr), cudaMemcpyDeviceToHost, st));
            B2P_CUDA(cudaStreamSynchronize(st));
            B2P_LAUNCH((k_z_finish<Fr>), div_up(n, 256), 256, 0, st, cz.p, Zf.p, Zg.p, n, gall.inverse());
        }
        B2P_CUDA(cudaMemsetAsync(cz.p + n, 0, (coeff_cap() - n) * sizeof(Fr), st));
        {
            int id = prof.begin(B2P_STAT_NTT_MS, st);
            to_canonical(cz.p);
            prof.end(id, st);
        }
        B2P_LAUNCH((k_blind<Fr>), 1, 32, 0, st, cz.p, n, 3, small.p + 6);
        srs->commit_async(cz.p, n + 3, 3);
        srs->fetch(3, 1, pts + 3);

        // -- transcript: alpha ---------------------------------------------------
        std::vector<uint8_t> bsb_bytes((size_t)k * PB);
        std::vector<Fr> bsb_hash(k);
        for (uint32_t c = 0; c < k; c++) {
            point_marshal<C>(bsb[c], &bsb_bytes[c * PB], true);
            bsb_hash[c] = hash_fr<C>(&bsb_bytes[c * PB], PB);
        }
        {
            Sha256 hsh;
            hsh.update("alpha");
            hsh.update(beta_pre, 32);
            hsh.update(bsb_bytes);
            point_marshal<C>(pts[3], pb, true);
            hsh.update(pb, PB);
            hsh.final(alpha_pre);
        }
        const Fr alpha = fr_from_be32_mod<Fr>(alpha_pre);

        // -- round 3: quotient ------------------------------------------------------
        // qk completed with public inputs and commitment hashes (gnark completeQk): with few of them the
        // quotient kernel adds their Lagrange terms itself (QuotientArgs::n_pi), otherwise qk is completed
        // in Lagrange form and transformed (iNTT(n) + coset NTT(4n)) like the wires
        const bool pi_direct = nb_public + k <= (uint32_t)MAX_PI_DIRECT && !env_no_pi_direct();
        if (!pi_direct) {
            Fr* tmp = h.p;   // h is free until the quotient kernel writes it
            B2P_CUDA(cudaMemcpyAsync(tmp, lag_qk.p, n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
            if (nb_public) B2P_CUDA(cudaMemcpyAsync(tmp, L.p, nb_public * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
            if (k) {
                std::vector<uint32_t> idx(k);
                for (uint32_t c = 0; c < k; c++) idx[c] = (uint32_t)(nb_public + commit_idx[c]);
                B2P_CUDA(cudaMemcpyAsync(small_idx.p, idx.data(), k * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                B2P_CUDA(cudaMemcpyAsync(small.p + 16, bsb_hash.data(), k * sizeof(Fr), cudaMemcpyHostToDevice, st));
                B2P_LAUNCH((k_scatter_small<Fr>), 1, 32, 0, st, tmp, small_idx.p, small.p + 16, (int)k);
                B2P_CUDA(cudaStreamSynchronize(st));   // idx / bsb_hash are stack/host temporaries
            }
            int id = prof.begin(B2P_STAT_NTT_MS, st);
            to_canonical(tmp);
            prof.end(id, st);
            to_coset(eqk.p, tmp, n);
        }
        // the four big forward transforms: here, or spread over the ranks of a shard group (each rank's combine
        // kernel stores its block of the evaluations straight into el / er / eo / ez)
        const bool shard_ntt = srs->router && srs->router->shards_ntt();
        if (shard_ntt) {
            B2P_REQUIRE(pi_direct && k == 0, "sharded transforms: circuits with BSB22 commitments or more than 8 public "
                                             "inputs are proved with sharded commitments only (ntt_rows = 0)");
            int id = prof.begin(B2P_STAT_NTT_MS, st);
            srs->router->ntt_forward(cl.p, n + 2, 0, st);
            srs->router->ntt_forward(cr.p, n + 2, 1, st);
            srs->router->ntt_forward(co.p, n + 2, 2, st);
            srs->router->ntt_forward(cz.p, n + 3, 3, st);
            srs->router->ntt_forward_wait(st);
            prof.end(id, st);
        } else {
            to_coset(el.p, cl.p, n + 2);
            to_coset(er.p, cr.p, n + 2);
            to_coset(eo.p, co.p, n + 2);
            to_coset(ez.p, cz.p, n + 3);
        }
        for (uint32_t c = 0; c < k; c++) {
            int id = prof.begin(B2P_STAT_NTT_MS, st);
            to_canonical(c_pi2[c].p);
            prof.end(id, st);
            to_coset(e_pi2[c].p, c_pi2[c].p, n);
        }
        {
            QuotientArgs<Fr> a;
            a.l = el.p; a.r = er.p; a.o = eo.p; a.z = ez.p;
            a.ql = e_ql.p; a.qr = e_qr.p; a.qm = e_qm.p; a.qo = e_qo.p;
            a.qk = pi_direct ? e_qk.p : eqk.p;
            a.n_pi = 0;
            if (pi_direct) {
                for (uint32_t i = 0; i < nb_public; i++) { a.pi_row[a.n_pi] = i; a.pi_val[a.n_pi++] = hLf[i]; }
                for (uint32_t c = 0; c < k; c++) {
                    a.pi_row[a.n_pi] = (uint32_t)(nb_public + commit_idx[c]);
                    a.pi_val[a.n_pi++] = bsb_hash[c];
                }
            }
            a.s1 = e_s1.p; a.s2 = e_s2.p; a.s3 = e_s3.p; a.x = e_x.p; a.l1 = e_l1.p;
            for (uint32_t c = 0; c < k; c++) { a.qcp[c] = e_qcp[c].p; a.pi2[c] = e_pi2[c].p; }
            a.k = (int)k;
            a.h = h.p;
            a.logm = logm; a.log_rho = log_rho;
            a.beta = beta; a.gamma = gamma; a.alpha = alpha; a.alpha2 = alpha * alpha; a.u = u; a.u2 = u2;
            for (int i = 0; i < 8; i++) a.zh_inv[i] = zh_inv[i & ((1 << log_rho) - 1)];
            int id = prof.begin(B2P_STAT_QUOTIENT_MS, st);
            B2P_LAUNCH((k_quotient<Fr>), div_up(m, 256), 256, 0, st, a);
            prof.end(id, st);
            id = prof.begin(B2P_STAT_NTT_MS, st);
            if (shard_ntt) srs->router->ntt_inverse(3 * (n + 2), st);     // deg h = 3n + 5: nothing above is read
            else d1.coset_inverse_dit(h.p, st);
            prof.end(id, st);
        }
        for (int j = 0; j < 3; j++) srs->commit_async(h.p + (uint64_t)j * (n + 2), n + 2, 4 + j);
        srs->fetch(4, 3, pts + 4);

        // -- transcript: zeta ----------------------------------------------------------
        {
            Sha256 hsh;
            hsh.update("zeta");
            hsh.update(alpha_pre, 32);
            for (int j = 0; j < 3; j++) { point_marshal<C>(pts[4 + j], pb, true); hsh.update(pb, PB); }
            hsh.final(zeta_pre);
        }
        const Fr zeta = fr_from_be32_mod<Fr>(zeta_pre);
        B2P_REQUIRE(!zeta.is_zero(), "degenerate challenge zeta = 0");
        const Fr zw = zeta * d0.omega;

        // -- round 4: evaluations, opening of z at omega*zeta -----------------------------
        pow_table(powz.p, zeta, n + 3);
        pow_table(powzi.p, zeta.inverse(), n + 3);
        pow_table(powzw.p, zw, n + 3);
        pow_table(powzwi.p, zw.inverse(), n + 3);
        std::vector<Fr> ev(5 + k);   // l r o s1 s2 qcp*  (lin comes later)
        {
            std::vector<std::pair<const Fr*, uint64_t>> polys = {
                {cl.p, n + 2}, {cr.p, n + 2}, {co.p, n + 2}, {c_s1.p, n}, {c_s2.p, n}};
            for (uint32_t c = 0; c < k; c++) polys.push_back({c_qcp[c].p, n});
            // z(omega zeta) needs the other power table: queue first, fetch with the rest
            divide_linear(cz.p, n + 3, powzw.p, powzwi.p);   // T[n+2] = z(omega zeta) as a by-product
            B2P_CUDA(cudaMemcpyAsync(dot_out.p + MAX_DOT - 1, T.p + (n + 2), sizeof(Fr), cudaMemcpyDeviceToDevice, st));
            eval_many(polys, powz.p, ev.data());
        }
        Fr z_zw;
        B2P_CUDA(cudaMemcpyAsync(&z_zw, dot_out.p + MAX_DOT - 1, sizeof(Fr), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        // the opening MSM runs on the device while the host derives the linearisation and [Lin]
        srs->commit_async(quot.p, n + 2, 8);
        const Fr l_z = ev[0], r_z = ev[1], o_z = ev[2], s1_z = ev[3], s2_z = ev[4];

        // -- round 5: linearised polynomial --------------------------------------------------
        const Fr one = Fr::one();
        const Fr zn = zeta.pow_u64(n);
        const Fr zh_z = zn - one;
        const Fr l1_z = zh_z * d0.n_inv * (zeta - one).inverse();
        const Fr a2l = alpha * alpha * l1_z;
        const Fr s1p = alpha * beta * z_zw * (l_z + beta * s1_z + gamma) * (r_z + beta * s2_z + gamma);
        const Fr bz = beta * zeta;
        const Fr s2p = a2l - alpha * (l_z + bz + gamma) * (r_z + bz * u + gamma) * (o_z + bz * u2 + gamma);
        const Fr zn2 = zeta.pow_u64(n + 2);
        // [Lin] is the same combination of commitments that are already known: S1 S2 S3 Ql Qr Qm Qo Qk Qcp*
        // from the key, the BSB22 commitments, [Z] and [h_j] from this proof
        std::vector<Aff> lin_pts;
        std::vector<Fr> lin_coef;
        {
            LinCombArgs<Fr> a;
            int t = 0;
            auto term = [&](const Fr* p, uint64_t len, const Fr& coef, bool unit, const Aff& com) {
                a.poly[t] = p; a.len[t] = len; a.coef[t] = coef; a.unit[t] = unit ? 1 : 0; t++;
                lin_pts.push_back(com); lin_coef.push_back(coef);
            };
            term(c_ql.p, n, l_z, false, vk_lin_points[3]);
            term(c_qr.p, n, r_z, false, vk_lin_points[4]);
            term(c_qm.p, n, l_z * r_z, false, vk_lin_points[5]);
            term(c_qo.p, n, o_z, false, vk_lin_points[6]);
            term(c_qk.p, n, one, true, vk_lin_points[7]);
            for (uint32_t c = 0; c < k; c++) term(c_pi2[c].p, n, ev[5 + c], false, bsb[c]);
            term(c_s3.p, n, s1p, false, vk_lin_points[2]);
            term(cz.p, n + 3, s2p, false, pts[3]);
            const Fr mzh = zh_z.neg();
            term(h.p, n + 2, mzh, false, pts[4]);
            term(h.p + (n + 2), n + 2, mzh * zn2, false, pts[5]);
            term(h.p + 2 * (n + 2), n + 2, mzh * zn2 * zn2, false, pts[6]);
            a.nterms = t;
            a.out = lin.p;
            a.out_len = n + 3;
            B2P_LAUNCH((k_lincomb<Fr>), div_up(n + 3, 256), 256, 0, st, a);
        }
        pts[9] = host_msm_small<C>(lin_pts.data(), lin_coef.data(), (int)lin_pts.size());
        Fr lin_z;
        eval_many({{lin.p, n + 3}}, powz.p, &lin_z);
        srs->fetch(8, 1, pts + 8);   // W_{omega zeta}

        // -- fold challenge (kzg.BatchOpenSinglePoint; templateLogicSigBN254.go:280-286) --------
        uint8_t v_pre[32], b32[32];
        {
            Sha256 hsh;
            hsh.update("gamma");
            field_to_be(zeta, b32); hsh.update(b32, 32);
            point_marshal<C>(pts[9], pb, true); hsh.update(pb, PB);
            hsh.update(lro_bytes);
            hsh.update(vk_bytes.data(), 2 * PB);                          // S1, S2
            hsh.update(vk_bytes.data() + 8 * PB, (size_t)k * PB);         // Qcp*
            field_to_be(lin_z, b32); hsh.update(b32, 32);
            for (size_t i = 0; i < ev.size(); i++) { field_to_be(ev[i], b32); hsh.update(b32, 32); }
            field_to_be(z_zw, b32); hsh.update(b32, 32);
            hsh.final(v_pre);
        }
        const Fr v = fr_from_be32_mod<Fr>(v_pre);
        {
            LinCombArgs<Fr> a;
            int t = 0;
            Fr acc = one;
            auto term = [&](const Fr* p, uint64_t len) {
                a.poly[t] = p; a.len[t] = len; a.coef[t] = acc; a.unit[t] = (t == 0) ? 1 : 0; t++;
                acc = acc * v;
            };
            term(lin.p, n + 3);
            term(cl.p, n + 2);
            term(cr.p, n + 2);
            term(co.p, n + 2);
            term(c_s1.p, n);
            term(c_s2.p, n);
            for (uint32_t c = 0; c < k; c++) term(c_qcp[c].p, n);
            a.nterms = t;
            a.out = folded.p;
            a.out_len = n + 3;
            B2P_LAUNCH((k_lincomb<Fr>), div_up(n + 3, 256), 256, 0, st, a);
        }
        divide_linear(folded.p, n + 3, powz.p, powzi.p);
        srs->commit_async(quot.p, n + 2, 7);
        srs->fetch(7, 1, pts + 7);

        // -- output: 9 points then 7+k scalars, gnark in-memory layout ---------------------------
        uint8_t* out = static_cast<uint8_t*>(out_raw);
        memcpy(out, pts, 9 * sizeof(Aff));
        Fr* fo = reinterpret_cast<Fr*>(out + 9 * sizeof(Aff));
        fo[0] = lin_z;
        for (size_t i = 0; i < ev.size(); i++) fo[1 + i] = ev[i];
        fo[6 + k] = z_zw;
        stats[B2P_STAT_D2H_BYTES] += 10.0 * sizeof(Aff) + (8.0 + k) * sizeof(Fr);

        if (prof.on) prof.collect(stats);
        srs->prof = nullptr;
        stats[B2P_STAT_MSM_CALLS] = 9;
        {
            unsigned long long adds = 0;
            B2P_CUDA(cudaMemcpyAsync(&adds, srs->msm.adds_total.p, sizeof adds, cudaMemcpyDeviceToHost, st));
            B2P_CUDA(cudaStreamSynchronize(st));
            stats[B2P_STAT_MSM_ACCUM_ADDS] = (double)adds;
        }
        stats[B2P_STAT_LAUNCHES] = (double)(g_launch_count - launches0);
        stats[B2P_STAT_TOTAL_MS] =
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

// ---------------------------------------------------------------------------
// curve-erased entry points used by capi.cu
// ---------------------------------------------------------------------------
template <class C>
struct CurveOpsImpl : CurveOps {
    using Fr = typename C::Fr;
    using Aff = Affine<typename C::Fp>;
    SrsBase* new_srs() const override { return new Srs<C>(); }
    CircuitBase* new_circuit() const override { return new Circuit<C>(); }
    ShardGroupBase* new_shard_group(uint32_t world, uint32_t rank, uint64_t total, SrsBase* shard,
                                    uint64_t ntt_rows) const override;   // shard_group.cuh

    // b2p_ntt: natural order in and out
    void ntt(void* data, uint64_t n, int flags) const override {
        B2P_REQUIRE(n >= 1 && (n & (n - 1)) == 0, "NTT length must be a power of two");
        int logn = 0;
        while ((1ull << logn) < n) logn++;
        cudaStream_t st;
        B2P_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        try {
            NttDomain<Fr> d;
            const bool coset = flags & B2P_NTT_COSET;
            d.init(logn, coset, st);
            DevBuf<Fr> buf(n);
            B2P_CUDA(cudaMemcpyAsync(buf.p, data, n * sizeof(Fr), cudaMemcpyHostToDevice, st));
            if (flags & B2P_NTT_INVERSE) {
                d.bitrev(buf.p, st);
                if (coset) d.coset_inverse_dit(buf.p, st);
                else d.inverse_dit(buf.p, st);
            } else {
                if (coset) d.coset_forward_dif(buf.p, st);
                else d.forward_dif(buf.p, st);
                d.bitrev(buf.p, st);
            }
            B2P_CUDA(cudaMemcpyAsync(data, buf.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, st));
            B2P_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            cudaStreamDestroy(st);
            throw;
        }
        cudaStreamDestroy(st);
    }

    // sum of n affine points on the host (the G-point add that follows the all_gather of a sharded MSM)
    void g1_sum(const void* points, uint64_t n, void* out_affine) const override {
        const Aff* p = static_cast<const Aff*>(points);
        XYZZ<typename C::Fp> acc = XYZZ<typename C::Fp>::inf();
        for (uint64_t i = 0; i < n; i++) acc.add_affine_signed(p[i], false);
        Aff a = acc.to_affine();
        memcpy(out_affine, &a, sizeof a);
    }

    // helper.go:27-88 (and gnark's MarshalSolidity for BN254, helper.go:16-17): same field order on both curves
    void marshal_proof(uint32_t k, const void* raw, const void* bsb22, uint8_t* out) const override {
        constexpr int PB = 2 * C::Fp::N * 4;
        const Aff* pts = static_cast<const Aff*>(raw);
        const Fr* fr = reinterpret_cast<const Fr*>(static_cast<const uint8_t*>(raw) + 9 * sizeof(Aff));
        const Aff* bs = static_cast<const Aff*>(bsb22);
        uint8_t* o = out;
        // BLS12-381 points go through G1Affine.RawBytes() (helper.go:35): infinity is 0x40 then zeros; BN254 through
        // MarshalSolidity (helper.go:16-17): raw X || Y, infinity all zero.  point_marshal's flag only acts on BLS12-381.
        auto P = [&](const Aff& a) { point_marshal<C>(a, o, true); o += PB; };
        auto S = [&](const Fr& f) { field_to_be(f, o); o += 32; };
        for (int i = 0; i < 3; i++) P(pts[i]);          // LRO
        for (int i = 0; i < 3; i++) P(pts[4 + i]);      // H
        for (int i = 1; i < 6; i++) S(fr[i]);           // l r o s1 s2 at zeta (ClaimedValues[0] is not serialised)
        P(pts[3]);                                      // Z
        S(fr[6 + k]);                                   // z(omega zeta)
        P(pts[7]);                                      // BatchedProof.H
        P(pts[8]);                                      // ZShiftedOpening.H
        for (uint32_t i = 0; i < k; i++) S(fr[6 + i]);  // qcp_i(zeta)
        for (uint32_t i = 0; i < k; i++) P(bs[i]);      // Bsb22Commitments
    }
    // helper.go:91-110
    void marshal_public_inputs(const void* values, uint32_t nb_public, uint8_t* out) const override {
        const Fr* v = static_cast<const Fr*>(values);
        for (uint32_t i = 0; i < nb_public; i++) field_to_be(v[i], out + 32 * i);
    }
};

#ifndef B2P_INSTANTIATE_PROVER
extern template struct Srs<Bn254>;
extern template struct Srs<Bls12381>;
extern template struct Circuit<Bn254>;
extern template struct Circuit<Bls12381>;
#endif

}  // namespace b2p
