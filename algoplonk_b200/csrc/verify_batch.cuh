// plonk.Verify for a BATCH of proofs with the group arithmetic on the GPU (the self-check of
// (*CompiledCircuit).Verify, /root/reference/algoplonk.go:93, testutils/testutils.go:51, for a prover service that
// checks what it emitted).  Same acceptance condition as b2p_verify / b2p_verify_batch (verify_host.hpp, which
// restates verifier/templateLogicSigBN254.go:126-397): the host keeps what is sequential and tiny per proof -- SHA-256
// transcript, a dozen scalar operations, parsing -- and the three point combinations of every proof ([Lin]: 9+k
// points, the folded digest: 5+k, the pairing pair: 5 + 2, all with full-width scalars) go to the device as
// SEGMENTS: one warp per combination, one lane per (point, scalar) pair doing a double-and-add over XYZZ, a warp
// tree sum, lane 0 converts to affine.  Three launches per batch (the transcript hashes [Lin] before the fold
// challenge and the digest before the last one), then one host pairing check on the 128-bit-weighted sums.
// BLS12-381's r-torsion tests of the proof's points ([r]P = O, 255 doublings each: 1 ms per proof on a host core)
// ride along in the first launch as one-pair segments.
#pragma once
#include <algorithm>
#include <thread>
#include <vector>

#include "msm.cuh"
#include "verify_host.hpp"

namespace b2p {

// out[s] = sum over the pairs of segment s of scal * pts; scal: 8 little-endian words of PLAIN bits (not reduced,
// not Montgomery: the group order itself is a legal scalar here)
template <class Fp>
__global__ void __launch_bounds__(128) k_segment_lincomb(const Affine<Fp>* __restrict__ pts, const uint32_t* __restrict__ scal,
                                                         const uint32_t* __restrict__ seg_off, uint32_t nseg,
                                                         Affine<Fp>* __restrict__ out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nseg) return;
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    const uint32_t end = seg_off[warp + 1];
#pragma unroll 1
    for (uint32_t i = seg_off[warp] + lane; i < end; i += 32) {
        Affine<Fp> p;
        p.x = ld_field(&pts[i].x);
        p.y = ld_field(&pts[i].y);
        if (p.is_inf()) continue;
        uint32_t s[8];
#pragma unroll
        for (int w = 0; w < 8; w++) s[w] = scal[8 * (size_t)i + w];
        int top = 255;
        while (top >= 0 && !((s[top >> 5] >> (top & 31)) & 1u)) top--;
        XYZZ<Fp> r = XYZZ<Fp>::inf();
#pragma unroll 1
        for (int b = top; b >= 0; b--) {
            r = r.dbl();
            if ((s[b >> 5] >> (b & 31)) & 1u) r.add_affine(p.x, p.y);
        }
        xyzz_add(acc, r);
    }
    acc = warp_sum_xyzz(acc);
    if (lane == 0) {
        const Affine<Fp> a = acc.to_affine();
        st_field(&out[warp].x, a.x);
        st_field(&out[warp].y, a.y);
    }
}

template <class C, class PC>
struct DeviceBatchVerifier {
    using V = hp::HostVerifier<PC>;
    using HFr = typename V::Fr;
    using HFp = typename V::Fp;
    using HAff = typename V::Aff;
    using Staged = typename V::Staged;
    using DAff = Affine<typename C::Fp>;
    static_assert(sizeof(DAff) == 2 * sizeof(HFp), "host and device points share gnark's memory layout");

    // the requests of one stage, flattened
    struct Segments {
        std::vector<DAff> pts;
        std::vector<uint32_t> scal, off{0};
        void pair(const HAff& p, const HFr& mont) {
            const HFr c = mont.from_mont();
            raw(p, reinterpret_cast<const uint32_t*>(c.v));
        }
        void raw(const HAff& p, const uint32_t* words8) {
            DAff d;
            if (p.inf) { d = DAff::inf(); } else { memcpy(&d.x, p.x.v, sizeof d.x); memcpy(&d.y, p.y.v, sizeof d.y); }
            pts.push_back(d);
            scal.insert(scal.end(), words8, words8 + 8);
        }
        void one(const HAff& p) {
            static const uint32_t w[8] = {1, 0, 0, 0, 0, 0, 0, 0};
            raw(p, w);
        }
        void close() { off.push_back((uint32_t)pts.size()); }
        size_t count() const { return off.size() - 1; }
    };

    cudaStream_t st = nullptr;
    DevBuf<DAff> d_pts, d_out;
    DevBuf<uint32_t> d_scal, d_off;
    ~DeviceBatchVerifier() { if (st) cudaStreamDestroy(st); }

    std::vector<HAff> run(const Segments& sg) {
        const size_t ns = sg.count(), np = sg.pts.size();
        std::vector<HAff> res(ns);
        if (!ns) return res;
        if (!st) B2P_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (d_pts.n < np) { d_pts.alloc(np + np / 2); d_scal.alloc(8 * (np + np / 2)); }
        if (d_off.n < ns + 1) { d_off.alloc(ns + 1 + ns / 2); d_out.alloc(ns + ns / 2); }
        B2P_CUDA(cudaMemcpyAsync(d_pts.p, sg.pts.data(), np * sizeof(DAff), cudaMemcpyHostToDevice, st));
        B2P_CUDA(cudaMemcpyAsync(d_scal.p, sg.scal.data(), np * 32, cudaMemcpyHostToDevice, st));
        B2P_CUDA(cudaMemcpyAsync(d_off.p, sg.off.data(), (ns + 1) * 4, cudaMemcpyHostToDevice, st));
        B2P_LAUNCH((k_segment_lincomb<typename C::Fp>), div_up(ns * 32, 128), 128, 0, st, d_pts.p, d_scal.p, d_off.p,
                   (uint32_t)ns, d_out.p);
        std::vector<DAff> out(ns);
        B2P_CUDA(cudaMemcpyAsync(out.data(), d_out.p, ns * sizeof(DAff), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        for (size_t i = 0; i < ns; i++) {
            res[i].inf = out[i].is_inf();
            memcpy(res[i].x.v, &out[i].x, sizeof(HFp));
            memcpy(res[i].y.v, &out[i].y, sizeof(HFp));
        }
        return res;
    }

    // fn(i) for i < count on up to T host threads; exceptions inside fn are the caller's to avoid (fn catches)
    template <class Fn>
    static void parallel_for(uint64_t count, Fn fn) {
        unsigned T = 1;
        if (count >= 8) {
            const char* e = getenv("B2P_VERIFY_THREADS");
            const int asked = e ? atoi(e) : 0;
            const unsigned want = asked > 0 ? (unsigned)asked : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
            T = (unsigned)std::min<uint64_t>(std::min(64u, want), count);
        }
        auto work = [&](unsigned t) { for (uint64_t i = t; i < count; i += T) fn(i); };
        std::vector<std::thread> pool;
        unsigned started = 1;
        try {
            for (unsigned t = 1; t < T; t++) { pool.emplace_back(work, t); started = t + 1; }
        } catch (...) {}
        work(0);
        for (unsigned t = started; t < T; t++) work(t);
        for (auto& th : pool) th.join();
    }

    static constexpr uint64_t CHUNK = 4096;      // proofs whose staged state is alive at once (a few KB each)

    bool verify(const typename V::Key& vk, const uint8_t* proofs, uint64_t proof_len, const uint8_t* pubs, uint64_t pub_len,
                uint64_t count, uint64_t* bad, std::string* why) {
        if (bad) *bad = count;
        if (count == 0) return true;
        const std::vector<HFr> rho = V::batch_weights(proofs, proof_len, pubs, pub_len, count);
        typename V::Ext lhs = V::Ext::inf(), rhs = V::Ext::inf();
        for (uint64_t first = 0; first < count; first += CHUNK)
            if (!reduce_chunk(vk, proofs, proof_len, pubs, pub_len, first, std::min(CHUNK, count - first), rho, lhs, rhs, bad, why))
                return false;
        return V::pair_is_one(vk, lhs.to_affine(), V::neg(rhs.to_affine()), why);
    }

    // proofs [first, first + cnt): all pre-pairing checks, then lhs += sum rho_i lhs_i, rhs += sum rho_i (W_zeta + u W_omega-zeta)_i
    bool reduce_chunk(const typename V::Key& vk, const uint8_t* proofs, uint64_t proof_len, const uint8_t* pubs,
                      uint64_t pub_len, uint64_t first, uint64_t cnt, const std::vector<HFr>& rho, typename V::Ext& lhs,
                      typename V::Ext& rhs, uint64_t* bad, std::string* why) {
        std::vector<Staged> sts(cnt);
        std::vector<std::string> whys(cnt);
        std::vector<uint8_t> okv(cnt, 1);
        auto first_failure = [&]() -> bool {     // the smallest index wins, as in the host batch
            for (uint64_t i = 0; i < cnt; i++)
                if (!okv[i]) {
                    if (bad) *bad = first + i;
                    if (why) *why = whys[i];
                    return true;
                }
            return false;
        };
        // nothing may escape a worker thread: an exception there would be std::terminate across the C ABI
        auto guarded_stage = [&](uint64_t i, auto&& fn) {
            try { fn(); } catch (...) { okv[i] = 0; try { whys[i] = "internal error while reducing the proof (out of memory?)"; } catch (...) {} }
        };
        // ---- stage 1: parse, challenges, the scalars of [Lin]; r-torsion tests deferred to the device
        parallel_for(cnt, [&](uint64_t i) {
            guarded_stage(i, [&] {
                const uint64_t g = first + i;
                okv[i] = V::stage1(vk, proofs + g * proof_len, proof_len, pubs + g * pub_len, pub_len, sts[i], &whys[i], true);
            });
        });
        uint32_t order[8];
        for (int i = 0; i < 4; i++) {
            order[2 * i] = (uint32_t)HFr::M(i);
            order[2 * i + 1] = (uint32_t)(HFr::M(i) >> 32);
        }
        auto pack = [&](Segments& sg, uint64_t i) {       // a proof already rejected keeps its (empty) segment
            const Staged& s = sts[i];
            if (okv[i]) {
                for (size_t j = 0; j < s.pts.size(); j++) sg.pair(s.pts[j], s.sc[j]);
                for (const HAff& p : s.plus) sg.one(p);
            }
            sg.close();
        };
        Segments sg;
        std::vector<uint32_t> first_check(cnt + 1, 0);
        for (uint64_t i = 0; i < cnt; i++) pack(sg, i);
        for (uint64_t i = 0; i < cnt; i++) {                  // after the cnt combinations: one segment per point to test
            first_check[i] = (uint32_t)sg.count();
            if (okv[i])
                for (const HAff& p : sts[i].to_check) { sg.raw(p, order); sg.close(); }
        }
        first_check[cnt] = (uint32_t)sg.count();
        std::vector<HAff> res = run(sg);
        for (uint64_t i = 0; i < cnt; i++)
            for (uint32_t j = first_check[i]; j < first_check[i + 1]; j++)
                if (!res[j].inf) { okv[i] = 0; whys[i] = "a point of the proof is not in the r-torsion subgroup"; }
        if (first_failure()) return false;
        // ---- stage 2: fold challenge (hashes [Lin]), the folded digest
        parallel_for(cnt, [&](uint64_t i) { guarded_stage(i, [&] { V::stage2(vk, sts[i], res[i]); }); });
        if (first_failure()) return false;
        sg = Segments();
        for (uint64_t i = 0; i < cnt; i++) pack(sg, i);
        res = run(sg);
        // ---- stage 3: last challenge (hashes the digest), the pair of every proof, weighted by the batch's rho_i
        parallel_for(cnt, [&](uint64_t i) { guarded_stage(i, [&] { V::stage3(sts[i], res[i]); }); });
        if (first_failure()) return false;
        sg = Segments();
        for (uint64_t i = 0; i < cnt; i++) {
            const Staged& s = sts[i];
            const HFr& w = rho[first + i];
            for (size_t j = 0; j < s.pts.size(); j++) sg.pair(s.pts[j], s.sc[j] * w);
            for (const HAff& p : s.plus) sg.pair(p, w);
            sg.close();
            for (size_t j = 0; j < s.pts2.size(); j++) sg.pair(s.pts2[j], s.sc2[j] * w);
            for (const HAff& p : s.plus2) sg.pair(p, w);
            sg.close();
        }
        res = run(sg);
        for (uint64_t i = 0; i < cnt; i++) {
            lhs = lhs.add_affine(res[2 * i]);
            rhs = rhs.add_affine(res[2 * i + 1]);
        }
        return true;
    }
};

}  // namespace b2p
