// NTT with the domain sharded over the G = 2^g GPUs of one box (BASELINE config 5):
// the multi-GPU form of gnark-crypto's fft.Domain.FFT / FFTInverse.  The reference is
// single-process, so nothing in /root/reference is replaced one to one; single-GPU semantics are b2p_ntt's.
//
// Distribution (one exchange per transform, in either direction):
//   coefficients  a_i           : CYCLIC   -- rank r holds a_{jG+r} at local index j, j < n/G
//   evaluations   A(w^brev(p))  : BLOCKS of the bit-reversed order -- rank r holds positions
//                                 p in [r n/G, (r+1) n/G), i.e. the layout the single-GPU DIF produces, cut in G
//   (both are "index mod G" distributions of the natural orders, so coefficient-wise and point-wise kernels
//    and the point-set-sharded MSM run on them with no communication.)
//
// Forward, k = k1 + (n/G) k2:   A(w^k) = sum_r  w_G^(r k2) * [ w_n^(r k1) * C_r(k1) ],   C_r = NTT_{n/G}(shard r)
//   1. local   : rank r runs the ordinary DIF passes of ntt.cuh on its shard -> C_r in bit-reversed order,
//                position q = brev(k1).  The top g bits of q name the rank that needs C_r(k1): the result
//                is already cut into G contiguous chunks of n/G^2, chunk d for rank d.
//   2. combine : ONE kernel on rank d reads chunk d of every rank's buffer (peer memory over NVLink, or the
//                receive buffer of an all_to_all), multiplies by w_n^(r k1), runs the size-G DIF over r in
//                registers and writes G neighbouring outputs: local position (q_low << g) | brev_g(k2).
//                The transposition is the kernel's load pattern; there is no separate transpose pass.
// Inverse = the mirror image: split (size-G DIT over the G neighbours, times w_n^(-r k1), chunk r written to
//   rank r: peer stores, or the send buffer of an all_to_all), then the ordinary DIT passes on the n/G local
//   elements with the 1/n (or g^-i / n) of the whole transform fused into the last pass.
//
// The per-column bodies live in ntt_shard_math.cuh (host + device) so tests/csrc/ntt_shard_shim.cpp runs
// the very same index and twiddle code on the host.
#pragma once
#include "ntt.cuh"
#include "ntt_shard_math.cuh"

namespace b2p {

template <class Fr>
struct NttShardArgs {
    Fr* chunk[1 << NTT_SHARD_MAX_LOGG];   // chunk[r]: n/G^2 elements exchanged with rank r
    const Fr* tw1;                        // w_n^(+-k1(first + q)), q < n/G^2
    Fr wg[1 << (NTT_SHARD_MAX_LOGG - 1)]; // w_G^(+-j), j < G/2
    Fr* local;                            // n/G local evaluations (written by combine, read by split)
    uint32_t chunk_len;
};

template <class Fr>
__global__ void k_ntt_shard_twiddles(Fr* __restrict__ out, Fr base, uint64_t first, uint32_t count, int local_logn) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    st_field(out + q, base.pow_u64(shard_k1(first + q, local_logn)));
}

// One thread per exchanged column q: G strided loads (coalesced across the warp, one per source rank), the
// twiddles and the size-G butterflies in registers, G neighbouring stores.  HBM/NVLink-bound:
// 2 * 32 B * n/G of traffic per rank and 2(G-1) + (G/2) log2 G multiplications per G elements.
template <class Fr, int LOGG>
__global__ void __launch_bounds__(128) k_ntt_shard_combine(const NttShardArgs<Fr> a) {
    constexpr int G = 1 << LOGG;
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.chunk_len) return;
    Fr e[G];
#pragma unroll
    for (int r = 0; r < G; r++) e[r] = ld_field(a.chunk[r] + q);
    if (LOGG > 0) shard_combine_body<Fr, LOGG>(e, ldg_field(a.tw1 + q), a.wg);
    Fr* out = a.local + ((uint64_t)q << LOGG);
#pragma unroll
    for (int t = 0; t < G; t++) st_field(out + t, e[t]);
}
template <class Fr, int LOGG>
__global__ void __launch_bounds__(128) k_ntt_shard_split(const NttShardArgs<Fr> a) {
    constexpr int G = 1 << LOGG;
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.chunk_len) return;
    Fr e[G];
    const Fr* in = a.local + ((uint64_t)q << LOGG);
#pragma unroll
    for (int t = 0; t < G; t++) e[t] = ld_field(in + t);
    if (LOGG > 0) shard_split_body<Fr, LOGG>(e, ldg_field(a.tw1 + q), a.wg);
#pragma unroll
    for (int r = 0; r < G; r++) st_field(a.chunk[r] + q, e[r]);
}

template <class Fr>
struct NttShard : NttShardBase {
    int logn = 0, logg = 0;
    uint32_t rank = 0, world = 1;
    uint64_t n = 0, local_n = 0, chunk_len = 0;
    NttDomain<Fr> dom;                    // the local transform: size n/G, root w_n^G
    DevBuf<Fr> tw1, tw1_inv;              // chunk_len entries each
    DevBuf<Fr> coset_pow, coset_pow_inv;  // g^(jG+r) and g^-(jG+r) / n, j < n/G
    Fr wg[1 << (NTT_SHARD_MAX_LOGG - 1)], wgi[1 << (NTT_SHARD_MAX_LOGG - 1)];
    Fr n_inv;                             // 1/n of the WHOLE transform

    void init(uint64_t n_, uint32_t world_, uint32_t rank_) {
        B2P_REQUIRE(n_ >= 1 && (n_ & (n_ - 1)) == 0, "NTT length must be a power of two");
        B2P_REQUIRE(world_ >= 1 && (world_ & (world_ - 1)) == 0 && world_ <= (1u << NTT_SHARD_MAX_LOGG),
                    "world size must be 1, 2, 4 or 8");
        B2P_REQUIRE(rank_ < world_, "rank out of range");
        while ((1ull << logn) < n_) logn++;
        while ((1u << logg) < world_) logg++;
        B2P_REQUIRE(logn <= Fr::Params::TWO_ADICITY, "NTT size exceeds the field's 2-adicity");
        B2P_REQUIRE(logn >= 2 * logg, "sharded NTT needs n >= world^2 (one exchanged element per pair of ranks)");
        n = n_; world = world_; rank = rank_;
        local_n = n >> logg;
        chunk_len = local_n >> logg;
        B2P_REQUIRE(chunk_len <= 0xffffffffull, "shard too large");
        cudaStream_t st;
        B2P_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        try {
            dom.init(logn - logg, false, st);
            const Fr w = NttDomain<Fr>::root_of_unity(logn), wi = w.inverse();
            const Fr g = NttDomain<Fr>::host_const(&Fr::Params::shift), gi = g.inverse();
            n_inv = Fr::from_u32(2).pow_u64(logn).inverse();
            const Fr w_g = w.pow_u64(local_n), w_gi = wi.pow_u64(local_n);     // primitive G-th roots
            for (int j = 0; j < (1 << (NTT_SHARD_MAX_LOGG - 1)); j++) { wg[j] = w_g.pow_u64(j); wgi[j] = w_gi.pow_u64(j); }
            tw1.alloc(chunk_len);
            tw1_inv.alloc(chunk_len);
            const uint64_t first = (uint64_t)rank * chunk_len;
            B2P_LAUNCH((k_ntt_shard_twiddles<Fr>), div_up(chunk_len, 128), 128, 0, st, tw1.p, w, first,
                       (uint32_t)chunk_len, logn - logg);
            B2P_LAUNCH((k_ntt_shard_twiddles<Fr>), div_up(chunk_len, 128), 128, 0, st, tw1_inv.p, wi, first,
                       (uint32_t)chunk_len, logn - logg);
            coset_pow.alloc(local_n);
            coset_pow_inv.alloc(local_n);
            B2P_LAUNCH((k_pow_table<Fr>), div_up(div_up(local_n, 16), 128), 128, 0, st, coset_pow.p, g.pow_u64(world),
                       local_n, g.pow_u64(rank));
            B2P_LAUNCH((k_pow_table<Fr>), div_up(div_up(local_n, 16), 128), 128, 0, st, coset_pow_inv.p,
                       gi.pow_u64(world), local_n, gi.pow_u64(rank) * n_inv);
            B2P_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            cudaStreamDestroy(st);
            throw;
        }
        cudaStreamDestroy(st);
    }

    uint64_t local_size() const override { return local_n; }
    uint64_t chunk_size() const override { return chunk_len; }

    void forward_local(const void* d_coeffs, uint64_t local_len, int flags, void* d_x, void* stream) const override {
        B2P_REQUIRE(local_len <= local_n, "more coefficients than local domain points");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        Fr* x = static_cast<Fr*>(d_x);
        const Fr* src = static_cast<const Fr*>(d_coeffs);
        if (flags & B2P_NTT_COSET) {
            dom.template passes<true, NTT_LOAD_PAD_SCALE, NTT_PLAIN>(x, dom.tw.p, st, src, local_len, coset_pow.p);
            return;
        }
        if (src != x && local_len)
            B2P_CUDA(cudaMemcpyAsync(x, src, local_len * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        if (local_len < local_n) B2P_CUDA(cudaMemsetAsync(x + local_len, 0, (local_n - local_len) * sizeof(Fr), st));
        dom.forward_dif(x, st);
    }

    template <int LOGG, bool COMBINE>
    void launch_exchange(const NttShardArgs<Fr>& a, cudaStream_t st) const {
        if (COMBINE) B2P_LAUNCH((k_ntt_shard_combine<Fr, LOGG>), div_up(chunk_len, 128), 128, 0, st, a);
        else B2P_LAUNCH((k_ntt_shard_split<Fr, LOGG>), div_up(chunk_len, 128), 128, 0, st, a);
    }
    template <bool COMBINE>
    void exchange(void* const* d_chunks, void* d_local, cudaStream_t st) const {
        NttShardArgs<Fr> a;
        for (uint32_t r = 0; r < (1u << NTT_SHARD_MAX_LOGG); r++) {
            B2P_REQUIRE(r >= world || d_chunks[r], "null chunk pointer");
            a.chunk[r] = r < world ? static_cast<Fr*>(d_chunks[r]) : nullptr;
        }
        a.tw1 = COMBINE ? tw1.p : tw1_inv.p;
        for (int j = 0; j < (1 << (NTT_SHARD_MAX_LOGG - 1)); j++) a.wg[j] = COMBINE ? wg[j] : wgi[j];
        a.local = static_cast<Fr*>(d_local);
        a.chunk_len = (uint32_t)chunk_len;
        switch (logg) {
            case 0: launch_exchange<0, COMBINE>(a, st); break;
            case 1: launch_exchange<1, COMBINE>(a, st); break;
            case 2: launch_exchange<2, COMBINE>(a, st); break;
            default: launch_exchange<3, COMBINE>(a, st); break;
        }
    }
    void forward_combine(const void* const* d_chunks, void* d_out, void* stream) const override {
        exchange<true>(const_cast<void* const*>(d_chunks), d_out, static_cast<cudaStream_t>(stream));
    }
    void inverse_split(const void* d_evals, void* const* d_chunks, void* stream) const override {
        exchange<false>(d_chunks, const_cast<void*>(d_evals), static_cast<cudaStream_t>(stream));
    }
    void inverse_local(void* d_x, int flags, void* d_out, void* stream) const override {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        Fr* x = static_cast<Fr*>(d_x);
        if (flags & B2P_NTT_COSET)
            dom.template passes<false, NTT_PLAIN, NTT_STORE_SCALE_TAB>(x, dom.tw_inv.p, st, nullptr, 0, coset_pow_inv.p);
        else
            dom.template passes<false, NTT_PLAIN, NTT_STORE_SCALE>(x, dom.tw_inv.p, st, nullptr, 0, nullptr, &n_inv);
        if (d_out && d_out != d_x)
            B2P_CUDA(cudaMemcpyAsync(d_out, d_x, local_n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    }
};
}  // namespace b2p
