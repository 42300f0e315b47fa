// Radix-2 NTT / iNTT over the scalar field -- replaces gnark-crypto fr/fft
// Domain.FFT / FFTInverse (+ coset variants).
//
// Forward = decimation in frequency: natural order in, bit-reversed order out.
// Inverse = decimation in time: bit-reversed order in, natural order out.
// Each launch ("pass") runs up to NTT_MAX_STAGES consecutive butterfly stages on
// a tile held in shared memory, so a 2^22 transform is two trips through HBM.
// Twiddles omega^k (k < n/2) are a resident table streamed through L2.
#pragma once
#include "common.cuh"

namespace b2p {

constexpr int NTT_MAX_STAGES = 11;   // tile = 2^11 elements = 64 KiB of shared memory

template <class Fr>
__global__ void k_pow_table(Fr* __restrict__ out, Fr base, uint64_t n, Fr scale) {
    // out[k] = scale * base^k ; each thread starts from base^(k0) by square-and-multiply
    constexpr int CHUNK = 16;
    uint64_t k0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * CHUNK;
    if (k0 >= n) return;
    Fr cur = base.pow_u64(k0) * scale;
    for (int j = 0; j < CHUNK && k0 + j < n; j++) {
        st_field(out + k0 + j, cur);
        cur = cur * base;
    }
}

static __device__ __forceinline__ uint32_t brev32(uint32_t x, int bits) { return __brev(x) >> (32 - bits); }

template <class Fr>
__global__ void k_bitrev_permute(Fr* __restrict__ data, int logn) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << logn)) return;
    uint32_t j = brev32(i, logn);
    if (logn == 0 || i >= j) return;
    Fr a = ld_field(data + i), b = ld_field(data + j);
    st_field(data + i, b);
    st_field(data + j, a);
}

// One pass: stages s_hi .. s_hi-nst+1 (DIF, descending) or s_lo .. s_lo+nst-1 (DIT, ascending).
// Tile element k (0 <= k < 2^nst) lives at global index (hi << (s_hi+1)) | (k << s_lo) | lo.
template <class Fr, bool DIF>
__global__ void __launch_bounds__(1 << (NTT_MAX_STAGES - 1))
k_ntt_pass(Fr* __restrict__ data, const Fr* __restrict__ tw, int logn, int s_lo, int nst) {
    extern __shared__ uint4 smem4[];
    constexpr int Q = Fr::N / 4;              // uint4 per element
    const int tile = 1 << nst;
    const uint32_t lo_mask = (1u << s_lo) - 1;
    const uint32_t lo = blockIdx.x & lo_mask;
    const uint32_t hi = blockIdx.x >> s_lo;
    const uint64_t base = ((uint64_t)hi << (s_lo + nst)) | lo;
    const int t = threadIdx.x;                // 2^(nst-1) threads

    // load: two elements per thread
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int k = t + r * (tile >> 1);
        const uint4* src = reinterpret_cast<const uint4*>(data + (base + ((uint64_t)k << s_lo)));
#pragma unroll
        for (int q = 0; q < Q; q++) smem4[q * tile + k] = src[q];
    }
    __syncthreads();

    for (int it = 0; it < nst; it++) {
        const int ls = DIF ? (nst - 1 - it) : it;
        const int half = 1 << ls;
        const int i = ((t >> ls) << (ls + 1)) | (t & (half - 1));
        const int j = i + half;
        // position inside the half-block at the global stage s = s_lo + ls
        const uint32_t pos = ((uint32_t)(i & (half - 1)) << s_lo) | lo;
        const int s = s_lo + ls;
        const Fr w = ldg_field(tw + ((uint64_t)pos << (logn - 1 - s)));
        Fr x, y;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            uint4 a = smem4[q * tile + i], b = smem4[q * tile + j];
            x.v[4 * q] = a.x; x.v[4 * q + 1] = a.y; x.v[4 * q + 2] = a.z; x.v[4 * q + 3] = a.w;
            y.v[4 * q] = b.x; y.v[4 * q + 1] = b.y; y.v[4 * q + 2] = b.z; y.v[4 * q + 3] = b.w;
        }
        Fr u, v;
        if (DIF) { u = x + y; v = (x - y) * w; }
        else     { Fr wy = y * w; u = x + wy; v = x - wy; }
#pragma unroll
        for (int q = 0; q < Q; q++) {
            smem4[q * tile + i] = make_uint4(u.v[4 * q], u.v[4 * q + 1], u.v[4 * q + 2], u.v[4 * q + 3]);
            smem4[q * tile + j] = make_uint4(v.v[4 * q], v.v[4 * q + 1], v.v[4 * q + 2], v.v[4 * q + 3]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int k = t + r * (tile >> 1);
        uint4* dst = reinterpret_cast<uint4*>(data + (base + ((uint64_t)k << s_lo)));
#pragma unroll
        for (int q = 0; q < Q; q++) dst[q] = smem4[q * tile + k];
    }
}

// out[i] = a[i] * table[i]   (coset scaling: table = g^i, or g^-i / n)
template <class Fr>
__global__ void k_mul_table(Fr* __restrict__ a, const Fr* __restrict__ table, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_field(a + i, ld_field(a + i) * ldg_field(table + i));
}
template <class Fr>
__global__ void k_mul_scalar(Fr* __restrict__ a, Fr s, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_field(a + i, ld_field(a + i) * s);
}

template <class Fr>
struct NttDomain {
    int logn = 0;
    uint64_t n = 0;
    Fr omega, omega_inv, n_inv;        // host copies (Montgomery form)
    Fr shift, shift_inv;               // coset generator g (FrMultiplicativeGen) and 1/g
    DevBuf<Fr> tw, tw_inv;             // omega^k, omega^-k, k < n/2
    DevBuf<Fr> coset_pow;              // g^i
    DevBuf<Fr> coset_pow_inv;          // g^-i / n
    bool has_coset = false;

    static Fr host_const(uint32_t (*f)(int)) {
        Fr r;
        for (int i = 0; i < Fr::N; i++) r.v[i] = f(i);
        return r;
    }
    static Fr root_of_unity(int logn) {
        Fr w = host_const(&Fr::Params::root);
        for (int i = logn; i < Fr::Params::TWO_ADICITY; i++) w = w.sqr();
        return w;
    }

    void init(int logn_, bool with_coset, cudaStream_t st) {
        B2P_REQUIRE(logn_ >= 0 && logn_ <= Fr::Params::TWO_ADICITY, "NTT size exceeds the field's 2-adicity");
        logn = logn_;
        n = 1ull << logn;
        omega = root_of_unity(logn);
        omega_inv = omega.inverse();
        n_inv = Fr::from_u32(2).pow_u64(logn).inverse();
        shift = host_const(&Fr::Params::shift);
        shift_inv = shift.inverse();
        const uint64_t half = n > 1 ? n / 2 : 1;
        tw.alloc(half);
        tw_inv.alloc(half);
        B2P_LAUNCH((k_pow_table<Fr>), div_up(div_up(half, 16), 128), 128, 0, st, tw.p, omega, half, Fr::one());
        B2P_LAUNCH((k_pow_table<Fr>), div_up(div_up(half, 16), 128), 128, 0, st, tw_inv.p, omega_inv, half, Fr::one());
        has_coset = with_coset;
        if (with_coset) {
            coset_pow.alloc(n);
            coset_pow_inv.alloc(n);
            B2P_LAUNCH((k_pow_table<Fr>), div_up(div_up(n, 16), 128), 128, 0, st, coset_pow.p, shift, n, Fr::one());
            B2P_LAUNCH((k_pow_table<Fr>), div_up(div_up(n, 16), 128), 128, 0, st, coset_pow_inv.p, shift_inv, n, n_inv);
        }
    }

    template <bool DIF>
    void passes(Fr* d, const Fr* table, cudaStream_t st) const {
        if (logn == 0) return;
        const int npass = (logn + NTT_MAX_STAGES - 1) / NTT_MAX_STAGES;
        // split stages as evenly as possible
        int sizes[8];
        for (int p = 0; p < npass; p++) sizes[p] = logn / npass + (p < logn % npass ? 1 : 0);
        auto kern = k_ntt_pass<Fr, DIF>;
        static bool attr_set = false;
        if (!attr_set) {
            B2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(Fr) << NTT_MAX_STAGES)));
            attr_set = true;
        }
        if (DIF) {
            int s_hi = logn - 1;
            for (int p = 0; p < npass; p++) {
                const int nst = sizes[p], s_lo = s_hi - nst + 1;
                B2P_LAUNCH(kern, (unsigned)(n >> nst), 1 << (nst - 1), sizeof(Fr) << nst, st, d, table, logn, s_lo, nst);
                s_hi = s_lo - 1;
            }
        } else {
            int s_lo = 0;
            for (int p = 0; p < npass; p++) {
                const int nst = sizes[p];
                B2P_LAUNCH(kern, (unsigned)(n >> nst), 1 << (nst - 1), sizeof(Fr) << nst, st, d, table, logn, s_lo, nst);
                s_lo += nst;
            }
        }
    }

    void bitrev(Fr* d, cudaStream_t st) const {
        B2P_LAUNCH((k_bitrev_permute<Fr>), div_up(n, 256), 256, 0, st, d, logn);
    }
    // natural -> bit-reversed evaluations
    void forward_dif(Fr* d, cudaStream_t st) const { passes<true>(d, tw.p, st); }
    // bit-reversed evaluations -> natural coefficients (scaled by 1/n)
    void inverse_dit(Fr* d, cudaStream_t st) const {
        passes<false>(d, tw_inv.p, st);
        B2P_LAUNCH((k_mul_scalar<Fr>), div_up(n, 256), 256, 0, st, d, n_inv, n);
    }
    // natural -> natural variants (explicit bit reversal)
    void forward_natural(Fr* d, cudaStream_t st) const {
        forward_dif(d, st);
        B2P_LAUNCH((k_bitrev_permute<Fr>), div_up(n, 256), 256, 0, st, d, logn);
    }
    void inverse_natural(Fr* d, cudaStream_t st) const {
        B2P_LAUNCH((k_bitrev_permute<Fr>), div_up(n, 256), 256, 0, st, d, logn);
        inverse_dit(d, st);
    }
    // coefficients (natural) -> evaluations on g*<omega> in bit-reversed order
    void coset_forward_dif(Fr* d, cudaStream_t st) const {
        B2P_REQUIRE(has_coset, "domain built without coset tables");
        B2P_LAUNCH((k_mul_table<Fr>), div_up(n, 256), 256, 0, st, d, coset_pow.p, n);
        forward_dif(d, st);
    }
    // evaluations on g*<omega> (bit-reversed) -> coefficients (natural)
    void coset_inverse_dit(Fr* d, cudaStream_t st) const {
        B2P_REQUIRE(has_coset, "domain built without coset tables");
        passes<false>(d, tw_inv.p, st);
        B2P_LAUNCH((k_mul_table<Fr>), div_up(n, 256), 256, 0, st, d, coset_pow_inv.p, n);
    }
};

#ifndef B2P_INSTANTIATE_NTT
extern template struct NttDomain<FrBn254>;
extern template struct NttDomain<FrBls12381>;
#endif

}  // namespace b2p
