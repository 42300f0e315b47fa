// Radix-2 NTT / iNTT over the scalar field -- replaces gnark-crypto fr/fft
// Domain.FFT / FFTInverse (+ coset variants).
//
// Forward = decimation in frequency: natural order in, bit-reversed order out.
// Inverse = decimation in time: bit-reversed order in, natural order out.
// Each launch ("pass") runs up to NTT_MAX_STAGES consecutive butterfly stages on
// a tile held in shared memory, so a 2^22 transform is two trips through HBM; zero
// padding, coset scaling and the 1/n of the inverse are fused into the first / last pass.
// Twiddles omega^k (k < n/2) are a resident table streamed through L2.
#pragma once
#include <cstdlib>
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

// What a pass does besides its butterflies (fused so the scaling steps of the coset transforms and
// the zero padding of the coefficient vectors cost no extra trip through HBM):
enum NttFuse {
    NTT_PLAIN = 0,
    NTT_LOAD_PAD_SCALE = 1,   // first DIF pass: in[i] = i < src_len ? src[i] * scale_tab[i] : 0   (out of place)
    NTT_STORE_SCALE_TAB = 2,  // last DIT pass:  out[i] = v[i] * scale_tab[i]                       (g^-i / n)
    NTT_STORE_SCALE = 3,      // last DIT pass:  out[i] = v[i] * scale                              (1 / n)
};

template <class Fr>
struct NttPassArgs {
    Fr* data;                 // in place (destination of NTT_LOAD_PAD_SCALE)
    const Fr* tw;             // omega^k, k < n/2
    const Fr* src;            // NTT_LOAD_PAD_SCALE: coefficient vector
    uint64_t src_len;
    const Fr* scale_tab;
    Fr scale;
    int logn, s_lo, nst;
};

// One pass: stages s_hi .. s_hi-nst+1 (DIF, descending) or s_lo .. s_lo+nst-1 (DIT, ascending).
// Tile element k (0 <= k < 2^nst) lives at global index (hi << (s_hi+1)) | (k << s_lo) | lo.
template <class Fr, bool DIF, int FUSE>
__global__ void __launch_bounds__(1 << (NTT_MAX_STAGES - 1))
k_ntt_pass(const NttPassArgs<Fr> a) {
    extern __shared__ uint4 smem4[];
    constexpr int Q = Fr::N / 4;              // uint4 per element
    Fr* __restrict__ data = a.data;
    const Fr* __restrict__ tw = a.tw;
    const int logn = a.logn, s_lo = a.s_lo, nst = a.nst;
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
        const uint64_t g = base + ((uint64_t)k << s_lo);
        if (FUSE == NTT_LOAD_PAD_SCALE) {
            Fr v = Fr::zero();
            if (g < a.src_len) v = ld_field(a.src + g) * ldg_field(a.scale_tab + g);
#pragma unroll
            for (int q = 0; q < Q; q++)
                smem4[q * tile + k] = make_uint4(v.v[4 * q], v.v[4 * q + 1], v.v[4 * q + 2], v.v[4 * q + 3]);
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(data + g);
#pragma unroll
            for (int q = 0; q < Q; q++) smem4[q * tile + k] = src[q];
        }
    }
    __syncthreads();

    // twiddle of this thread's butterfly at iteration `it` (depends on the thread, not on the data):
    // fetched one stage ahead so its L2 latency hides behind the butterfly and the barrier
    auto twiddle = [&](int it) {
        const int ls = DIF ? (nst - 1 - it) : it;
        const int half = 1 << ls;
        // position inside the half-block at the global stage s = s_lo + ls
        const uint32_t pos = ((uint32_t)(t & (half - 1)) << s_lo) | lo;
        return ldg_field(tw + ((uint64_t)pos << (logn - 1 - (s_lo + ls))));
    };
    Fr w = twiddle(0);
    for (int it = 0; it < nst; it++) {
        const int ls = DIF ? (nst - 1 - it) : it;
        const int half = 1 << ls;
        const int i = ((t >> ls) << (ls + 1)) | (t & (half - 1));
        const int j = i + half;
        Fr w_next;
        if (it + 1 < nst) w_next = twiddle(it + 1);
        Fr x, y;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            uint4 a4 = smem4[q * tile + i], b4 = smem4[q * tile + j];
            x.v[4 * q] = a4.x; x.v[4 * q + 1] = a4.y; x.v[4 * q + 2] = a4.z; x.v[4 * q + 3] = a4.w;
            y.v[4 * q] = b4.x; y.v[4 * q + 1] = b4.y; y.v[4 * q + 2] = b4.z; y.v[4 * q + 3] = b4.w;
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
        if (it + 1 < nst) w = w_next;
    }

#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int k = t + r * (tile >> 1);
        const uint64_t g = base + ((uint64_t)k << s_lo);
        uint4* dst = reinterpret_cast<uint4*>(data + g);
        if (FUSE == NTT_STORE_SCALE_TAB || FUSE == NTT_STORE_SCALE) {
            Fr v;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                uint4 c = smem4[q * tile + k];
                v.v[4 * q] = c.x; v.v[4 * q + 1] = c.y; v.v[4 * q + 2] = c.z; v.v[4 * q + 3] = c.w;
            }
            v = v * (FUSE == NTT_STORE_SCALE_TAB ? ldg_field(a.scale_tab + g) : a.scale);
            st_field(data + g, v);
        } else {
#pragma unroll
            for (int q = 0; q < Q; q++) dst[q] = smem4[q * tile + k];
        }
    }
}

// ---------------------------------------------------------------------------
// The same pass with eight elements per thread: three butterfly stages run in registers between two
// trips through shared memory (the kernel above makes one trip per stage), the four butterflies of a
// stage are independent (ILP for the IMAD pipe, which is what bounds a 254-bit NTT), and a thread
// fetches 7 twiddles per 12 butterflies instead of 12.  Used for tiles of >= 2^8 elements.
//   round: a window of three bits [p, p+3) of the tile index; thread t holds the elements whose
//   index is t with j = 0..7 spliced in at bit p; the stages of the round are the live bits of the
//   window (all three, except in the last round when nst is not a multiple of 3).
// Shared-memory slot of element k is k ^ ((k >> 3) & 7): conflict-free 128-bit accesses both for
// the window at p = 0 (a thread's elements are neighbours) and for the higher ones.
// ---------------------------------------------------------------------------
constexpr int NTT_R8_MIN_STAGES = 8;
__device__ __forceinline__ int ntt_swz(int k) { return k ^ ((k >> 3) & 7); }

template <class Fr>
__device__ __forceinline__ Fr ntt_lds(const uint4* smem4, int tile, int k) {
    constexpr int Q = Fr::N / 4;
    Fr x;
    const int sk = ntt_swz(k);
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const uint4 c = smem4[q * tile + sk];
        x.v[4 * q] = c.x; x.v[4 * q + 1] = c.y; x.v[4 * q + 2] = c.z; x.v[4 * q + 3] = c.w;
    }
    return x;
}
template <class Fr>
__device__ __forceinline__ void ntt_sts(uint4* smem4, int tile, int k, const Fr& x) {
    constexpr int Q = Fr::N / 4;
    const int sk = ntt_swz(k);
#pragma unroll
    for (int q = 0; q < Q; q++)
        smem4[q * tile + sk] = make_uint4(x.v[4 * q], x.v[4 * q + 1], x.v[4 * q + 2], x.v[4 * q + 3]);
}

// butterflies of window bit B on the 8 register-resident elements
template <class Fr, bool DIF, int B>
__device__ __forceinline__ void ntt_r8_stage(Fr (&e)[8], const Fr* __restrict__ tw, uint32_t tl, int p, int s_lo,
                                             uint32_t lo, int logn) {
    const int s = s_lo + p + B;               // global stage
#pragma unroll
    for (int grp = 0; grp < (1 << B); grp++) {
        // twiddle shared by the pairs whose lower index has low window bits == grp
        const uint32_t pos = ((tl | ((uint32_t)grp << p)) << s_lo) | lo;
        const Fr w = ldg_field(tw + ((uint64_t)pos << (logn - 1 - s)));
#pragma unroll
        for (int up = 0; up < (4 >> B); up++) {
            const int j = grp | (up << (B + 1));
            const int jj = j | (1 << B);
            if (DIF) {
                const Fr u = e[j] + e[jj];
                e[jj] = (e[j] - e[jj]) * w;
                e[j] = u;
            } else {
                const Fr wy = e[jj] * w;
                e[jj] = e[j] - wy;
                e[j] = e[j] + wy;
            }
        }
    }
}

template <class Fr, bool DIF, int FUSE>
__global__ void __launch_bounds__(1 << (NTT_MAX_STAGES - 3), 2)
k_ntt_pass8(const NttPassArgs<Fr> a) {
    extern __shared__ uint4 smem4[];
    Fr* __restrict__ data = a.data;
    const Fr* __restrict__ tw = a.tw;
    const int logn = a.logn, s_lo = a.s_lo, nst = a.nst;
    const int tile = 1 << nst, nthr = tile >> 3;
    const uint32_t lo = blockIdx.x & ((1u << s_lo) - 1);
    const uint32_t hi = blockIdx.x >> s_lo;
    const uint64_t base = ((uint64_t)hi << (s_lo + nst)) | lo;
    const int t = threadIdx.x;                // 2^(nst-3) threads

    // load: eight elements per thread, neighbouring threads take neighbouring elements
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int k = t + r * nthr;
        const uint64_t g = base + ((uint64_t)k << s_lo);
        Fr v;
        if (FUSE == NTT_LOAD_PAD_SCALE) {
            v = Fr::zero();
            if (g < a.src_len) v = ld_field(a.src + g) * ldg_field(a.scale_tab + g);
        } else {
            v = ld_field(data + g);
        }
        ntt_sts(smem4, tile, k, v);
    }
    __syncthreads();

    for (int done = 0; done < nst;) {
        const int r = min(3, nst - done);
        int p, b0;                           // window position, first live bit of the window
        if (DIF) { p = r == 3 ? nst - done - 3 : 0; b0 = 0; }             // stages nst-done-1 .. nst-done-r
        else     { p = r == 3 ? done : nst - 3; b0 = 3 - r; }             // stages done .. done+r-1
        const uint32_t tl = (uint32_t)t & ((1u << p) - 1);
        const int kbase = ((t >> p) << (p + 3)) | (int)tl;
        Fr e[8];
#pragma unroll
        for (int j = 0; j < 8; j++) e[j] = ntt_lds<Fr>(smem4, tile, kbase | (j << p));
        if (DIF) {
            if (b0 + r > 2) ntt_r8_stage<Fr, DIF, 2>(e, tw, tl, p, s_lo, lo, logn);
            if (b0 + r > 1) ntt_r8_stage<Fr, DIF, 1>(e, tw, tl, p, s_lo, lo, logn);
            ntt_r8_stage<Fr, DIF, 0>(e, tw, tl, p, s_lo, lo, logn);
        } else {
            if (b0 == 0) ntt_r8_stage<Fr, DIF, 0>(e, tw, tl, p, s_lo, lo, logn);
            if (b0 <= 1) ntt_r8_stage<Fr, DIF, 1>(e, tw, tl, p, s_lo, lo, logn);
            ntt_r8_stage<Fr, DIF, 2>(e, tw, tl, p, s_lo, lo, logn);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) ntt_sts(smem4, tile, kbase | (j << p), e[j]);
        __syncthreads();
        done += r;
    }

#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int k = t + r * nthr;
        const uint64_t g = base + ((uint64_t)k << s_lo);
        Fr v = ntt_lds<Fr>(smem4, tile, k);
        if (FUSE == NTT_STORE_SCALE_TAB) v = v * ldg_field(a.scale_tab + g);
        if (FUSE == NTT_STORE_SCALE) v = v * a.scale;
        st_field(data + g, v);
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

    // FIRST / LAST: fusion applied by the first / last pass of the transform (NttFuse)
    template <bool DIF, int FIRST, int LAST>
    void passes(Fr* d, const Fr* table, cudaStream_t st, const Fr* src = nullptr, uint64_t src_len = 0,
                const Fr* scale_tab = nullptr, const Fr* scale = nullptr) const {
        if (logn == 0) {
            // a single point: only the fused copy / scaling remains
            if (FIRST == NTT_LOAD_PAD_SCALE) {
                if (!src_len) {
                    B2P_CUDA(cudaMemsetAsync(d, 0, sizeof(Fr), st));
                } else {
                    if (src != d) B2P_CUDA(cudaMemcpyAsync(d, src, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
                    B2P_LAUNCH((k_mul_table<Fr>), 1, 32, 0, st, d, scale_tab, 1);
                }
            }
            if (LAST == NTT_STORE_SCALE_TAB) B2P_LAUNCH((k_mul_table<Fr>), 1, 32, 0, st, d, scale_tab, 1);
            if (LAST == NTT_STORE_SCALE) B2P_LAUNCH((k_mul_scalar<Fr>), 1, 32, 0, st, d, *scale, 1);
            return;
        }
        const int npass = (logn + NTT_MAX_STAGES - 1) / NTT_MAX_STAGES;
        // split stages as evenly as possible
        int sizes[8];
        for (int p = 0; p < npass; p++) sizes[p] = logn / npass + (p < logn % npass ? 1 : 0);
        NttPassArgs<Fr> a;
        a.data = d; a.tw = table; a.src = src; a.src_len = src_len; a.scale_tab = scale_tab;
        a.scale = scale ? *scale : Fr::zero();
        a.logn = logn;
        int s_next = DIF ? logn - 1 : 0;
        for (int p = 0; p < npass; p++) {
            a.nst = sizes[p];
            a.s_lo = DIF ? s_next - a.nst + 1 : s_next;
            s_next = DIF ? a.s_lo - 1 : a.s_lo + a.nst;
            const int fuse = (p == 0 ? FIRST : NTT_PLAIN) | (p == npass - 1 ? LAST : NTT_PLAIN);
            if (fuse == NTT_PLAIN) launch_pass<DIF, NTT_PLAIN>(a, st);
            else if (fuse == FIRST) launch_pass<DIF, FIRST>(a, st);
            else if (fuse == LAST) launch_pass<DIF, LAST>(a, st);
            else B2P_REQUIRE(false, "unsupported NTT fusion");   // a first-pass and a last-pass fusion never meet
        }
    }
    template <bool DIF, int FUSE>
    void launch_pass(const NttPassArgs<Fr>& a, cudaStream_t st) const {
        if (a.nst >= NTT_R8_MIN_STAGES && !force_radix2()) {
            auto kern = k_ntt_pass8<Fr, DIF, FUSE>;
            static std::atomic<bool> attr_set[64];     // the attribute is per device; any host thread may launch
            int dev = 0;
            B2P_CUDA(cudaGetDevice(&dev));
            if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
                B2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(sizeof(Fr) << NTT_MAX_STAGES)));
                attr_set[dev & 63].store(true, std::memory_order_release);
            }
            B2P_LAUNCH(kern, (unsigned)(n >> a.nst), 1 << (a.nst - 3), sizeof(Fr) << a.nst, st, a);
            return;
        }
        auto kern = k_ntt_pass<Fr, DIF, FUSE>;
        static std::atomic<bool> attr_set[64];
        int dev = 0;
        B2P_CUDA(cudaGetDevice(&dev));
        if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
            B2P_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(Fr) << NTT_MAX_STAGES)));
            attr_set[dev & 63].store(true, std::memory_order_release);
        }
        B2P_LAUNCH(kern, (unsigned)(n >> a.nst), 1 << (a.nst - 1), sizeof(Fr) << a.nst, st, a);
    }
    // B2P_NTT_RADIX2=1 keeps every pass on the one-butterfly-per-thread kernel (tests compare the two)
    static bool force_radix2() {
        static const bool v = [] { const char* e = getenv("B2P_NTT_RADIX2"); return e && atoi(e) != 0; }();
        return v;
    }

    void bitrev(Fr* d, cudaStream_t st) const {
        B2P_LAUNCH((k_bitrev_permute<Fr>), div_up(n, 256), 256, 0, st, d, logn);
    }
    // natural -> bit-reversed evaluations
    void forward_dif(Fr* d, cudaStream_t st) const { passes<true, NTT_PLAIN, NTT_PLAIN>(d, tw.p, st); }
    // bit-reversed evaluations -> natural coefficients (scaled by 1/n in the last pass)
    void inverse_dit(Fr* d, cudaStream_t st) const {
        passes<false, NTT_PLAIN, NTT_STORE_SCALE>(d, tw_inv.p, st, nullptr, 0, nullptr, &n_inv);
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
    // coefficients (natural) -> evaluations on g*<omega> in bit-reversed order, in place
    void coset_forward_dif(Fr* d, cudaStream_t st) const { coset_forward_dif_from(d, d, n, st); }
    // same, reading len <= n coefficients from src (zero padded) and writing the n evaluations to dst
    void coset_forward_dif_from(Fr* dst, const Fr* src, uint64_t len, cudaStream_t st) const {
        B2P_REQUIRE(has_coset, "domain built without coset tables");
        B2P_REQUIRE(len <= n, "more coefficients than domain points");
        passes<true, NTT_LOAD_PAD_SCALE, NTT_PLAIN>(dst, tw.p, st, src, len, coset_pow.p);
    }
    // evaluations on g*<omega> (bit-reversed) -> coefficients (natural); g^-i / n applied by the last pass
    void coset_inverse_dit(Fr* d, cudaStream_t st) const {
        B2P_REQUIRE(has_coset, "domain built without coset tables");
        passes<false, NTT_PLAIN, NTT_STORE_SCALE_TAB>(d, tw_inv.p, st, nullptr, 0, coset_pow_inv.p);
    }
};

#ifndef B2P_INSTANTIATE_NTT
extern template struct NttDomain<FrBn254>;
extern template struct NttDomain<FrBls12381>;
#endif

}  // namespace b2p
