// Coefficient-wise / evaluation-wise polynomial kernels of the PLONK prover:
// grand product (iop.BuildRatioCopyConstraint), quotient numerator
// (computeNumerator + divideByZH), evaluation and division by (X - z)
// (kzg.Open / BatchOpenSinglePoint), linearised polynomial and fold
// (innerComputeLinearizedPoly) -- SURVEY 8a-5..8a-7, Appendix A.
#pragma once
#include "common.cuh"
#include "ntt.cuh"

namespace b2p {

// ---------------------------------------------------------------------------
// generic inclusive scan over field elements (Op = multiply or add)
// ---------------------------------------------------------------------------
struct OpMul {
    template <class F> __device__ static F id() { return F::one(); }
    template <class F> __device__ static F apply(const F& a, const F& b) { return a * b; }
};
struct OpAdd {
    template <class F> __device__ static F id() { return F::zero(); }
    template <class F> __device__ static F apply(const F& a, const F& b) { return a + b; }
};

constexpr int FSCAN_THREADS = 256;
constexpr int FSCAN_ITEMS = 8;
constexpr int FSCAN_TILE = FSCAN_THREADS * FSCAN_ITEMS;

// phase 1: in-tile inclusive scan, tile totals out
template <class F, class Op>
__global__ void __launch_bounds__(FSCAN_THREADS)
k_fscan_tiles(F* __restrict__ data, F* __restrict__ totals, uint64_t n) {
    __shared__ uint4 sm[F::N / 4][FSCAN_THREADS];
    const uint64_t base = (uint64_t)blockIdx.x * FSCAN_TILE + (uint64_t)threadIdx.x * FSCAN_ITEMS;
    F acc = Op::template id<F>();
    F v[FSCAN_ITEMS];
#pragma unroll
    for (int i = 0; i < FSCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? ld_field(data + base + i) : Op::template id<F>();
        acc = Op::apply(acc, v[i]);
        v[i] = acc;
    }
    // Hillis-Steele over the per-thread totals
    auto put = [&](const F& x) {
#pragma unroll
        for (int q = 0; q < F::N / 4; q++)
            sm[q][threadIdx.x] = make_uint4(x.v[4 * q], x.v[4 * q + 1], x.v[4 * q + 2], x.v[4 * q + 3]);
    };
    auto get = [&](int t) {
        F x;
#pragma unroll
        for (int q = 0; q < F::N / 4; q++) {
            uint4 a = sm[q][t];
            x.v[4 * q] = a.x; x.v[4 * q + 1] = a.y; x.v[4 * q + 2] = a.z; x.v[4 * q + 3] = a.w;
        }
        return x;
    };
    F incl = acc;
    put(incl);
    __syncthreads();
    for (int d = 1; d < FSCAN_THREADS; d <<= 1) {
        F other;
        const bool take = (int)threadIdx.x >= d;
        if (take) other = get(threadIdx.x - d);
        __syncthreads();
        if (take) { incl = Op::apply(other, incl); put(incl); }
        __syncthreads();
    }
    // exclusive prefix of this thread
    F excl = threadIdx.x ? get(threadIdx.x - 1) : Op::template id<F>();
#pragma unroll
    for (int i = 0; i < FSCAN_ITEMS; i++)
        if (base + i < n) st_field(data + base + i, threadIdx.x ? Op::apply(excl, v[i]) : v[i]);
    if (threadIdx.x == FSCAN_THREADS - 1) st_field(totals + blockIdx.x, incl);
}

// phase 3: combine tile b with the inclusive scan of the totals of tiles < b
template <class F, class Op>
__global__ void __launch_bounds__(FSCAN_THREADS)
k_fscan_apply(F* __restrict__ data, const F* __restrict__ totals_scanned, uint64_t n) {
    if (blockIdx.x == 0) return;
    const F pre = ld_field(totals_scanned + blockIdx.x - 1);
    const uint64_t base = (uint64_t)blockIdx.x * FSCAN_TILE + (uint64_t)threadIdx.x * FSCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < FSCAN_ITEMS; i++)
        if (base + i < n) st_field(data + base + i, Op::apply(pre, ld_field(data + base + i)));
}

inline size_t fscan_scratch_elems(uint64_t n) {
    size_t e = 0;
    while (n > 1) {
        n = div_up(n, FSCAN_TILE);
        e += n;
        if (n == 1) break;
    }
    return e + 1;
}

template <class F, class Op>
void field_scan_inclusive(F* data, uint64_t n, F* scratch, cudaStream_t st) {
    if (n == 0) return;
    const unsigned tiles = div_up(n, FSCAN_TILE);
    B2P_LAUNCH((k_fscan_tiles<F, Op>), tiles, FSCAN_THREADS, 0, st, data, scratch, n);
    if (tiles == 1) return;
    field_scan_inclusive<F, Op>(scratch, tiles, scratch + tiles, st);
    B2P_LAUNCH((k_fscan_apply<F, Op>), tiles, FSCAN_THREADS, 0, st, data, scratch, n);
}

// ---------------------------------------------------------------------------
// batch inversion, in place: each thread inverts BINV_CHUNK elements with one
// field inversion (Montgomery's trick).  Zero elements stay zero.
// ---------------------------------------------------------------------------
constexpr int BINV_CHUNK = 16;
template <class F>
__global__ void k_batch_inverse(F* __restrict__ data, uint64_t n) {
    const uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * BINV_CHUNK;
    if (base >= n) return;
    F pre[BINV_CHUNK];
    F acc = F::one();
    const int cnt = (int)min((uint64_t)BINV_CHUNK, n - base);
    for (int i = 0; i < cnt; i++) {
        pre[i] = acc;
        F x = ld_field(data + base + i);
        if (!x.is_zero()) acc = acc * x;
    }
    F inv = acc.inverse();
    for (int i = cnt - 1; i >= 0; i--) {
        F x = ld_field(data + base + i);
        if (x.is_zero()) continue;
        st_field(data + base + i, inv * pre[i]);
        inv = inv * x;
    }
}

// ---------------------------------------------------------------------------
// small element-wise helpers
// ---------------------------------------------------------------------------
template <class F>
__global__ void k_mul_pointwise(F* __restrict__ out, const F* __restrict__ a, const F* __restrict__ b, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_field(out + i, ld_field(a + i) * ld_field(b + i));
}

// p(X) += b(X) * (X^n - 1) for a blinding polynomial of `nb` coefficients (gnark getBlindedCoefficients);
// p has room for n + nb coefficients and p[n..] is zero on entry.
template <class F>
__global__ void k_blind(F* __restrict__ p, uint64_t n, int nb, const F* __restrict__ b) {
    int i = threadIdx.x;
    if (i >= nb) return;
    F bi = ld_field(b + i);
    st_field(p + i, ld_field(p + i) - bi);
    st_field(p + n + i, ld_field(p + n + i) + bi);
}

// writes values[j] into dst[idx[j]]
template <class F>
__global__ void k_scatter_small(F* __restrict__ dst, const uint32_t* __restrict__ idx, const F* __restrict__ values, int cnt) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cnt) st_field(dst + idx[j], ld_field(values + j));
}

// omega^i for i < n from the half-size twiddle table
template <class F>
__device__ __forceinline__ F omega_pow(const F* __restrict__ tw, uint64_t i, uint64_t n) {
    const uint64_t half = n >> 1;
    if (n == 1) return F::one();
    if (i < half) return ldg_field(tw + i);
    return ldg_field(tw + (i - half)).neg();
}

// S_j in Lagrange form: s[t] = id[perm[t]], id = [w^i, u w^i, u^2 w^i]  (gnark computePermutationPolynomials)
template <class F>
__global__ void k_perm_to_lagrange(F* __restrict__ s, const int64_t* __restrict__ perm, const F* __restrict__ tw,
                                   uint64_t n, F u, F u2) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const uint64_t q = (uint64_t)perm[t];
    const uint64_t col = q / n, row = q % n;
    F v = omega_pow(tw, row, n);
    if (col == 1) v = v * u;
    else if (col == 2) v = v * u2;
    st_field(s + t, v);
}

// ---------------------------------------------------------------------------
// grand product terms:  f[i+1] = num_i (i < n-1), f[0] = 1;  the denominators are stored REVERSED,
// gr[n-2-i] = den_i, gr[n-1] = 1, so that the prefix scan of gr yields the suffix products of den:
//   Z_k = prod_{i<k} num_i / den_i = F_k * (prod_{i>=k} den_i) / (prod_all den_i) = F_k * PR[n-2-k] * Ginv
// with F, PR the inclusive scans of f, gr and Ginv = 1 / PR[n-1]: ONE field inversion (on the host)
// instead of a batch inversion of n prefix products.
// ---------------------------------------------------------------------------
template <class F>
__global__ void k_z_terms(F* __restrict__ f, F* __restrict__ gr, const F* __restrict__ L, const F* __restrict__ R,
                          const F* __restrict__ O, const F* __restrict__ S /* 3n Lagrange */, const F* __restrict__ tw,
                          uint64_t n, F beta, F gamma, F u, F u2) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) {
        st_field(f, F::one());
        st_field(gr + (n - 1), F::one());
        return;
    }
    const F l = ld_field(L + i) + gamma, r = ld_field(R + i) + gamma, o = ld_field(O + i) + gamma;
    const F bw = beta * omega_pow(tw, i, n);
    constexpr uint32_t US = F::Params::SHIFT_SMALL;     // u and u^2 are small integers: additions, not products
    const F num = (l + bw) * (r + bw.template mul_small<US>()) * (o + bw.template mul_small<US * US>());
    const F den = (l + beta * ld_field(S + i)) * (r + beta * ld_field(S + n + i)) * (o + beta * ld_field(S + 2 * n + i));
    st_field(f + i + 1, num);
    st_field(gr + (n - 2 - i), den);
}
// Z_k = F_k * PR[n-2-k] * ginv  (k = n-1: the empty suffix product)
template <class F>
__global__ void k_z_finish(F* __restrict__ z, const F* __restrict__ Fs, const F* __restrict__ PR, uint64_t n, F ginv) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    F v = ld_field(Fs + k) * ginv;
    if (k + 1 < n) v = v * ld_field(PR + (n - 2 - k));
    st_field(z + k, v);
}

// ---------------------------------------------------------------------------
// quotient on the big coset, bit-reversed layout (position p <-> natural index brev(p))
// ---------------------------------------------------------------------------
constexpr int MAX_QCP = 8;
constexpr int MAX_PI_DIRECT = 8;
template <class F>
struct QuotientArgs {
    const F *l, *r, *o, *z;               // blinded wire / grand product evaluations
    const F *ql, *qr, *qm, *qo, *qk;      // selector evaluations (qk completed with public inputs)
    const F *s1, *s2, *s3;
    const F *x;                           // coset points g * w^i
    const F *l1;                          // L_1 evaluations
    const F *qcp[MAX_QCP];
    const F *pi2[MAX_QCP];
    int k;
    F *h;                                 // out
    int logm, log_rho;
    F beta, gamma, alpha, alpha2, u, u2;
    F zh_inv[8];                          // 1 / (X^n - 1) by natural index mod rho
    // Public inputs and BSB22 commitment hashes sit in rows of qk (gnark completeQk).  With only a few of
    // them, qk is NOT completed and transformed per proof: a.qk holds the key's qk and the missing part
    //   sum_t pi_val[t] * L_{pi_row[t]}(x),   L_i(x) = L_0(x * omega^-i),
    // is read off the resident L_0 evaluations (a rotation by rho * i in natural order).
    int n_pi;
    uint32_t pi_row[MAX_PI_DIRECT];
    F pi_val[MAX_PI_DIRECT];
};

template <class F>
__global__ void __launch_bounds__(256) k_quotient(const QuotientArgs<F> a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = 1u << a.logm;
    if (p >= m) return;
    const uint32_t rho = 1u << a.log_rho;
    const uint32_t nat = brev32(p, a.logm);
    const uint32_t pshift = brev32((nat + rho) & (m - 1), a.logm);

    const F l = ld_field(a.l + p), r = ld_field(a.r + p), o = ld_field(a.o + p);
    F gate = (ld_field(a.ql + p) + ld_field(a.qm + p) * r) * l + ld_field(a.qr + p) * r
           + ld_field(a.qo + p) * o + ld_field(a.qk + p);
    for (int c = 0; c < a.k; c++) gate = gate + ld_field(a.qcp[c] + p) * ld_field(a.pi2[c] + p);
    for (int t = 0; t < a.n_pi; t++) {
        const uint32_t q = brev32((nat - rho * a.pi_row[t]) & (m - 1), a.logm);
        gate = gate + a.pi_val[t] * ld_field(a.l1 + q);
    }

    const F z = ld_field(a.z + p), zs = ld_field(a.z + pshift);
    const F lg = l + a.gamma, rg = r + a.gamma, og = o + a.gamma;
    const F pa = (lg + a.beta * ld_field(a.s1 + p)) * (rg + a.beta * ld_field(a.s2 + p))
               * (og + a.beta * ld_field(a.s3 + p)) * zs;
    const F bx = a.beta * ld_field(a.x + p);
    constexpr uint32_t US = F::Params::SHIFT_SMALL;     // u = 5 / 7: multiples by additions (ALU pipe)
    const F pb = (lg + bx) * (rg + bx.template mul_small<US>()) * (og + bx.template mul_small<US * US>()) * z;
    const F loc = ld_field(a.l1 + p) * (z - F::one());
    const F num = gate + a.alpha * (pa - pb) + a.alpha2 * loc;
    st_field(a.h + p, num * a.zh_inv[nat & (rho - 1)]);
}

// L_1 on the coset: (X^n - 1) / (n (X - 1)); input x (bit-reversed layout), out = denominators n (X - 1)
// (batch-inverted afterwards, then multiplied by X^n - 1 via k_l1_finish).
template <class F>
__global__ void k_l1_denoms(F* __restrict__ out, const F* __restrict__ x, uint64_t m, F nfr) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    st_field(out + p, (ld_field(x + p) - F::one()) * nfr);
}
template <class F>
struct ZhVals { F v[8]; };
template <class F>
__global__ void k_l1_finish(F* __restrict__ out, uint64_t m, int logm, int log_rho, const ZhVals<F> zh) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    const uint32_t nat = brev32((uint32_t)p, logm);
    st_field(out + p, ld_field(out + p) * zh.v[nat & ((1u << log_rho) - 1)]);
}
// x[p] = g * w^brev(p)
template <class F>
__global__ void k_coset_points(F* __restrict__ x, const F* __restrict__ tw, int logm, F g) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t m = 1ull << logm;
    if (p >= m) return;
    st_field(x + p, omega_pow(tw, brev32((uint32_t)p, logm), m) * g);
}

// ---------------------------------------------------------------------------
// evaluation: dot products of up to MAX_DOT polynomials with one power table
// ---------------------------------------------------------------------------
constexpr int MAX_DOT = 16;
constexpr int DOT_THREADS = 256;
constexpr int DOT_ITEMS = 8;
template <class F>
struct DotArgs {
    const F* poly[MAX_DOT];
    uint64_t len[MAX_DOT];
    int npoly;
    const F* pow;      // z^j
    F* partial;        // [npoly][gridDim.x]
};

template <class F>
__device__ __forceinline__ F block_sum(F v) {
    __shared__ uint32_t sm[F::N][DOT_THREADS / 32];
    // warp tree via shuffles, limb by limb
    for (int d = 16; d >= 1; d >>= 1) {
        F o;
#pragma unroll
        for (int i = 0; i < F::N; i++) o.v[i] = __shfl_down_sync(0xffffffffu, v.v[i], d);
        v = v + o;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < F::N; i++) sm[i][wid] = v.v[i];
    }
    __syncthreads();
    F tot = F::zero();
    if (threadIdx.x == 0) {
        for (int w = 0; w < DOT_THREADS / 32; w++) {
            F o;
#pragma unroll
            for (int i = 0; i < F::N; i++) o.v[i] = sm[i][w];
            tot = tot + o;
        }
    }
    return tot;   // valid on thread 0
}

template <class F>
__global__ void __launch_bounds__(DOT_THREADS) k_dot_pow(const DotArgs<F> a) {
    const int pi = blockIdx.y;
    const uint64_t len = a.len[pi];
    const F* poly = a.poly[pi];
    F acc = F::zero();
    for (uint64_t j = (uint64_t)blockIdx.x * DOT_THREADS + threadIdx.x; j < len; j += (uint64_t)gridDim.x * DOT_THREADS)
        acc = acc + ld_field(poly + j) * ldg_field(a.pow + j);
    F tot = block_sum(acc);
    if (threadIdx.x == 0) st_field(a.partial + (uint64_t)pi * gridDim.x + blockIdx.x, tot);
}
template <class F>
__global__ void __launch_bounds__(DOT_THREADS) k_dot_finish(const F* __restrict__ partial, int nblocks, F* __restrict__ out) {
    const int pi = blockIdx.x;
    F acc = F::zero();
    for (int j = threadIdx.x; j < nblocks; j += DOT_THREADS) acc = acc + ld_field(partial + (uint64_t)pi * nblocks + j);
    F tot = block_sum(acc);
    if (threadIdx.x == 0) st_field(out + pi, tot);
}

// ---------------------------------------------------------------------------
// division by (X - z):  q_{i-1} = (p(z) - sum_{j<i} p_j z^j) * z^-i , i = 1..len-1
//   step 1: t_j = p_j z^j          (k_mul_pointwise with the power table)
//   step 2: inclusive additive scan of t
//   step 3: q_{i-1} = (T[len-1] - T[i-1]) * zinv^i
// ---------------------------------------------------------------------------
template <class F>
__global__ void k_div_finish(F* __restrict__ q, const F* __restrict__ T, const F* __restrict__ pow_inv, uint64_t len) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (i >= len) return;
    const F total = ld_field(T + len - 1);
    st_field(q + i - 1, (total - ld_field(T + i - 1)) * ldg_field(pow_inv + i));
}

// ---------------------------------------------------------------------------
// out[i] = sum_t coef[t] * poly[t][i]  over polynomials of different lengths
// (linearised polynomial, opening fold)
// ---------------------------------------------------------------------------
constexpr int MAX_LC = 20;
template <class F>
struct LinCombArgs {
    const F* poly[MAX_LC];
    uint64_t len[MAX_LC];
    F coef[MAX_LC];
    int unit[MAX_LC];      // coefficient is exactly 1: skip the multiplication
    int nterms;
    F* out;
    uint64_t out_len;
};
template <class F>
__global__ void __launch_bounds__(256) k_lincomb(const LinCombArgs<F> a) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.out_len) return;
    F acc = F::zero();
    for (int t = 0; t < a.nterms; t++) {
        if (i < a.len[t]) {
            F v = ld_field(a.poly[t] + i);
            acc = acc + (a.unit[t] ? v : v * a.coef[t]);
        }
    }
    st_field(a.out + i, acc);
}

}  // namespace b2p
