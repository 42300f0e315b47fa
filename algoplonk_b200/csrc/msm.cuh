// Multi-scalar multiplication  sum_i s_i * P_i  over G1 -- the replacement for
// gnark-crypto's G1Affine.MultiExp reached through kzg.Commit (SURVEY 8a-3;
// call sites /root/reference/setup/setup.go:11,13 and gnark prove.go).
//
// B200-first design (not gnark's per-window goroutine Pippenger):
//   * The SRS is static, HBM is 180 GB: at load time every base point P_i gets
//     W-1 shifted copies 2^(c*w) * P_i (affine).  A c-bit signed digit d of
//     window w of scalar i then contributes  sign(d) * T[w][i]  to bucket |d|
//     of ONE shared bucket set -- no per-window bucket sets, no window
//     combination, and the window width can grow to c = 20 at n = 2^20.
//   * digits -> buckets is a counting sort written here: histogram with
//     global reductions, exclusive scan, scatter with fetch-add cursors.
//   * bucket accumulation: one thread per <=CAP-entry slice of a bucket, XYZZ
//     accumulator in registers, mixed additions (8M+2S), points gathered as
//     full 64 B / 96 B sectors.
//   * bucket reduction sum_b (b+1) * B_b: a radix-SEG tree of running sums.
#pragma once
#include <vector>
#include "common.cuh"
#include "scan.cuh"

namespace b2p {

constexpr int MSM_CAP = 64;        // max entries accumulated by one thread
constexpr int MSM_SEG = 8;         // radix of the bucket-reduction tree
constexpr int MSM_THREADS = 128;

struct MsmPlan {
    int c = 0;        // window bits
    int W = 0;        // windows
    uint32_t nbuckets = 0;   // 2^(c-1)
};

inline MsmPlan msm_plan(uint64_t npoints, int scalar_bits, int force_c = 0) {
    MsmPlan best;
    double best_cost = 1e300;
    for (int c = 2; c <= 22; c++) {
        if (force_c && c != force_c) continue;
        int W = (scalar_bits + 1 + c - 1) / c;
        double cost = (double)npoints * W + 3.0 * (double)(1u << (c - 1));
        if (cost < best_cost) { best_cost = cost; best.c = c; best.W = W; best.nbuckets = 1u << (c - 1); }
    }
    return best;
}

// ---------------------------------------------------------------------------
// SRS table:  T[w * npoints + i] = 2^(c*w) * P_i
// ---------------------------------------------------------------------------
template <class Fp>
__global__ void k_msm_build_table(Affine<Fp>* __restrict__ table, uint64_t npoints, int c, int W) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npoints) return;
    Affine<Fp> p;
    p.x = ld_field(&table[i].x);
    p.y = ld_field(&table[i].y);
    XYZZ<Fp> acc = XYZZ<Fp>::from_affine(p);
    for (int w = 1; w < W; w++) {
        for (int k = 0; k < c; k++) acc = acc.dbl();
        Affine<Fp> a = acc.to_affine();
        st_field(&table[(uint64_t)w * npoints + i].x, a.x);
        st_field(&table[(uint64_t)w * npoints + i].y, a.y);
        // continue from the affine form: keeps ZZ = ZZZ = 1 so later doublings stay cheap
        acc = XYZZ<Fp>::from_affine(a);
    }
}

// ---------------------------------------------------------------------------
// digit extraction
// ---------------------------------------------------------------------------
template <class Fr>
__device__ __forceinline__ uint32_t window_bits(const Fr& s, int off, int c) {
    const int limb = off >> 5, sh = off & 31;
    uint64_t lo = limb < Fr::N ? s.v[limb] : 0u;
    uint64_t hi = limb + 1 < Fr::N ? s.v[limb + 1] : 0u;
    uint64_t t = (lo | (hi << 32)) >> sh;
    return (uint32_t)(t & ((1u << c) - 1));
}

// Calls f(w, bucket, neg) for every non-zero signed digit of s (canonical form).
template <class Fr, class Fn>
__device__ __forceinline__ void for_each_digit(const Fr& s, int c, int W, Fn f) {
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    for (int w = 0; w < W; w++) {
        uint32_t d = window_bits(s, w * c, c) + carry;
        carry = 0;
        bool neg = false;
        if (d > half) { d = (1u << c) - d; neg = true; carry = 1; }
        if (d) f(w, d - 1, neg);
    }
}

template <class Fr>
__global__ void k_msm_count(const Fr* __restrict__ scalars, uint64_t n, int c, int W, int mont,
                            uint32_t* __restrict__ counts) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = ld_field(scalars + i);
    if (mont) s = s.from_mont();
    for_each_digit(s, c, W, [&](int, uint32_t b, bool) { atomicAdd(counts + b, 1u); });
}

template <class Fr>
__global__ void k_msm_scatter(const Fr* __restrict__ scalars, uint64_t n, uint64_t npoints, int c, int W, int mont,
                              const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                              uint32_t* __restrict__ entries) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = ld_field(scalars + i);
    if (mont) s = s.from_mont();
    for_each_digit(s, c, W, [&](int w, uint32_t b, bool neg) {
        uint32_t pos = atomicAdd(cursor + b, 1u);
        entries[offsets[b] + pos] = (uint32_t)((uint64_t)w * npoints + i) | (neg ? 0x80000000u : 0u);
    });
}

// ---------------------------------------------------------------------------
// bucket accumulation: thread t handles slice `sub` of bucket b
// ---------------------------------------------------------------------------
template <class Fp>
__global__ void __launch_bounds__(MSM_THREADS)
k_msm_accumulate(const Affine<Fp>* __restrict__ table, const uint32_t* __restrict__ entries,
                 const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                 const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ total_items,
                 uint32_t nbuckets, XYZZ<Fp>* __restrict__ partial) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *total_items) return;
    // bucket = last b with item_off[b] <= t
    uint32_t lo = 0, hi = nbuckets - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (item_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const uint32_t b = lo;
    const uint32_t sub = t - item_off[b];
    const uint32_t cnt = counts[b];
    const uint32_t begin = sub * MSM_CAP;
    const uint32_t len = min((uint32_t)MSM_CAP, cnt - begin);
    const uint32_t* e = entries + offsets[b] + begin;
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    for (uint32_t j = 0; j < len; j++) {
        const uint32_t ent = e[j];
        const Affine<Fp>* src = table + (ent & 0x7fffffffu);
        Affine<Fp> p;
        p.x = ldg_field(&src->x);
        p.y = ldg_field(&src->y);
        acc.add_affine_signed(p, ent >> 31);
    }
    XYZZ<Fp>* dst = partial + t;
    st_field(&dst->X, acc.X);
    st_field(&dst->Y, acc.Y);
    st_field(&dst->ZZ, acc.ZZ);
    st_field(&dst->ZZZ, acc.ZZZ);
}

template <class Fp>
__device__ __forceinline__ XYZZ<Fp> ld_xyzz(const XYZZ<Fp>* p) {
    XYZZ<Fp> r;
    r.X = ld_field(&p->X); r.Y = ld_field(&p->Y); r.ZZ = ld_field(&p->ZZ); r.ZZZ = ld_field(&p->ZZZ);
    return r;
}
template <class Fp>
__device__ __forceinline__ void st_xyzz(XYZZ<Fp>* p, const XYZZ<Fp>& r) {
    st_field(&p->X, r.X); st_field(&p->Y, r.Y); st_field(&p->ZZ, r.ZZ); st_field(&p->ZZZ, r.ZZZ);
}

// ---------------------------------------------------------------------------
// bucket reduction.  Invariant over levels (m entries, index i has weight i):
//     result = sum_i A_i + scale * sum_i i * P_i + sum_i P_i
// One step groups SEG consecutive entries:  A'_s = sum_j A_j + scale * sum_j j * P_j,
// P'_s = sum_j P_j, scale' = scale * SEG.  At m == 1: result = A_0 + P_0.
// Level 0 reads the per-item partial sums (A absent).
// ---------------------------------------------------------------------------
template <class Fp, bool FIRST>
__global__ void __launch_bounds__(MSM_THREADS)
k_msm_reduce_level(const XYZZ<Fp>* __restrict__ Pin, const XYZZ<Fp>* __restrict__ Ain,
                   const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ total_items,
                   uint32_t m, int log_scale, XYZZ<Fp>* __restrict__ Pout, XYZZ<Fp>* __restrict__ Aout) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t mout = (m + MSM_SEG - 1) / MSM_SEG;
    if (s >= mout) return;
    XYZZ<Fp> running = XYZZ<Fp>::inf();   // sum of P_j seen so far (descending j)
    XYZZ<Fp> weighted = XYZZ<Fp>::inf();  // sum_j j * P_j
    XYZZ<Fp> asum = XYZZ<Fp>::inf();
    for (int j = MSM_SEG - 1; j >= 0; j--) {
        const uint32_t i = s * MSM_SEG + j;
        if (i < m) {
            XYZZ<Fp> p;
            if (FIRST) {
                // bucket sum = sum of its item partials
                const uint32_t b0 = item_off[i];
                const uint32_t b1 = (i + 1 < m) ? item_off[i + 1] : *total_items;
                p = XYZZ<Fp>::inf();
                for (uint32_t t = b0; t < b1; t++) p.add(ld_xyzz(Pin + t));
            } else {
                p = ld_xyzz(Pin + i);
                asum.add(ld_xyzz(Ain + i));
            }
            running.add(p);
        }
        if (j > 0) weighted.add(running);
    }
    for (int k = 0; k < log_scale; k++) weighted = weighted.dbl();
    asum.add(weighted);
    st_xyzz(Pout + s, running);
    st_xyzz(Aout + s, asum);
}

template <class Fp>
__global__ void k_msm_final(const XYZZ<Fp>* __restrict__ P, const XYZZ<Fp>* __restrict__ A, XYZZ<Fp>* __restrict__ out,
                            const uint32_t* __restrict__ total_entries, unsigned long long* __restrict__ adds_total) {
    XYZZ<Fp> r = ld_xyzz(A);
    r.add(ld_xyzz(P));
    st_xyzz(out, r);
    *adds_total += *total_entries;   // single thread, stream ordered: mixed additions done by the accumulation
}

// ---------------------------------------------------------------------------
// host-side driver
// ---------------------------------------------------------------------------
template <class C>
struct MsmEngine {
    using Fr = typename C::Fr;
    using Fp = typename C::Fp;
    using Aff = Affine<Fp>;
    using Ext = XYZZ<Fp>;

    uint64_t npoints = 0;
    MsmPlan plan;
    DevBuf<Aff> table;

    // scratch (sized for npoints scalars)
    DevBuf<uint32_t> counts, offsets, cursor, item_off, entries, scan_scratch, total_items;
    DevBuf<Ext> partial, lvlP[2], lvlA[2], result;
    uint32_t max_items = 0;
    Profiler* prof = nullptr;
    DevBuf<uint32_t> total_entries;
    DevBuf<unsigned long long> adds_total;   // running count of mixed additions (non-zero digits), device side

    void load(const void* host_points, uint64_t n, int force_c, cudaStream_t st) {
        npoints = n;
        plan = msm_plan(n, Fr::Params::BITS, force_c);
        B2P_REQUIRE((uint64_t)plan.W * n < (1ull << 31), "SRS too large for 31-bit table indices");
        table.alloc((size_t)plan.W * n);
        B2P_CUDA(cudaMemcpyAsync(table.p, host_points, n * sizeof(Aff), cudaMemcpyHostToDevice, st));
        B2P_LAUNCH((k_msm_build_table<Fp>), div_up(n, 128), 128, 0, st, table.p, n, plan.c, plan.W);
        alloc_scratch();
    }
    // points already on the device (first npoints entries of a W*npoints table buffer)
    void load_device(DevBuf<Aff>&& tbl, uint64_t n, const MsmPlan& pl, cudaStream_t st) {
        npoints = n;
        plan = pl;
        table = std::move(tbl);
        B2P_LAUNCH((k_msm_build_table<Fp>), div_up(n, 128), 128, 0, st, table.p, n, plan.c, plan.W);
        alloc_scratch();
    }
    void alloc_scratch() {
        const uint32_t nb = plan.nbuckets;
        counts.alloc(nb); offsets.alloc(nb); cursor.alloc(nb); item_off.alloc(nb);
        entries.alloc((size_t)plan.W * npoints);
        scan_scratch.alloc(scan_scratch_words(nb));
        total_items.alloc(1);
        total_entries.alloc(1);
        adds_total.alloc(1);
        B2P_CUDA(cudaMemset(adds_total.p, 0, sizeof(unsigned long long)));
        max_items = nb + (uint32_t)(((uint64_t)plan.W * npoints) / MSM_CAP) + 1;
        partial.alloc(max_items);
        const uint32_t m1 = div_up(nb, MSM_SEG);
        for (int k = 0; k < 2; k++) { lvlP[k].alloc(m1); lvlA[k].alloc(m1); }
        result.alloc(1);
    }

    // d_scalars: device pointer, n <= npoints scalars; result (XYZZ, device) in this->result.
    void run_async(const Fr* d_scalars, uint64_t n, bool mont, cudaStream_t st) {
        B2P_REQUIRE(n <= npoints, "MSM: more scalars than SRS points");
        const uint32_t nb = plan.nbuckets;
        B2P_CUDA(cudaMemsetAsync(counts.p, 0, nb * sizeof(uint32_t), st));
        B2P_CUDA(cudaMemsetAsync(cursor.p, 0, nb * sizeof(uint32_t), st));
        if (n) B2P_LAUNCH((k_msm_count<Fr>), div_up(n, 256), 256, 0, st, d_scalars, n, plan.c, plan.W, (int)mont, counts.p);
        exclusive_scan_u32(counts.p, offsets.p, nb, scan_scratch.p, total_entries.p, st, ScanIdentity{});
        exclusive_scan_u32(counts.p, item_off.p, nb, scan_scratch.p, total_items.p, st, ScanCeilDiv{MSM_CAP});
        if (n) B2P_LAUNCH((k_msm_scatter<Fr>), div_up(n, 256), 256, 0, st, d_scalars, n, npoints, plan.c, plan.W, (int)mont,
                          offsets.p, cursor.p, entries.p);
        const int span = prof ? prof->begin(B2P_STAT_MSM_ACCUM_MS, st) : -1;
        B2P_LAUNCH((k_msm_accumulate<Fp>), div_up(max_items, MSM_THREADS), MSM_THREADS, 0, st, table.p, entries.p,
                   counts.p, offsets.p, item_off.p, total_items.p, nb, partial.p);
        if (prof) prof->end(span, st);
        // reduction tree
        uint32_t m = nb;
        int level = 0;
        int log_seg = 0;
        while ((1 << log_seg) < MSM_SEG) log_seg++;
        const Ext* Pin = partial.p;
        const Ext* Ain = nullptr;
        do {
            const uint32_t mout = div_up(m, MSM_SEG);
            Ext* Pout = lvlP[level & 1].p;
            Ext* Aout = lvlA[level & 1].p;
            if (level == 0)
                B2P_LAUNCH((k_msm_reduce_level<Fp, true>), div_up(mout, MSM_THREADS), MSM_THREADS, 0, st, Pin, Ain,
                           item_off.p, total_items.p, m, 0, Pout, Aout);
            else
                B2P_LAUNCH((k_msm_reduce_level<Fp, false>), div_up(mout, MSM_THREADS), MSM_THREADS, 0, st, Pin, Ain,
                           item_off.p, total_items.p, m, level * log_seg, Pout, Aout);
            Pin = Pout; Ain = Aout; m = mout; level++;
        } while (m > 1);
        B2P_LAUNCH((k_msm_final<Fp>), 1, 1, 0, st, Pin, Ain, result.p, total_entries.p, adds_total.p);
    }

    // synchronous convenience: returns the affine result (host)
    Aff run(const Fr* d_scalars, uint64_t n, bool mont, cudaStream_t st) {
        run_async(d_scalars, n, mont, st);
        Ext h;
        B2P_CUDA(cudaMemcpyAsync(&h, result.p, sizeof(Ext), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        return h.to_affine();
    }
};

#ifndef B2P_INSTANTIATE_MSM
extern template struct MsmEngine<Bn254>;
extern template struct MsmEngine<Bls12381>;
#endif

}  // namespace b2p
