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
//   * digits -> buckets is a counting sort written here: histogram whose
//     fetch-adds also rank every digit inside its bucket, exclusive scan, scatter.
//   * bucket accumulation: one thread per <=CAP-entry slice of a bucket (slices sorted
//     by length so a warp's lanes run equally long), XYZZ accumulator in registers,
//     mixed additions (8M+2S), points gathered as full 64 B / 96 B sectors one
//     addition ahead; bucket sums land in a dense array.
//   * bucket reduction sum_b (b+1) * B_b: bucket index split b = h * 2^s + l, plain
//     row / column sums (level 1: one wave of long per-thread chains; level 2: a
//     shared-memory tree), then a bit-decomposed weighted sum of the ~1.5 k row and
//     column sums.  Levels 2 and 3 are latency-bound and therefore deferred: they run
//     once per fetch for all MSMs queued since (result slots).
#pragma once
#include <vector>
#include "common.cuh"
#include "scan.cuh"
#include "msm_digits.cuh"

namespace b2p {

constexpr int MSM_CAP = 128;       // max entries accumulated by one thread (the partial top window puts
                                   // ~n / 2^(bits - (W-1)c) entries into each of its buckets: ~90 at n = 2^20, c = 20)
constexpr int MSM_THREADS = 128;
// resident accumulation blocks per SM: 4 x 128 registers for 8-limb base fields (BN254); a 12-limb field
// (BLS12-381) spills at 128 or 168 registers and loses more than the extra warps give (measured: 46.5 / 45.0 /
// 44.2 ms of accumulation per 2^20 proof at 4 / 3 / 2 blocks)
constexpr int MSM_ACC_BLOCKS_WIDE = 2;
constexpr int MSM_SLOTS = 16;      // MSMs that can be queued before their results are fetched

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

// Histogram pass.  The fetch-add that counts a digit also hands out its rank inside the bucket; the rank
// is kept (ranks[w * n + i], coalesced) so that the scatter pass needs no second round of atomics.
template <class Fr>
__global__ void k_msm_count(const Fr* __restrict__ scalars, uint64_t n, int c, int W, int mont,
                            uint32_t* __restrict__ counts, uint32_t* __restrict__ ranks) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = ld_field(scalars + i);
    if (mont) s = s.from_mont();
    for_each_digit(s, c, W, [&](int w, uint32_t b, bool) { ranks[(uint64_t)w * n + i] = atomicAdd(counts + b, 1u); });
}

template <class Fr>
__global__ void k_msm_scatter(const Fr* __restrict__ scalars, uint64_t n, uint64_t npoints, int c, int W, int mont,
                              const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ranks,
                              uint32_t* __restrict__ entries) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = ld_field(scalars + i);
    if (mont) s = s.from_mont();
    for_each_digit(s, c, W, [&](int w, uint32_t b, bool neg) {
        entries[offsets[b] + ranks[(uint64_t)w * n + i]] = (uint32_t)((uint64_t)w * npoints + i) | (neg ? 0x80000000u : 0u);
    });
}

// ---------------------------------------------------------------------------
// Work items.  A bucket with cnt entries is cut into ceil(cnt / CAP) items of at
// most CAP entries; one thread accumulates one item.  Bucket sizes are Poisson
// distributed, so handing a warp 32 neighbouring buckets leaves ~30 % of its lanes
// idle (measured: 22 of 32 lanes active).  Items are therefore counting-sorted by
// length, longest first: lanes of a warp get equal trip counts and the long items
// do not end up in the tail of the launch.
//   hist[(CAP - len) * nblk + blk]  -> exclusive scan ->  start of (len, blk) in `order`
// A bucket with a single item (the rule: Poisson(26) against CAP = 64) is written
// straight into the dense bucket array; buckets with more items (skewed scalars: many
// 0/1/small witness values land in one bucket) are listed and their item sums are
// added by one warp each (k_msm_multi_buckets).
// ---------------------------------------------------------------------------
constexpr int MSM_SORT_THREADS = 256;

static __global__ void __launch_bounds__(MSM_SORT_THREADS)
k_msm_len_hist(const uint32_t* __restrict__ counts, uint32_t nb, uint32_t nblk, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[MSM_CAP + 1];
    for (int i = threadIdx.x; i <= MSM_CAP; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) {
        const uint32_t cnt = counts[b], nfull = cnt / MSM_CAP, rem = cnt % MSM_CAP;
        if (nfull) atomicAdd(&sh[MSM_CAP], nfull);
        if (rem) atomicAdd(&sh[rem], 1u);
    }
    __syncthreads();
    for (int len = 1 + threadIdx.x; len <= MSM_CAP; len += blockDim.x)
        hist[(uint32_t)(MSM_CAP - len) * nblk + blockIdx.x] = sh[len];
}

static __global__ void __launch_bounds__(MSM_SORT_THREADS)
k_msm_len_scatter(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ item_off, uint32_t nb,
                  uint32_t nblk, const uint32_t* __restrict__ start, uint2* __restrict__ order,
                  uint32_t* __restrict__ big_list, uint32_t* __restrict__ big_count) {
    __shared__ uint32_t sh[MSM_CAP + 1];
    for (int len = 1 + threadIdx.x; len <= MSM_CAP; len += blockDim.x)
        sh[len] = start[(uint32_t)(MSM_CAP - len) * nblk + blockIdx.x];
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t cnt = counts[b], nfull = cnt / MSM_CAP, rem = cnt % MSM_CAP;
    if (nfull) {
        const uint32_t p = atomicAdd(&sh[MSM_CAP], nfull);
        for (uint32_t k = 0; k < nfull; k++) order[p + k] = make_uint2(b, k);
    }
    if (rem) order[atomicAdd(&sh[rem], 1u)] = make_uint2(b, nfull);
    if (nfull + (rem ? 1u : 0u) > 1u) big_list[atomicAdd(big_count, 1u)] = b;   // more than one item
    (void)item_off;
}

// ---------------------------------------------------------------------------
// bucket accumulation: thread t handles item order[t] = (bucket, slice)
// ---------------------------------------------------------------------------
template <class Fp> struct XYZZ;
template <class Fp> __device__ __forceinline__ void st_xyzz_fwd(XYZZ<Fp>* p, const XYZZ<Fp>& r) {
    st_field(&p->X, r.X); st_field(&p->Y, r.Y); st_field(&p->ZZ, r.ZZ); st_field(&p->ZZZ, r.ZZZ);
}
template <class Fp>
__device__ __forceinline__ Affine<Fp> ldg_point(const Affine<Fp>* src) {
    Affine<Fp> p;
    p.x = ldg_field(&src->x);
    p.y = ldg_field(&src->y);
    return p;
}

template <class Fp>
__global__ void __launch_bounds__(MSM_THREADS, (Fp::N <= 8 ? 4 : MSM_ACC_BLOCKS_WIDE))
k_msm_accumulate(const Affine<Fp>* __restrict__ table, const uint32_t* __restrict__ entries,
                 const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                 const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ total_items,
                 const uint2* __restrict__ order, XYZZ<Fp>* __restrict__ partial, XYZZ<Fp>* __restrict__ buckets) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *total_items) return;
    const uint2 it = order[t];
    const uint32_t b = it.x, sub = it.y;
    const uint32_t cnt = counts[b];
    const uint32_t begin = sub * MSM_CAP;
    const uint32_t len = min((uint32_t)MSM_CAP, cnt - begin);
    const uint32_t* e = entries + offsets[b] + begin;
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    // software pipeline: the gather of point j+1 is in flight while point j is added
    uint32_t ent = e[0];
    Affine<Fp> p = ldg_point(table + (ent & 0x7fffffffu));
#pragma unroll 1
    for (uint32_t j = 0; j < len; j++) {
        const uint32_t ent_cur = ent;
        const Affine<Fp> cur = p;
        if (j + 1 < len) {
            ent = e[j + 1];
            p = ldg_point(table + (ent & 0x7fffffffu));
        }
        acc.add_affine_signed(cur, ent_cur >> 31);
    }
    st_xyzz_fwd(cnt <= MSM_CAP ? buckets + b : partial + item_off[b] + sub, acc);
}

template <class Fp>
__device__ __forceinline__ XYZZ<Fp> ld_xyzz(const XYZZ<Fp>* p) {
    XYZZ<Fp> r;
    r.X = ld_field(&p->X); r.Y = ld_field(&p->Y); r.ZZ = ld_field(&p->ZZ); r.ZZZ = ld_field(&p->ZZZ);
    return r;
}
template <class Fp>
__device__ __forceinline__ XYZZ<Fp> ld_xyzz_cg(const XYZZ<Fp>* p) {   // L2-coherent loads
    XYZZ<Fp> r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<Fp>) / 16); i++) {
        const uint4 t = __ldcg(q + i);
        w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
    }
    return r;
}
template <class Fp>
__device__ __forceinline__ void st_xyzz(XYZZ<Fp>* p, const XYZZ<Fp>& r) {
    st_field(&p->X, r.X); st_field(&p->Y, r.Y); st_field(&p->ZZ, r.ZZ); st_field(&p->ZZZ, r.ZZZ);
}

// ---------------------------------------------------------------------------
// warp-level sums of XYZZ points
// ---------------------------------------------------------------------------

template <class Fp>
__device__ __forceinline__ XYZZ<Fp> shfl_down_xyzz(const XYZZ<Fp>& p, int d) {
    XYZZ<Fp> r;
#pragma unroll
    for (int i = 0; i < Fp::N; i++) {
        r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], d);
        r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], d);
        r.ZZ.v[i] = __shfl_down_sync(0xffffffffu, p.ZZ.v[i], d);
        r.ZZZ.v[i] = __shfl_down_sync(0xffffffffu, p.ZZZ.v[i], d);
    }
    return r;
}
// Out-of-line point addition for the reduction kernels: keeps their register count at the cost of one
// call per ~2000-instruction addition (the accumulation kernel inlines its mixed addition instead).
template <class Fp>
__device__ __noinline__ XYZZ<Fp> xyzz_add_fn(XYZZ<Fp> a, const XYZZ<Fp> b) { a.add(b); return a; }
#define xyzz_add(a, b) ((a) = xyzz_add_fn((a), (b)))

// sum over the warp; valid in lane 0 (upper lanes add garbage-free copies of valid points, harmlessly)
template <class Fp>
__device__ __forceinline__ XYZZ<Fp> warp_sum_xyzz(XYZZ<Fp> v) {
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ<Fp> o = shfl_down_xyzz(v, d);
        xyzz_add(v, o);
    }
    return v;
}

// Buckets with more than one item: one warp adds the bucket's item sums and writes the dense slot.
template <class Fp>
__global__ void __launch_bounds__(128)
k_msm_multi_buckets(const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count,
                    const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ total_items, uint32_t nb,
                    const XYZZ<Fp>* __restrict__ partial, XYZZ<Fp>* __restrict__ buckets) {
    const uint32_t nlist = *list_count;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < nlist; k += nwarps) {
        const uint32_t b = list[k];
        const uint32_t b0 = item_off[b];
        const uint32_t b1 = (b + 1 < nb) ? item_off[b + 1] : *total_items;
        XYZZ<Fp> acc = XYZZ<Fp>::inf();
#pragma unroll 1
        for (uint32_t t = b0 + lane; t < b1; t += 32) xyzz_add(acc, ld_xyzz(partial + t));
        acc = warp_sum_xyzz(acc);
        if (lane == 0) st_xyzz(buckets + b, acc);
    }
}

// one thread, once per MSM: running count of the mixed additions the accumulation performed
static __global__ void k_msm_count_adds(const uint32_t* __restrict__ total_entries, unsigned long long* __restrict__ adds_total) {
    *adds_total += *total_entries;
}

// ---------------------------------------------------------------------------
// Bucket reduction  result = sum_b (b+1) * B_b  over the dense array of nb = 2^(c-1) bucket sums.
// Write b = h * 2^s + l (l: s low bits).  With column sums C_l = sum_h B_{h,l} and row
// sums R_h = sum_l B_{h,l}
//     result = sum_l (l+1) * C_l  +  2^s * sum_h h * R_h :
// 2^s + 2^(c-1-s) PLAIN sums, followed by one small weighted sum over ncols + nrows points, which
// is split by the bits of the weights into plain sums again
//     sum_i w_i X_i = sum_j 2^j Q_j,   Q_j = sum_{i : bit j of w_i} X_i.
// Nothing here is a long dependent chain of point additions (the first version, a
// radix-8 running-sum level plus bit-decomposed sums over 65536 entries, was
// latency-bound: 1.07 ms of a 3.4 ms MSM at c = 20).
// ---------------------------------------------------------------------------
constexpr int MSM_RC_WARPS = 8;     // rows/columns per block of k_msm_rowcol
constexpr int MSM_TAIL_THREADS = 256;
constexpr int MSM_RC_MIN_CHUNK = 16;  // buckets per thread in level 1 (raised until the launch is one wave)
constexpr int MSM_RC_BLOCKS_PER_SM = 3; // resident blocks of k_msm_rowcol_partial (its __launch_bounds__)

// Level 1: every thread adds `chunk` buckets of one column (threads [0, ncols * nch_r)) or of one row
// (the rest): 2 * nb point additions spread over ~2 * nb / 16 threads with no tree in the way -- this is where
// the reduction's work is.  The next bucket is fetched while the current one is added.
//   Pc[l * nch_r + ch] = sum_{h in chunk ch} B_{h,l}      Pr[h * nch_c + ch] = sum_{l in chunk ch} B_{h,l}
template <class Fp>
__global__ void __launch_bounds__(128, MSM_RC_BLOCKS_PER_SM)
k_msm_rowcol_partial(const XYZZ<Fp>* __restrict__ B, int s, uint32_t nb, uint32_t chunk, XYZZ<Fp>* __restrict__ Pc,
                     XYZZ<Fp>* __restrict__ Pr) {
    const uint32_t ncols = 1u << s, nrows = nb >> s;
    const uint32_t nch_r = (nrows + chunk - 1) / chunk, nch_c = (ncols + chunk - 1) / chunk;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t first, stride, cnt;
    XYZZ<Fp>* dst;
    if (t < ncols * nch_r) {
        const uint32_t l = t % ncols, ch = t / ncols;      // lanes along l: neighbouring buckets
        first = ((ch * chunk) << s) | l;
        stride = ncols;
        cnt = min(nrows, (ch + 1) * chunk) - ch * chunk;
        dst = Pc + (size_t)l * nch_r + ch;
    } else {
        t -= ncols * nch_r;
        if (t >= nrows * nch_c) return;
        const uint32_t h = t / nch_c, ch = t % nch_c;
        first = (h << s) | (ch * chunk);
        stride = 1;
        cnt = min(ncols, (ch + 1) * chunk) - ch * chunk;
        dst = Pr + (size_t)h * nch_c + ch;
    }
    XYZZ<Fp> acc = ld_xyzz(B + first);
    XYZZ<Fp> nxt = cnt > 1 ? ld_xyzz(B + first + stride) : XYZZ<Fp>::inf();
#pragma unroll 1
    for (uint32_t i = 1; i < cnt; i++) {
        const XYZZ<Fp> cur = nxt;
        if (i + 1 < cnt) nxt = ld_xyzz(B + first + (size_t)(i + 1) * stride);
        acc.add(cur);
    }
    st_xyzz(dst, acc);
}

// Level 2: X[l] = C_l for l < ncols, X[ncols + h] = R_h for h < nrows: each is the sum of its <= 64 chunk
// partials.  A block takes MSM_RC_WARPS sums at a time and adds them up as one pairwise tree through
// shared memory, so that all lanes work on the wide levels (a shuffle tree per warp spends 5 of its 7
// additions with most lanes idle).
// Levels 2 and 3 are latency-bound (~10 dependent additions each), so they are deferred until the
// results are fetched and run once for all MSMs queued since (grid.y = slots).
template <class Fp>
__global__ void __launch_bounds__(32 * MSM_RC_WARPS)
k_msm_rowcol(const XYZZ<Fp>* __restrict__ Pc, const XYZZ<Fp>* __restrict__ Pr, int s, uint32_t nb, uint32_t chunk,
             XYZZ<Fp>* __restrict__ X) {
    extern __shared__ uint4 rc_smem[];
    XYZZ<Fp>* sh = reinterpret_cast<XYZZ<Fp>*>(rc_smem);       // 32 * MSM_RC_WARPS points
    const uint32_t ncols = 1u << s, nrows = nb >> s;
    const uint32_t nch_r = (nrows + chunk - 1) / chunk, nch_c = (ncols + chunk - 1) / chunk;
    // blockIdx.y: result slot (several MSMs are finished by one launch)
    Pc += (size_t)blockIdx.y * ncols * nch_r;
    Pr += (size_t)blockIdx.y * nrows * nch_c;
    X += (size_t)blockIdx.y * (ncols + nrows);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * MSM_RC_WARPS + wid;         // the sum this warp's slice of the tree belongs to
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
    if (w < ncols + nrows) {
        const XYZZ<Fp>* src = w < ncols ? Pc + (size_t)w * nch_r : Pr + (size_t)(w - ncols) * nch_c;
        const uint32_t cnt = w < ncols ? nch_r : nch_c;
#pragma unroll 1
        for (uint32_t i = lane; i < cnt; i += 32) xyzz_add(acc, ld_xyzz(src + i));
    }
    // tree over the 32 lane sums of every warp: entry (wid, lane) at sh[wid * 32 + lane]
    sh[threadIdx.x] = acc;
    __syncthreads();
#pragma unroll 1
    for (uint32_t width = 16; width >= 1; width >>= 1) {
        // MSM_RC_WARPS * width additions, packed onto the first threads of the block
        const uint32_t t = threadIdx.x;
        XYZZ<Fp> r;
        const bool active = t < MSM_RC_WARPS * width;
        if (active) {
            const uint32_t g = t / width, i = t % width;
            r = sh[g * 32 + i];
            xyzz_add(r, sh[g * 32 + i + width]);
        }
        __syncthreads();
        if (active) sh[(t / width) * 32 + (t % width)] = r;
        __syncthreads();
    }
    if (lane == 0 && w < ncols + nrows) st_xyzz(X + w, sh[wid * 32]);
}

// weight of entry i of X in the final sum
__device__ __forceinline__ uint32_t msm_tail_weight(uint32_t i, int s) {
    const uint32_t ncols = 1u << s;
    return i < ncols ? i + 1 : (i - ncols) << s;
}

// block j: T[j] = 2^j * sum_{i : bit j of w_i} X_i; the last block to finish adds the T[j] up.
template <class Fp>
__global__ void __launch_bounds__(MSM_TAIL_THREADS)
k_msm_tail(const XYZZ<Fp>* __restrict__ X, uint32_t ntot, int s, XYZZ<Fp>* __restrict__ T, uint32_t* __restrict__ done,
           XYZZ<Fp>* __restrict__ out) {
    __shared__ XYZZ<Fp> wsum[MSM_TAIL_THREADS / 32];
    __shared__ bool last;
    const int j = blockIdx.x;
    X += (size_t)blockIdx.y * ntot;       // blockIdx.y: result slot
    T += (size_t)blockIdx.y * 32;
    done += blockIdx.y;
    out += blockIdx.y;
    XYZZ<Fp> acc = XYZZ<Fp>::inf();
#pragma unroll 1
    for (uint32_t i = threadIdx.x; i < ntot; i += MSM_TAIL_THREADS)
        if ((msm_tail_weight(i, s) >> j) & 1u) xyzz_add(acc, ld_xyzz(X + i));
    acc = warp_sum_xyzz(acc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsum[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        XYZZ<Fp> v = lane < MSM_TAIL_THREADS / 32 ? wsum[lane] : XYZZ<Fp>::inf();
#pragma unroll 1
        for (int d = MSM_TAIL_THREADS / 64; d >= 1; d >>= 1) {
            XYZZ<Fp> o = shfl_down_xyzz(v, d);
            xyzz_add(v, o);
        }
        if (lane == 0) {
#pragma unroll 1
            for (int k = 0; k < j; k++) v = v.dbl();
            st_xyzz(T + j, v);
            __threadfence();
            last = atomicAdd(done, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (!last || wid != 0) return;
    __threadfence();
    XYZZ<Fp> v = lane < (int)gridDim.x ? ld_xyzz_cg(T + lane) : XYZZ<Fp>::inf();   // written by other SMs: bypass L1
    v = warp_sum_xyzz(v);
    if (lane == 0) {
        st_xyzz(out, v);
        *done = 0;                       // ready for the next MSM that uses this slot
    }
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
    DevBuf<uint32_t> counts, offsets, ranks, item_off, entries, scan_scratch, total_items;
    DevBuf<Ext> partial, buckets, rc_col, rc_row, rc_sums, tail_T, result;
    DevBuf<uint32_t> tail_done;
    uint32_t max_items = 0;
    Profiler* prof = nullptr;
    DevBuf<uint32_t> total_entries, len_hist, len_start, big_list, big_count;
    DevBuf<uint2> order;
    uint32_t sort_blocks = 0;
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
    // low bits of the bucket index that select the column in the 2D reduction
    int split_bits() const { return (plan.c - 1) / 2; }
    uint32_t rc_chunk = MSM_RC_MIN_CHUNK;
    size_t col_partials() const { return (size_t)(1u << split_bits()) * div_up(plan.nbuckets >> split_bits(), rc_chunk); }
    size_t row_partials() const { return (size_t)(plan.nbuckets >> split_bits()) * div_up(1u << split_bits(), rc_chunk); }
    // Level 1 of the reduction is one long dependent chain per thread: a second, partly filled wave of blocks
    // would double its time, so the chunk grows until all blocks are resident at once.
    void choose_rc_chunk() {
        int dev = 0, sms = 148;
        B2P_CUDA(cudaGetDevice(&dev));
        B2P_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const size_t capacity = (size_t)sms * MSM_RC_BLOCKS_PER_SM * 128;
        const uint32_t longest = plan.nbuckets >> split_bits();          // rows >= columns
        for (rc_chunk = MSM_RC_MIN_CHUNK; rc_chunk < longest && col_partials() + row_partials() > capacity; rc_chunk++) {}
    }
    void alloc_scratch() {
        const uint32_t nb = plan.nbuckets;
        counts.alloc(nb); offsets.alloc(nb); item_off.alloc(nb);
        entries.alloc((size_t)plan.W * npoints);
        ranks.alloc((size_t)plan.W * npoints);
        sort_blocks = div_up(nb, MSM_SORT_THREADS);
        const uint32_t hist_len = MSM_CAP * sort_blocks;
        scan_scratch.alloc(scan_scratch_words(nb > hist_len ? nb : hist_len));
        len_hist.alloc(hist_len); len_start.alloc(hist_len);
        total_items.alloc(1);
        total_entries.alloc(1);
        adds_total.alloc(1);
        B2P_CUDA(cudaMemset(adds_total.p, 0, sizeof(unsigned long long)));
        max_items = nb + (uint32_t)(((uint64_t)plan.W * npoints) / MSM_CAP) + 1;
        partial.alloc(max_items);
        order.alloc(max_items);
        buckets.alloc(nb);
        big_list.alloc(nb);
        big_count.alloc(1);
        choose_rc_chunk();
        rc_col.alloc((size_t)MSM_SLOTS * col_partials());
        rc_row.alloc((size_t)MSM_SLOTS * row_partials());
        rc_sums.alloc((size_t)MSM_SLOTS * ((1u << split_bits()) + (nb >> split_bits())));
        tail_T.alloc((size_t)MSM_SLOTS * 32);
        tail_done.alloc(MSM_SLOTS);
        B2P_CUDA(cudaMemset(tail_done.p, 0, MSM_SLOTS * sizeof(uint32_t)));
        result.alloc(MSM_SLOTS);
    }

    // Queues sort + accumulation + level 1 of the reduction for n <= npoints device scalars; the MSM's
    // row/column partial sums land in result slot `slot`.  finish_async() completes the queued slots.
    void run_async(const Fr* d_scalars, uint64_t n, bool mont, cudaStream_t st, int slot = 0) {
        B2P_REQUIRE(n <= npoints, "MSM: more scalars than SRS points");
        B2P_REQUIRE(slot >= 0 && slot < MSM_SLOTS, "MSM result slot out of range");
        const uint32_t nb = plan.nbuckets;
        B2P_CUDA(cudaMemsetAsync(counts.p, 0, nb * sizeof(uint32_t), st));
        if (n) B2P_LAUNCH((k_msm_count<Fr>), div_up(n, 256), 256, 0, st, d_scalars, n, plan.c, plan.W, (int)mont, counts.p,
                          ranks.p);
        exclusive_scan_u32(counts.p, offsets.p, nb, scan_scratch.p, total_entries.p, st, ScanIdentity{});
        exclusive_scan_u32(counts.p, item_off.p, nb, scan_scratch.p, total_items.p, st, ScanCeilDiv{MSM_CAP});
        if (n) B2P_LAUNCH((k_msm_scatter<Fr>), div_up(n, 256), 256, 0, st, d_scalars, n, npoints, plan.c, plan.W, (int)mont,
                          offsets.p, ranks.p, entries.p);
        // items sorted by length, longest first
        B2P_CUDA(cudaMemsetAsync(big_count.p, 0, sizeof(uint32_t), st));
        B2P_LAUNCH(k_msm_len_hist, sort_blocks, MSM_SORT_THREADS, 0, st, counts.p, nb, sort_blocks, len_hist.p);
        exclusive_scan_u32(len_hist.p, len_start.p, MSM_CAP * sort_blocks, scan_scratch.p, (uint32_t*)nullptr, st,
                           ScanIdentity{});
        B2P_LAUNCH(k_msm_len_scatter, sort_blocks, MSM_SORT_THREADS, 0, st, counts.p, item_off.p, nb, sort_blocks,
                   len_start.p, order.p, big_list.p, big_count.p);
        B2P_CUDA(cudaMemsetAsync(buckets.p, 0, (size_t)nb * sizeof(Ext), st));     // empty buckets = infinity
        const int span = prof ? prof->begin(B2P_STAT_MSM_ACCUM_MS, st) : -1;
        B2P_LAUNCH((k_msm_accumulate<Fp>), div_up(max_items, MSM_THREADS), MSM_THREADS, 0, st, table.p, entries.p,
                   counts.p, offsets.p, item_off.p, total_items.p, order.p, partial.p, buckets.p);
        if (prof) prof->end(span, st);
        B2P_LAUNCH((k_msm_multi_buckets<Fp>), 148, 128, 0, st, big_list.p, big_count.p, item_off.p, total_items.p, nb,
                   partial.p, buckets.p);
        B2P_LAUNCH(k_msm_count_adds, 1, 1, 0, st, total_entries.p, adds_total.p);
        // reduction level 1: row/column chunk sums of the dense bucket array
        const int s = split_bits();
        const uint32_t nthreads = (uint32_t)(col_partials() + row_partials());
        B2P_LAUNCH((k_msm_rowcol_partial<Fp>), div_up(nthreads, 128), 128, 0, st, buckets.p, s, nb, rc_chunk,
                   rc_col.p + (size_t)slot * col_partials(), rc_row.p + (size_t)slot * row_partials());
    }
    // Reduction levels 2 and 3 for slots [first, first + cnt), one launch each; results (XYZZ) in result[slot].
    void finish_async(int first, int cnt, cudaStream_t st) {
        B2P_REQUIRE(first >= 0 && cnt >= 1 && first + cnt <= MSM_SLOTS, "MSM result slot out of range");
        const uint32_t nb = plan.nbuckets;
        const int s = split_bits();
        const uint32_t ncols = 1u << s, nrows = nb >> s, ntot = ncols + nrows;
        constexpr size_t rc_smem_bytes = 32 * MSM_RC_WARPS * sizeof(Ext);
        if (rc_smem_bytes > 48 * 1024)                 // G2 points (Fp2 coordinates): 64 / 96 KiB, opt-in size
            B2P_CUDA(cudaFuncSetAttribute(k_msm_rowcol<Fp>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)rc_smem_bytes));
        B2P_LAUNCH((k_msm_rowcol<Fp>), dim3(div_up(ntot, MSM_RC_WARPS), cnt), 32 * MSM_RC_WARPS, rc_smem_bytes, st,
                   rc_col.p + (size_t)first * col_partials(), rc_row.p + (size_t)first * row_partials(), s, nb, rc_chunk,
                   rc_sums.p + (size_t)first * ntot);
        const int nbits = plan.c;                     // weights are < 2^(c-1) + 1
        B2P_LAUNCH((k_msm_tail<Fp>), dim3(nbits, cnt), MSM_TAIL_THREADS, 0, st, rc_sums.p + (size_t)first * ntot, ntot, s,
                   tail_T.p + (size_t)first * 32, tail_done.p + first, result.p + first);
    }

    // synchronous convenience: returns the affine result (host)
    Aff run(const Fr* d_scalars, uint64_t n, bool mont, cudaStream_t st) {
        run_async(d_scalars, n, mont, st, 0);
        finish_async(0, 1, st);
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
