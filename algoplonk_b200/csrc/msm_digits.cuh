// Window plan and signed-digit recoding of the MSM (split out of msm.cuh so that the host tests can run the very
// code the kernels execute: tests/csrc/msm_digits_shim.cpp, tests/test_host.py).
#pragma once
#include <cstdint>
#include "ptx.cuh"

namespace b2p {

struct MsmPlan {
    int c = 0;        // window bits
    int W = 0;        // windows
    uint32_t nbuckets = 0;   // 2^(c-1)
};

// Cost model, in mixed additions, fitted to profiles/msm_plan_sweep_r2_*.jsonl (one MSM of n = 2^15 ... 2^21 points at
// every window width, B200) and the per-kernel times of profiles/msm_once_r2_c*.txt:
//   n W                     digit additions (the accumulation: one thread per bucket slice of <= 128 entries),
//   x (1 + 0.19 log2(300 k / items))  when there are fewer work items than fill the GPU a few times over
//                           (148 SMs x 512 resident threads x 4): the accumulation is then latency-bound -- 2^19 points:
//                           c = 17 (88 k items) 2.31 ms, c = 20 (524 k) 1.81 ms, although c = 17 makes fewer additions;
//   >= 30 k x longest chain the accumulation cannot finish before its longest item: one dependent mixed addition is
//                           ~3.3 us (22 k additions' worth of throughput; 30 k with the merging of split buckets).  The partial TOP window matters here: it has
//                           t = bits + 1 - (W-1) c real bits, so its n digits land in only 2^(t-1) buckets -- 2^17
//                           points, c = 19: t = 8, 1024 entries per bucket, 128-entry items: 0.67 ms of accumulation
//                           against 0.21 ms at c = 20 (t = 15);
//   + 3 * 2^(c-1)           the bucket reduction.
// Only the smallest c of each window count W is a candidate: a wider window with the same W adds buckets, not speed.
inline MsmPlan msm_plan(uint64_t npoints, int scalar_bits, int force_c = 0) {
    MsmPlan best;
    double best_cost = 1e300;
    for (int c = 2; c <= 22; c++) {
        if (force_c && c != force_c) continue;
        const int W = (scalar_bits + 1 + c - 1) / c;
        if (!force_c && c > 2 && (scalar_bits + 1 + c - 2) / (c - 1) == W) continue;     // c - 1 reaches the same W
        const double adds = (double)npoints * W, nb = (double)(1u << (c - 1));
        double items = adds / 128.0 > nb ? adds / 128.0 : nb;
        if (items > adds) items = adds;
        double slow = 1.0;
        for (double t = items; t < 300000.0 && t >= 1.0; t *= 2.0) slow += 0.19;
        const int t_top = scalar_bits + 1 - (W - 1) * c;                                  // >= 1
        double chain = (double)npoints * (W - 1) / nb + (double)npoints / (double)(1ull << (t_top - 1));
        if (chain > 128.0) chain = 128.0;
        double acc = adds * slow;
        if (acc < 30000.0 * chain) acc = 30000.0 * chain;
        const double cost = acc + 3.0 * nb;
        if (cost < best_cost) { best_cost = cost; best.c = c; best.W = W; best.nbuckets = 1u << (c - 1); }
    }
    return best;
}

// ---------------------------------------------------------------------------
// digit extraction
// ---------------------------------------------------------------------------
template <class Fr>
HD uint32_t window_bits(const Fr& s, int off, int c) {
    const int limb = off >> 5, sh = off & 31;
    uint64_t lo = limb < Fr::N ? s.v[limb] : 0u;
    uint64_t hi = limb + 1 < Fr::N ? s.v[limb + 1] : 0u;
    uint64_t t = (lo | (hi << 32)) >> sh;
    return (uint32_t)(t & ((1u << c) - 1));
}

// Calls f(w, bucket, neg) for every non-zero signed digit of s (canonical form).
template <class Fr, class Fn>
HD void for_each_digit(const Fr& s, int c, int W, Fn f) {
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

}  // namespace b2p
