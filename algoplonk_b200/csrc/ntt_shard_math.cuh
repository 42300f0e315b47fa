// Index and twiddle arithmetic of the domain-sharded NTT (ntt_shard.cuh): the per-column bodies of the
// combine / split kernels.  Host + device, depends on field.cuh only, so that
// tests/csrc/ntt_shard_shim.cpp runs the very code the kernels run on a machine without a GPU.
#pragma once
#include "field.cuh"

namespace b2p {

constexpr int NTT_SHARD_MAX_LOGG = 3;            // up to the 8 GPUs of one NVSwitch box

// size-G DIF over e[0..G) (natural in, bit-reversed out); wg[j] = w_G^j, j < G/2
template <class Fr, int LOGG>
HD void shard_dft_dif(Fr (&e)[1 << LOGG], const Fr* wg) {
#pragma unroll
    for (int s = LOGG - 1; s >= 0; s--) {
#pragma unroll
        for (int j = 0; j < (1 << LOGG); j++) {
            if (j & (1 << s)) continue;
            const int jj = j | (1 << s);
            const int ti = (j & ((1 << s) - 1)) << (LOGG - 1 - s);
            const Fr u = e[j] + e[jj];
            const Fr d = e[j] - e[jj];
            e[jj] = ti == 0 ? d : d * wg[ti];
            e[j] = u;
        }
    }
}
// size-G DIT over e[0..G) (bit-reversed in, natural out, no 1/G); wgi[j] = w_G^-j
template <class Fr, int LOGG>
HD void shard_dft_dit(Fr (&e)[1 << LOGG], const Fr* wgi) {
#pragma unroll
    for (int s = 0; s < LOGG; s++) {
#pragma unroll
        for (int j = 0; j < (1 << LOGG); j++) {
            if (j & (1 << s)) continue;
            const int jj = j | (1 << s);
            const int ti = (j & ((1 << s) - 1)) << (LOGG - 1 - s);
            const Fr wy = ti == 0 ? e[jj] : e[jj] * wgi[ti];
            e[jj] = e[j] - wy;
            e[j] = e[j] + wy;
        }
    }
}
// e[r] *= w1^r
template <class Fr, int LOGG>
HD void shard_twist(Fr (&e)[1 << LOGG], const Fr& w1) {
    if (LOGG == 0) return;
    Fr w = w1;
    e[1] = e[1] * w;
#pragma unroll
    for (int r = 2; r < (1 << LOGG); r++) {
        w = w * w1;
        e[r] = e[r] * w;
    }
}
// in: e[r] = C_r(k1) for the G ranks r; w1 = w_n^k1.  out: e[t] = A(w^(k1 + (n/G) brev_g(t)))
template <class Fr, int LOGG>
HD void shard_combine_body(Fr (&e)[1 << LOGG], const Fr& w1, const Fr* wg) {
    shard_twist<Fr, LOGG>(e, w1);
    shard_dft_dif<Fr, LOGG>(e, wg);
}
// in: e[t] = A(w^(k1 + (n/G) brev_g(t))); iw1 = w_n^-k1.  out: e[r] = G * C_r(k1)
template <class Fr, int LOGG>
HD void shard_split_body(Fr (&e)[1 << LOGG], const Fr& iw1, const Fr* wgi) {
    shard_dft_dit<Fr, LOGG>(e, wgi);
    shard_twist<Fr, LOGG>(e, iw1);
}
// exponent k1 of the twiddle that belongs to position q of a rank's local transform (bit-reversed order)
HD uint64_t shard_k1(uint64_t q, int local_logn) {
    uint64_t k = 0;
    for (int b = 0; b < local_logn; b++) k |= ((q >> b) & 1) << (local_logn - 1 - b);
    return k;
}

}  // namespace b2p
