// Test-only shim: runs the per-column bodies of the domain-sharded NTT (algoplonk_b200/csrc/ntt_shard_math.cuh:
// the code the combine / split kernels execute) on the host, around a plain radix-2 local transform that
// follows the stage / twiddle rule of ntt.cuh's pass kernels, for all `world` ranks in one process.
// pytest compares the result with the big-int oracle on a machine without a GPU.
//
//   hs_forward(field, logn, logg, coeffs, evals, coset): coeffs natural order (n Fr, Montgomery limbs);
//       evals[r * n/G + p] = what rank r ends up holding = A(w^brev(r n/G + p))  (DIF order, cut in G blocks)
//   hs_inverse(...): the way back.
#include <cstring>
#include <vector>
#include "../../algoplonk_b200/csrc/ntt_shard_math.cuh"
using namespace b2p;

template <class Fr> static Fr constant(uint32_t (*f)(int)) {
    Fr r;
    for (int i = 0; i < Fr::N; i++) r.v[i] = f(i);
    return r;
}
template <class Fr> static Fr root_of_unity(int logn) {
    Fr w = constant<Fr>(&Fr::Params::root);
    for (int i = logn; i < Fr::Params::TWO_ADICITY; i++) w = w.sqr();
    return w;
}
// in-place DIF (natural -> bit-reversed) / DIT (bit-reversed -> natural, unscaled): stage s pairs i, i + 2^s,
// twiddle w^(pos << (logn-1-s)) with pos = i mod 2^s  (k_ntt_pass)
template <class Fr> static void local_dif(std::vector<Fr>& a, int logn, const Fr& w) {
    for (int s = logn - 1; s >= 0; s--)
        for (size_t i = 0; i < a.size(); i++) {
            if (i >> s & 1) continue;
            const Fr tw = w.pow_u64((uint64_t)(i & ((1ull << s) - 1)) << (logn - 1 - s));
            const Fr x = a[i], y = a[i + (1ull << s)];
            a[i] = x + y;
            a[i + (1ull << s)] = (x - y) * tw;
        }
}
template <class Fr> static void local_dit(std::vector<Fr>& a, int logn, const Fr& wi) {
    for (int s = 0; s < logn; s++)
        for (size_t i = 0; i < a.size(); i++) {
            if (i >> s & 1) continue;
            const Fr tw = wi.pow_u64((uint64_t)(i & ((1ull << s) - 1)) << (logn - 1 - s));
            const Fr x = a[i], wy = a[i + (1ull << s)] * tw;
            a[i] = x + wy;
            a[i + (1ull << s)] = x - wy;
        }
}

template <class Fr, int LOGG> static void run(int logn, bool inverse, const Fr* in, Fr* out, bool coset) {
    constexpr int G = 1 << LOGG;
    const int ll = logn - LOGG;
    const uint64_t n = 1ull << logn, ln = n >> LOGG, chunk = ln >> LOGG;
    const Fr w = root_of_unity<Fr>(logn), wi = w.inverse();
    const Fr wl = root_of_unity<Fr>(ll), wli = wl.inverse();
    const Fr g = constant<Fr>(&Fr::Params::shift), gi = g.inverse();
    const Fr n_inv = Fr::from_u32(2).pow_u64(logn).inverse();
    Fr wg[4], wgi[4];
    for (int j = 0; j < 4; j++) { wg[j] = w.pow_u64(ln).pow_u64(j); wgi[j] = wi.pow_u64(ln).pow_u64(j); }
    std::vector<std::vector<Fr>> x(G, std::vector<Fr>(ln));   // every rank's exchange buffer
    if (!inverse) {
        for (int r = 0; r < G; r++) {                          // forward_local
            for (uint64_t j = 0; j < ln; j++) {
                Fr v = in[j * G + r];
                if (coset) v = v * g.pow_u64(j * G + r);
                x[r][j] = v;
            }
            local_dif(x[r], ll, wl);
        }
        for (int d = 0; d < G; d++)                            // forward_combine on rank d
            for (uint64_t q = 0; q < chunk; q++) {
                Fr e[G];
                for (int r = 0; r < G; r++) e[r] = x[r][d * chunk + q];
                shard_combine_body<Fr, LOGG>(e, w.pow_u64(shard_k1(d * chunk + q, ll)), wg);
                for (int t = 0; t < G; t++) out[d * ln + ((q << LOGG) | t)] = e[t];
            }
    } else {
        for (int s = 0; s < G; s++)                            // inverse_split on rank s
            for (uint64_t q = 0; q < chunk; q++) {
                Fr e[G];
                for (int t = 0; t < G; t++) e[t] = in[s * ln + ((q << LOGG) | t)];
                shard_split_body<Fr, LOGG>(e, wi.pow_u64(shard_k1(s * chunk + q, ll)), wgi);
                for (int r = 0; r < G; r++) x[r][s * chunk + q] = e[r];
            }
        for (int r = 0; r < G; r++) {                          // inverse_local
            local_dit(x[r], ll, wli);
            for (uint64_t j = 0; j < ln; j++) {
                Fr v = x[r][j] * n_inv;
                if (coset) v = v * gi.pow_u64(j * G + r);
                out[j * G + r] = v;
            }
        }
    }
}

template <class Fr> static int dispatch(int logn, int logg, bool inverse, const uint32_t* in, uint32_t* out, bool coset) {
    if (logg < 0 || logg > NTT_SHARD_MAX_LOGG || logn < 2 * logg) return -1;
    const Fr* a = reinterpret_cast<const Fr*>(in);
    Fr* o = reinterpret_cast<Fr*>(out);
    switch (logg) {
        case 0: run<Fr, 0>(logn, inverse, a, o, coset); break;
        case 1: run<Fr, 1>(logn, inverse, a, o, coset); break;
        case 2: run<Fr, 2>(logn, inverse, a, o, coset); break;
        default: run<Fr, 3>(logn, inverse, a, o, coset); break;
    }
    return 0;
}
extern "C" int hs_transform(int field, int logn, int logg, int inverse, const uint32_t* in, uint32_t* out, int coset) {
    return field == 0 ? dispatch<FrBn254>(logn, logg, inverse != 0, in, out, coset != 0)
                      : dispatch<FrBls12381>(logn, logg, inverse != 0, in, out, coset != 0);
}
