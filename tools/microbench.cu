// microbench.cu -- measured INT32 roofs for the 256/384-bit kernels (SURVEY 8d asks for a measured
// IMAD peak next to the HBM one).  Build: make -C tools   Run on the GPU box: tools/microbench
//   1. mad.lo.u32 / mad.wide.u32 issue peaks (dependency-free chains, all SMs)
//   2. Montgomery multiplications / s for the four fields (register resident, ILP 2)
//   3. XYZZ mixed additions / s (the MSM bucket-accumulation inner operation)
// Prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <string>
#include "../algoplonk_b200/csrc/common.cuh"
#include "experiments/field29.cuh"
using namespace b2p;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// MODE 0: mad.lo.u32 (IMAD), 16 independent accumulators.
// (A carry-free IMAD.WIDE.U32 variant was dropped: with constant multiplicands ptxas hoists the products and the loop
//  measures IADD3, with data-dependent ones it surrounds every IMAD.WIDE with three MOV / IMAD.MOV -- neither number
//  says anything about the instruction.  The reduced-radix multiplier built on the assumption that it is cheap
//  measured 34 % SLOWER than the carry-chained one: tools/experiments/field29.cuh, reduced_radix below.)
// MODE 2: the carry-chained pair the field multiplier is made of -- mad.lo.cc.u32 / madc.hi.cc.u32 on adjacent
//         registers, which ptxas fuses into IMAD.WIDE.U32.X -- 4 chains of 4 pairs.
template <int MODE>
__global__ void __launch_bounds__(256) k_imad(uint32_t* out, int iters, uint32_t seed) {
    const uint32_t a = seed + threadIdx.x * 2654435761u, b = seed * 3 + 1 + blockIdx.x;
    if (MODE == 2) {
        uint32_t c[4][9];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int i = 0; i < 9; i++) c[k][i] = i * 7 + k + threadIdx.x;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // one row of a 4-pair product: 8 instructions = 4 IMAD.WIDE.U32(.X) after fusion
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                    "madc.lo.cc.u32 %2, %10, %9, %2;\n\tmadc.hi.cc.u32 %3, %10, %9, %3;\n\t"
                    "madc.lo.cc.u32 %4, %11, %9, %4;\n\tmadc.hi.cc.u32 %5, %11, %9, %5;\n\t"
                    "madc.lo.cc.u32 %6, %12, %9, %6;\n\tmadc.hi.u32 %7, %12, %9, %7;"
                    : "+r"(c[k][0]), "+r"(c[k][1]), "+r"(c[k][2]), "+r"(c[k][3]), "+r"(c[k][4]), "+r"(c[k][5]),
                      "+r"(c[k][6]), "+r"(c[k][7])
                    : "r"(a + k), "r"(b), "r"(a ^ 0x55u), "r"(a + 77u), "r"(a * 3u));
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int i = 0; i < 8; i++) s += c[k][i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        uint32_t c[16];
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = i + threadIdx.x;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c[i]) : "r"(b), "r"(a));
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += c[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

template <class F>
__global__ void __launch_bounds__(256) k_fmul(F* out, int iters) {
    F a = F::one(), b = F::r2(), c = F::r2();
#pragma unroll
    for (int i = 0; i < F::N - 1; i++) {      // every limb differs per lane: nothing lands on the uniform datapath
        a.v[i] += (threadIdx.x + 1) * (2 * i + 1);
        b.v[i] ^= (threadIdx.x + 7) * (2 * i + 3);
        c.v[i] += 3 * threadIdx.x + blockIdx.x + i;
    }
    for (int it = 0; it < iters; it++) {
        a = a * b;      // two independent chains
        c = c * b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + c;
}

// the same two measurements in the reduced-radix representation (field29.cuh: carry-free IMAD.WIDE products)
template <class F>
__global__ void __launch_bounds__(256) k_fmul29(F* out, int iters) {
    F a = F::one(), b = F::one(), c = F::one();
#pragma unroll
    for (int i = 0; i < F::L - 1; i++) {
        a.v[i] = (a.v[i] + (threadIdx.x + 1) * (2 * i + 1)) & F::MASK;
        b.v[i] = (b.v[i] ^ ((threadIdx.x + 7) * (2 * i + 3))) & F::MASK;
        c.v[i] = (c.v[i] + 3 * threadIdx.x + blockIdx.x + i) & F::MASK;
    }
    for (int it = 0; it < iters; it++) {
        a = a * b;
        c = c * b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = add(a, c);
}
template <class F>
__global__ void __launch_bounds__(128) k_madd29(XYZZ29<F>* out, int iters) {
    F x = F::one(), y = F::one();
#pragma unroll
    for (int i = 0; i < F::L - 1; i++) {
        x.v[i] = (x.v[i] + (threadIdx.x + 1) * (2 * i + 1)) & F::MASK;
        y.v[i] = (y.v[i] ^ ((threadIdx.x + 5) * (2 * i + 3))) & F::MASK;
    }
    XYZZ29<F> acc{x, y, F::one(), x};
    for (int it = 0; it < iters; it++) {
        acc.add_affine(x, y);
        x = add(x, y);
        x.template cond_sub<2>();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <class F> __global__ void k_one_mul29(F* io) { io[threadIdx.x + 2] = io[threadIdx.x] * io[threadIdx.x + 1]; }
template <class F> __global__ void k_one_sqr29(F* io) { io[threadIdx.x + 1] = io[threadIdx.x].sqr(); }
template __global__ void k_one_mul29<Fp29Bn254>(Fp29Bn254*);
template __global__ void k_one_sqr29<Fp29Bn254>(Fp29Bn254*);
template __global__ void k_one_mul29<Fp29Bls12381>(Fp29Bls12381*);
template __global__ void k_one_sqr29<Fp29Bls12381>(Fp29Bls12381*);

template <class Fp>
__global__ void __launch_bounds__(128) k_madd(XYZZ<Fp>* out, int iters) {
    // accumulate a fixed affine-looking operand; values need not be on the curve for timing
    Fp x = Fp::r2(), y = Fp::one();
#pragma unroll
    for (int i = 0; i < Fp::N - 1; i++) { x.v[i] += (threadIdx.x + 1) * (2 * i + 1); y.v[i] ^= (threadIdx.x + 5) * (2 * i + 3); }
    XYZZ<Fp> acc = XYZZ<Fp>::from_affine(Affine<Fp>{x, y});
    acc.ZZ = Fp::r2(); acc.ZZZ = x;
    for (int it = 0; it < iters; it++) {
        acc.add_affine(x, y);
        x = x + y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- one field / curve operation per thread, straight line: the kernels tools/sass_counts.py counts SASS
// instructions of (profiles/sass_counts_r2.json); never timed
template <class F> __global__ void k_one_ldst(F* io) { io[threadIdx.x + 1] = io[threadIdx.x]; }
template <class F> __global__ void k_one_mul(F* io) { io[threadIdx.x + 2] = io[threadIdx.x] * io[threadIdx.x + 1]; }
template <class F> __global__ void k_one_sqr(F* io) { io[threadIdx.x + 1] = io[threadIdx.x].sqr(); }
template <class F> __global__ void k_one_mul_sub(F* io) {
    io[threadIdx.x + 4] = F::mul_sub(io[threadIdx.x], io[threadIdx.x + 1], io[threadIdx.x + 2], io[threadIdx.x + 3]);
}
template <class F> __global__ void k_one_add(F* io) { io[threadIdx.x + 2] = io[threadIdx.x] + io[threadIdx.x + 1]; }
template <class Fp> __global__ void k_one_madd(XYZZ<Fp>* acc, const Affine<Fp>* p) {
    XYZZ<Fp> a = acc[threadIdx.x];
    a.add_affine(p[threadIdx.x].x, p[threadIdx.x].y);
    acc[threadIdx.x] = a;
}
template __global__ void k_one_ldst<FpBn254>(FpBn254*);
template __global__ void k_one_mul<FpBn254>(FpBn254*);
template __global__ void k_one_sqr<FpBn254>(FpBn254*);
template __global__ void k_one_mul_sub<FpBn254>(FpBn254*);
template __global__ void k_one_add<FpBn254>(FpBn254*);
template __global__ void k_one_madd<FpBn254>(XYZZ<FpBn254>*, const Affine<FpBn254>*);
template __global__ void k_one_ldst<FpBls12381>(FpBls12381*);
template __global__ void k_one_mul<FpBls12381>(FpBls12381*);
template __global__ void k_one_sqr<FpBls12381>(FpBls12381*);
template __global__ void k_one_mul_sub<FpBls12381>(FpBls12381*);
template __global__ void k_one_add<FpBls12381>(FpBls12381*);
template __global__ void k_one_madd<FpBls12381>(XYZZ<FpBls12381>*, const Affine<FpBls12381>*);

// ---- batch-affine addition, the alternative to the XYZZ accumulator -------------------
// out[i] = A[i] + B[i] for n independent pairs of affine points with distinct x (the bucket-pair rounds of a
// batch-affine Pippenger).  Thread t owns pairs t, t + T, t + 2T, ... (coalesced), K = ceil(n / T) of them:
//   pass A  d_k = xB - xA, running product p_k = p_(k-1) d_k  -> scratch          1 M
//   one Fermat inversion of p_(K-1) per thread                                      ~380 M / K
//   pass B  (reverse) 1/d_k = inv * p_(k-1); inv *= d_k;  lambda = (yB - yA)/d_k;  x3 = lambda^2 - xA - xB;
//           y3 = lambda (xA - x3) - yA                                              2 M + 2 M + 1 S
// i.e. 5 M + 1 S + 380 M / K per addition against 6 M + 2 S + (a b - c d) for the XYZZ mixed addition.
template <class Fp>
__global__ void __launch_bounds__(128) k_batch_affine(const Affine<Fp>* __restrict__ A, const Affine<Fp>* __restrict__ B,
                                                      Affine<Fp>* __restrict__ out, Fp* __restrict__ scratch, uint32_t n) {
    const uint32_t T = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    Fp p = Fp::one();
    for (uint32_t i = t; i < n; i += T) {
        const Fp d = ld_field(&B[i].x) - ld_field(&A[i].x);
        p = p * d;
        st_field(scratch + i, p);
    }
    Fp inv = p.inverse();
    const uint32_t cnt = t < n ? (n - 1 - t) / T + 1 : 0;
    for (uint32_t k = cnt; k-- > 0;) {
        const uint32_t i = t + k * T;
        const Fp xa = ld_field(&A[i].x), ya = ld_field(&A[i].y), xb = ld_field(&B[i].x), yb = ld_field(&B[i].y);
        const Fp prev = k ? ld_field(scratch + (i - T)) : Fp::one();
        const Fp dinv = inv * prev;
        inv = inv * (xb - xa);
        const Fp lam = (yb - ya) * dinv;
        const Fp x3 = lam.sqr() - xa - xb;
        const Fp y3 = Fp::mul_sub(lam, xa - x3, ya, Fp::one());
        st_field(&out[i].x, x3);
        st_field(&out[i].y, y3);
    }
}
// the same additions with the XYZZ formulas the MSM uses today (accumulator = A[i] lifted, one mixed addition, no
// conversion back): the per-addition cost the batch-affine kernel has to beat, on the same memory traffic pattern
template <class Fp>
__global__ void __launch_bounds__(128) k_pair_xyzz(const Affine<Fp>* __restrict__ A, const Affine<Fp>* __restrict__ B,
                                                   XYZZ<Fp>* __restrict__ out, uint32_t n) {
    const uint32_t T = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += T) {
        XYZZ<Fp> acc;
        acc.X = ld_field(&A[i].x); acc.Y = ld_field(&A[i].y); acc.ZZ = ld_field(&B[i].y); acc.ZZZ = ld_field(&B[i].x);
        acc.add_affine(ld_field(&B[i].x), ld_field(&B[i].y));
        st_field(&out[i].X, acc.X); st_field(&out[i].Y, acc.Y); st_field(&out[i].ZZ, acc.ZZ); st_field(&out[i].ZZZ, acc.ZZZ);
    }
}
// points (i + 1 + off) * G, affine: correct inputs so that the result can be checked
template <class Fp>
__global__ void k_make_points(Affine<Fp>* out, uint32_t n, uint32_t off, Affine<Fp> g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const XYZZ<Fp> P = XYZZ<Fp>::from_affine(g).mul_small((uint64_t)i + 1 + off);
    const Affine<Fp> a = P.to_affine();
    st_field(&out[i].x, a.x);
    st_field(&out[i].y, a.y);
}
// err += 1 for every i (sampled) where out[i] != A[i] + B[i] computed with the XYZZ formulas
template <class Fp>
__global__ void k_check_pairs(const Affine<Fp>* A, const Affine<Fp>* B, const Affine<Fp>* out, uint32_t n, uint32_t step,
                              uint32_t* err) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * step;
    if (i >= n) return;
    XYZZ<Fp> acc = XYZZ<Fp>::from_affine(Affine<Fp>{ld_field(&A[i].x), ld_field(&A[i].y)});
    acc.add_affine(ld_field(&B[i].x), ld_field(&B[i].y));
    const Affine<Fp> want = acc.to_affine();
    if (want.x != ld_field(&out[i].x) || want.y != ld_field(&out[i].y)) atomicAdd(err, 1u);
}

template <class K, class... A>
static float time_kernel(K kern, dim3 grid, dim3 block, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // warm-up: clocks ramp from idle, the first launches of a fresh process are not representative
    for (int r = 0; r < 8; r++) kern<<<grid, block>>>(args...);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kern<<<grid, block>>>(args...);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void* buf;
    CK(cudaMalloc(&buf, (size_t)sms * 16 * 256 * sizeof(XYZZ<FpBls12381>)));
    const int iters = 4096;
    dim3 grid(sms * 8), block(256);
    double thr = (double)grid.x * block.x;
    float t_lo = time_kernel(k_imad<0>, grid, block, (uint32_t*)buf, iters, 7u);
    float t_chain = time_kernel(k_imad<2>, grid, block, (uint32_t*)buf, iters, 7u);
    double imad_lo = thr * iters * 16 / (t_lo * 1e-3);
    double imad_wide_x = thr * iters * 16 / (t_chain * 1e-3);     // fused pairs: 4 chains x 4 IMAD.WIDE.U32.X
    const int fi = 2048;
    float t1 = time_kernel(k_fmul<FrBn254>, grid, block, (FrBn254*)buf, fi);
    float t2 = time_kernel(k_fmul<FpBn254>, grid, block, (FpBn254*)buf, fi);
    float t3 = time_kernel(k_fmul<FrBls12381>, grid, block, (FrBls12381*)buf, fi);
    float t4 = time_kernel(k_fmul<FpBls12381>, grid, block, (FpBls12381*)buf, fi);
    dim3 g2(sms * 8), b2(128);
    double thr2 = (double)g2.x * b2.x;
    const int mi = 512;
    float t5 = time_kernel(k_madd<FpBn254>, g2, b2, (XYZZ<FpBn254>*)buf, mi);
    float t6 = time_kernel(k_madd<FpBls12381>, g2, b2, (XYZZ<FpBls12381>*)buf, mi);
    float u1 = time_kernel(k_fmul29<Fp29Bn254>, grid, block, (Fp29Bn254*)buf, fi);
    float u2 = time_kernel(k_fmul29<Fr29Bls12381>, grid, block, (Fr29Bls12381*)buf, fi);
    float u3 = time_kernel(k_fmul29<Fp29Bls12381>, grid, block, (Fp29Bls12381*)buf, fi);
    float u5 = time_kernel(k_madd29<Fp29Bn254>, g2, b2, (XYZZ29<Fp29Bn254>*)buf, mi);
    float u6 = time_kernel(k_madd29<Fp29Bls12381>, g2, b2, (XYZZ29<Fp29Bls12381>*)buf, mi);
    // batch-affine vs XYZZ on n independent pairs (BN254): n = 6.8 M is the first pairing round of ONE 2^20-point
    // MSM (13.6 M digits), 20.4 M that of three MSMs taken together
    std::string ba = "[";
    {
        typedef FpBn254 Fp;
        const uint32_t nmax = 20400000;
        Affine<Fp>*A, *B, *O; Fp* S; XYZZ<Fp>* X; uint32_t* err;
        CK(cudaMalloc(&A, (size_t)nmax * sizeof(Affine<Fp>))); CK(cudaMalloc(&B, (size_t)nmax * sizeof(Affine<Fp>)));
        CK(cudaMalloc(&O, (size_t)nmax * sizeof(Affine<Fp>))); CK(cudaMalloc(&S, (size_t)nmax * sizeof(Fp)));
        CK(cudaMalloc(&X, (size_t)nmax * sizeof(XYZZ<Fp>))); CK(cudaMalloc(&err, 4));
        Affine<Fp> g; g.x = Fp::from_u32(1); g.y = Fp::from_u32(2);
        k_make_points<Fp><<<(nmax + 127) / 128, 128>>>(A, nmax, 0u, g);
        k_make_points<Fp><<<(nmax + 127) / 128, 128>>>(B, nmax, 0x40000000u, g);
        CK(cudaDeviceSynchronize());
        const uint32_t ns[2] = {6800000u, 20400000u};
        const int Ks[6] = {32, 64, 128, 256, 512, 1024};
        for (int a = 0; a < 2; a++) {
            const uint32_t n = ns[a];
            const float tx = time_kernel(k_pair_xyzz<Fp>, dim3(sms * 16), dim3(128), A, B, X, n);
            for (int b = 0; b < 6; b++) {
                const uint32_t T = (n + Ks[b] - 1) / Ks[b];
                const dim3 gr((T + 127) / 128), bl(128);
                CK(cudaMemset(err, 0, 4));
                const float tb = time_kernel(k_batch_affine<Fp>, gr, bl, A, B, O, S, n);
                k_check_pairs<Fp><<<(n / 997 + 127) / 128, 128>>>(A, B, O, n, 997u, err);
                uint32_t herr = 0;
                CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
                char buf2[256];
                snprintf(buf2, sizeof buf2, "%s{\"pairs\": %u, \"K\": %d, \"threads\": %u, \"batch_affine_adds_per_s\": %.4g, "
                         "\"xyzz_adds_per_s\": %.4g, \"mismatches\": %u}", (a || b) ? ", " : "", n, Ks[b], gr.x * 128,
                         n / (tb * 1e-3), n / (tx * 1e-3), herr);
                ba += buf2;
            }
        }
        ba += "]";
        cudaFree(A); cudaFree(B); cudaFree(O); cudaFree(S); cudaFree(X); cudaFree(err);
    }
    CK(cudaGetLastError());
    printf("{\"sms\": %d, \"imad_lo_per_s\": %.4g, \"imad_wide_x_carry_chain_per_s\": %.4g, "
           "\"fmul_per_s\": {\"fr_bn254\": %.4g, \"fp_bn254\": %.4g, \"fr_bls12381\": %.4g, \"fp_bls12381\": %.4g}, "
           "\"xyzz_madd_per_s\": {\"bn254\": %.4g, \"bls12381\": %.4g}, "
           "\"reduced_radix\": {\"fmul_per_s\": {\"fp_bn254_9x29\": %.4g, \"fr_bls12381_9x29\": %.4g, \"fp_bls12381_14x28\": %.4g}, "
           "\"xyzz_madd_per_s\": {\"bn254\": %.4g, \"bls12381\": %.4g}}, \"batch_affine_bn254\": %s}\n",
           sms, imad_lo, imad_wide_x, thr * fi * 2 / (t1 * 1e-3), thr * fi * 2 / (t2 * 1e-3), thr * fi * 2 / (t3 * 1e-3),
           thr * fi * 2 / (t4 * 1e-3), thr2 * mi / (t5 * 1e-3), thr2 * mi / (t6 * 1e-3),
           thr * fi * 2 / (u1 * 1e-3), thr * fi * 2 / (u2 * 1e-3), thr * fi * 2 / (u3 * 1e-3),
           thr2 * mi / (u5 * 1e-3), thr2 * mi / (u6 * 1e-3), ba.c_str());
    return 0;
}
