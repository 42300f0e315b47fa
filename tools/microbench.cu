// microbench.cu -- measured INT32 roofs for the 256/384-bit kernels (SURVEY 8d asks for a measured
// IMAD peak next to the HBM one).  Build: make -C tools   Run on the GPU box: tools/microbench
//   1. mad.lo.u32 / mad.wide.u32 issue peaks (dependency-free chains, all SMs)
//   2. Montgomery multiplications / s for the four fields (register resident, ILP 2)
//   3. XYZZ mixed additions / s (the MSM bucket-accumulation inner operation)
// Prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../algoplonk_b200/csrc/ec.cuh"
using namespace b2p;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int WIDE>
__global__ void k_imad(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
    if (WIDE) {
        uint64_t c[8];
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = i + threadIdx.x;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(c[i]) : "r"(b));
        }
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
    } else {
        uint32_t c[8];
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = i + threadIdx.x;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c[i]) : "r"(b), "r"(a));
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[i];
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
    float t_wide = time_kernel(k_imad<1>, grid, block, (uint32_t*)buf, iters, 7u);
    double imad_lo = thr * iters * 8 / (t_lo * 1e-3), imad_wide = thr * iters * 8 / (t_wide * 1e-3);
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
    CK(cudaGetLastError());
    printf("{\"sms\": %d, \"imad_lo_per_s\": %.4g, \"imad_wide_per_s\": %.4g, "
           "\"fmul_per_s\": {\"fr_bn254\": %.4g, \"fp_bn254\": %.4g, \"fr_bls12381\": %.4g, \"fp_bls12381\": %.4g}, "
           "\"xyzz_madd_per_s\": {\"bn254\": %.4g, \"bls12381\": %.4g}}\n",
           sms, imad_lo, imad_wide, thr * fi * 2 / (t1 * 1e-3), thr * fi * 2 / (t2 * 1e-3), thr * fi * 2 / (t3 * 1e-3),
           thr * fi * 2 / (t4 * 1e-3), thr2 * mi / (t5 * 1e-3), thr2 * mi / (t6 * 1e-3));
    return 0;
}
