// Explicit instantiation: the MSM engine over G2 (coordinates in Fp2) for both curves, and the one-shot entry point
// b2p_msm_g2 -- G2Affine.MultiExp, the other half of "MSM over G1/G2" (BASELINE north_star; not on plonk.Prove's
// path: a PLONK proving key holds two G2 points, /root/reference/setup/setup.go:124,216-225).
#define B2P_INSTANTIATE_MSM
#include "msm.cuh"
#include "fp2.cuh"
namespace b2p {

struct Bn254G2 {
    static constexpr int ID = 0;
    using Fr = FrBn254;
    using Fp = Fp2<FpBn254>;
};
struct Bls12381G2 {
    static constexpr int ID = 1;
    using Fr = FrBls12381;
    using Fp = Fp2<FpBls12381>;
};
template struct MsmEngine<Bn254G2>;
template struct MsmEngine<Bls12381G2>;

template <class C>
static void msm_g2_t(const void* points, const void* scalars, uint64_t n, void* out) {
    using Fr = typename C::Fr;
    using Aff = Affine<typename C::Fp>;
    cudaStream_t st;
    B2P_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    try {
        Aff res = Aff::inf();
        if (n) {
            MsmEngine<C> eng;
            const char* e = getenv("B2P_MSM_C");
            eng.load(points, n, e ? atoi(e) : 0, st);
            DevBuf<Fr> sc(n);
            B2P_CUDA(cudaMemcpyAsync(sc.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, st));
            res = eng.run(sc.p, n, true, st);
        }
        memcpy(out, &res, sizeof res);
    } catch (...) {
        cudaStreamDestroy(st);
        throw;
    }
    cudaStreamDestroy(st);
}
void msm_g2(int curve, const void* points, const void* scalars, uint64_t n, void* out) {
    if (curve == 0) msm_g2_t<Bn254G2>(points, scalars, n, out);
    else msm_g2_t<Bls12381G2>(points, scalars, n, out);
}

}  // namespace b2p
