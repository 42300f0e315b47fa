// Test-only shim: the MSM's window plan and signed-digit recoding (algoplonk_b200/csrc/msm_digits.cuh, the code
// k_msm_count / k_msm_scatter execute per scalar) through a C ABI, for tests/test_host.py.
#include <cstring>
#include "../../algoplonk_b200/csrc/msm_digits.cuh"
using namespace b2p;

struct Scalar8 { static constexpr int N = 8; uint32_t v[8]; };

extern "C" void ht_msm_plan(uint64_t npoints, int bits, int force_c, int* c, int* W, uint32_t* nbuckets) {
    MsmPlan p = msm_plan(npoints, bits, force_c);
    *c = p.c; *W = p.W; *nbuckets = p.nbuckets;
}
// digits of the canonical scalar s (8 little-endian words): returns how many are non-zero
extern "C" int ht_msm_digits(const uint32_t* s, int c, int W, int* win, uint32_t* bucket, int* neg) {
    Scalar8 x;
    memcpy(x.v, s, sizeof x.v);
    int cnt = 0;
    for_each_digit(x, c, W, [&](int w, uint32_t b, bool n) { win[cnt] = w; bucket[cnt] = b; neg[cnt] = n ? 1 : 0; cnt++; });
    return cnt;
}
