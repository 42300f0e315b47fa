// Shared plumbing: curve configs, CUDA error handling, device buffers, vector loads.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>
#include "ec.cuh"
#include "../../include/b200plonk.h"
#include "iface.hpp"

namespace b2p {

struct Bn254 {
    static constexpr int ID = 0;
    using Fr = FrBn254;
    using Fp = FpBn254;
    static constexpr const char* NAME = "BN254";
};
struct Bls12381 {
    static constexpr int ID = 1;
    using Fr = FrBls12381;
    using Fp = FpBls12381;
    static constexpr const char* NAME = "BLS12-381";
};


#define B2P_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw ::b2p::Error(-2, std::string("CUDA error ") + cudaGetErrorString(_e) + " at " + \
                                       __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
    } while (0)

#define B2P_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) throw ::b2p::Error(-1, std::string(msg));   \
    } while (0)

// Launch counter: every kernel launch of this library goes through B2P_LAUNCH so
// that bench.py can report "gpu_launches" as a measured number.
extern std::atomic<unsigned long long> g_launch_count;

#define B2P_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
        ::b2p::g_launch_count++;                                                \
        B2P_CUDA(cudaGetLastError());                                           \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) B2P_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
    operator T*() const { return p; }
};

// ---------------------------------------------------------------------------
// profiling spans (CUDA events on the prover stream; only when enabled)
// ---------------------------------------------------------------------------
struct Profiler {
    bool on = false;
    struct Span { int slot; cudaEvent_t a, b; };
    std::vector<Span> spans;
    int begin(int slot, cudaStream_t st) {
        if (!on) return -1;
        Span s{slot, nullptr, nullptr};
        cudaEventCreate(&s.a);
        cudaEventCreate(&s.b);
        cudaEventRecord(s.a, st);
        spans.push_back(s);
        return (int)spans.size() - 1;
    }
    void end(int id, cudaStream_t st) {
        if (id >= 0) cudaEventRecord(spans[id].b, st);
    }
    void collect(double* stats) {
        for (auto& s : spans) {
            float ms = 0;
            cudaEventSynchronize(s.b);
            cudaEventElapsedTime(&ms, s.a, s.b);
            stats[s.slot] += ms;
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        spans.clear();
    }
};


inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// 16-byte vector loads/stores of field elements (32 B or 48 B, 16 B aligned).
template <class F>
__device__ __forceinline__ F ld_field(const F* p) {
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) {
        uint4 t = q[i];
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
template <class F>
__device__ __forceinline__ F ldg_field(const F* p) {   // read-only path
    F r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) {
        uint4 t = __ldg(q + i);
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
template <class F>
__device__ __forceinline__ void st_field(F* p, const F& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < F::N / 4; i++) q[i] = make_uint4(r.v[4 * i], r.v[4 * i + 1], r.v[4 * i + 2], r.v[4 * i + 3]);
}

}  // namespace b2p
