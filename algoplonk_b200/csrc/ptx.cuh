// Carry-chain primitives.  Device: inline PTX (add.cc / madc.lo.cc / madc.hi.cc ...),
// which ptxas turns into IADD3.X / IMAD.WIDE.U32(.X) chains on sm_100a.
// Host (gcc, or the host pass of nvcc): a bit-exact emulation with an explicit
// carry flag, so that the very same field code can be unit-tested on a machine
// without a GPU and reused by the host side of the prover for the handful of
// scalar computations it does (challenges, zeta^n, Jacobian -> affine, ...).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#else
#define HD inline
#define HDN
#endif

namespace b2p {
namespace ptx {

#if defined(__CUDA_ARCH__)

__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

#else  // host emulation -----------------------------------------------------

namespace detail {
inline uint32_t& cf() { static thread_local uint32_t flag = 0; return flag; }
}
inline uint32_t add_cc(uint32_t a, uint32_t b) {
    uint64_t s = (uint64_t)a + b; detail::cf() = (uint32_t)(s >> 32); return (uint32_t)s;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint64_t s = (uint64_t)a + b + detail::cf(); detail::cf() = (uint32_t)(s >> 32); return (uint32_t)s;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + detail::cf(); }
// PTX sub.cc: CC.CF holds the *borrow* after sub (1 = borrow occurred).
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint64_t d = (uint64_t)a - b; detail::cf() = (uint32_t)((d >> 32) & 1); return (uint32_t)d;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint64_t d = (uint64_t)a - b - detail::cf(); detail::cf() = (uint32_t)((d >> 32) & 1); return (uint32_t)d;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - detail::cf(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }

#endif

}  // namespace ptx
}  // namespace b2p
