// Exclusive prefix sum over uint32 (bucket histogram -> bucket offsets).
// Three launches: per-block scan + block totals, scan of the totals (one block,
// recursive if needed), uniform add.  Sizes here are <= 2^22 so this is far off
// the critical path.
#pragma once
#include "common.cuh"

namespace b2p {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                       // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048

// XF: transform applied to the input before scanning (e.g. ceil(cnt / CAP)).
struct ScanIdentity { __device__ uint32_t operator()(uint32_t x) const { return x; } };
struct ScanCeilDiv {
    uint32_t d;
    __device__ uint32_t operator()(uint32_t x) const { return (x + d - 1) / d; }
};

template <class XF>
__global__ void k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                             uint32_t* __restrict__ tile_sums, uint32_t n, XF xf) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? xf(in[base + i]) : 0u;
        sum += v[i];
    }
    // inclusive warp scan of per-thread sums
    uint32_t incl = sum;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    uint32_t excl = incl - sum + (wid ? warp_sums[wid - 1] : 0u);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == SCAN_THREADS - 1 && tile_sums) tile_sums[blockIdx.x] = excl;
}

static __global__ void k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_offsets, uint32_t n) {
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const uint32_t add = tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) out[base + i] += add;
}

// scratch must hold scan_scratch_words(n) uint32.  If total != nullptr the grand
// total is written there (device pointer).
inline size_t scan_scratch_words(uint32_t n) {
    size_t words = 0;
    while (n > 1) {
        n = div_up(n, SCAN_TILE);
        words += n + 1;
        if (n == 1) break;
    }
    return words + 2;
}

template <class XF>
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch, uint32_t* total,
                        cudaStream_t st, XF xf) {
    if (n == 0) return;
    const uint32_t tiles = div_up(n, SCAN_TILE);
    uint32_t* sums = scratch;
    B2P_LAUNCH((k_scan_tiles<XF>), tiles, SCAN_THREADS, 0, st, in, out, sums, n, xf);
    if (tiles == 1) {
        if (total) B2P_CUDA(cudaMemcpyAsync(total, sums, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        return;
    }
    // scan the tile sums in place (recursive), then add back
    exclusive_scan_u32(sums, sums, tiles, scratch + tiles + 1, total, st, ScanIdentity{});
    B2P_LAUNCH(k_scan_add, tiles, SCAN_THREADS, 0, st, out, sums, n);
}

}  // namespace b2p
