// extern "C" surface of libb200plonk.so -- see include/b200plonk.h for the contract
// and the reference interface each entry point replaces.  This file only does
// argument checking, curve dispatch and error translation; the work is in the
// per-curve instantiations (inst_*.cu).
#include <atomic>
#include <cuda_runtime.h>
#include <mutex>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>
#include <cstdio>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "../../include/b200plonk.h"
#include "iface.hpp"

namespace b2p {
std::atomic<unsigned long long> g_launch_count{0};
}
using namespace b2p;

#define API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_err;

static void require(bool cond, const char* msg) {
    if (!cond) throw Error(B2P_ERR_ARG, msg);
}

template <class Fn>
static int guarded(Fn fn) {
    try {
        fn();
        return B2P_OK;
    } catch (const Error& e) {
        g_err = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        g_err = "host allocation failed";
        return B2P_ERR_INTERNAL;
    } catch (const std::exception& e) {
        g_err = e.what();
        return B2P_ERR_INTERNAL;
    } catch (...) {
        g_err = "unknown error";
        return B2P_ERR_INTERNAL;
    }
}

// Handles are bound to the device that was current when they were created; calls may arrive on any host
// thread (cgo moves goroutines between OS threads, Python worker threads start on device 0), so every entry
// point that touches a handle makes that device current for the duration of the call.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        if (cudaGetDevice(&prev) != cudaSuccess) throw Error(B2P_ERR_CUDA, "CUDA error: no current device");
        if (prev != dev) {
            if (cudaSetDevice(dev) != cudaSuccess) throw Error(B2P_ERR_CUDA, "CUDA error: cannot switch device");
            switched = true;
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
static int current_device() {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) throw Error(B2P_ERR_CUDA, "CUDA error: no current device");
    return d;
}

static const CurveOps* ops_for(int curve) {
    if (curve == B2P_BN254) return curve_ops_bn254();
    if (curve == B2P_BLS12_381) return curve_ops_bls12381();
    throw Error(B2P_ERR_ARG, "unsupported curve id (B2P_BN254 = 0, B2P_BLS12_381 = 1)");
}

API const char* b2p_last_error(void) { return g_err.c_str(); }
API const char* b2p_version(void) { return "b200plonk 0.1 (sm_100a)"; }
API uint64_t b2p_launch_count(void) { return g_launch_count; }

API int b2p_init(int device) {
    return guarded([&] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error(B2P_ERR_CUDA, std::string("CUDA error: no usable device: ") + cudaGetErrorString(e));
        if (device >= 0) {
            require(device < count, "device index out of range");
            e = cudaSetDevice(device);
            if (e != cudaSuccess) throw Error(B2P_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
        }
        e = cudaFree(nullptr);
        if (e != cudaSuccess) throw Error(B2P_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
    });
}

API int b2p_host_alloc(uint64_t bytes, void** out) {
    return guarded([&] {
        require(out && bytes > 0, "null argument");
        cudaError_t e = cudaMallocHost(out, bytes);
        if (e != cudaSuccess) throw Error(B2P_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
    });
}
API void b2p_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

API int b2p_srs_load(int curve, const void* g1, uint64_t n_can, const void* g1_lag, uint64_t n_lag, b2p_srs** out) {
    (void)g1_lag; (void)n_lag;   // see header: Lagrange commitments are iNTT + canonical MSM
    return guarded([&] {
        require(g1 && out, "null argument");
        SrsBase* s = ops_for(curve)->new_srs();
        s->device = current_device();
        try { s->load(g1, n_can); } catch (...) { delete s; throw; }
        *out = reinterpret_cast<b2p_srs*>(s);
    });
}

API int b2p_srs_load_compressed(int curve, const void* pk_bin, uint64_t len, uint64_t count, b2p_srs** out) {
    return guarded([&] {
        require(pk_bin && out, "null argument");
        const uint64_t nb = curve == B2P_BN254 ? 32 : 48;
        const uint8_t* p = static_cast<const uint8_t*>(pk_bin);
        require(len >= 4, "pk.bin too small: no header");
        const uint64_t declared = ((uint64_t)p[0] << 24) | ((uint64_t)p[1] << 16) | ((uint64_t)p[2] << 8) | p[3];
        // setup/setup.go:219-223: "pk.bin too small for %d elements"
        if (declared < count || len < 4 + count * nb)
            throw Error(B2P_ERR_ARG, "pk.bin too small for " + std::to_string(count) + " elements");
        SrsBase* s = ops_for(curve)->new_srs();
        s->device = current_device();
        try { s->load_compressed(p + 4, count); } catch (...) { delete s; throw; }
        *out = reinterpret_cast<b2p_srs*>(s);
    });
}

API int b2p_srs_generate_unsafe(int curve, const void* tau, uint64_t n_can, b2p_srs** out) {
    return guarded([&] {
        require(tau && out, "null argument");
        SrsBase* s = ops_for(curve)->new_srs();
        s->device = current_device();
        try { s->generate_unsafe(tau, 0, 1, n_can); } catch (...) { delete s; throw; }
        *out = reinterpret_cast<b2p_srs*>(s);
    });
}

API int b2p_srs_generate_unsafe_range(int curve, const void* tau, uint64_t first, uint64_t count, b2p_srs** out) {
    return guarded([&] {
        require(tau && out, "null argument");
        SrsBase* s = ops_for(curve)->new_srs();
        s->device = current_device();
        try { s->generate_unsafe(tau, first, 1, count); } catch (...) { delete s; throw; }
        *out = reinterpret_cast<b2p_srs*>(s);
    });
}

API int b2p_srs_generate_unsafe_strided(int curve, const void* tau, uint64_t first, uint64_t stride, uint64_t count,
                                        b2p_srs** out) {
    return guarded([&] {
        require(tau && out, "null argument");
        require(stride >= 1, "stride must be at least 1");
        SrsBase* s = ops_for(curve)->new_srs();
        s->device = current_device();
        try { s->generate_unsafe(tau, first, stride, count); } catch (...) { delete s; throw; }
        *out = reinterpret_cast<b2p_srs*>(s);
    });
}

API int b2p_srs_get_points(const b2p_srs* srs, uint64_t first, uint64_t count, void* out) {
    return guarded([&] {
        require(srs && out, "null argument");
        DeviceGuard g(reinterpret_cast<const SrsBase*>(srs)->device);
        reinterpret_cast<const SrsBase*>(srs)->get_points(first, count, out);
    });
}
API int b2p_srs_to_lagrange(b2p_srs* srs, uint64_t n, void* out_points) {
    return guarded([&] {
        require(srs && out_points, "null argument");
        SrsBase* s = reinterpret_cast<SrsBase*>(srs);
        std::lock_guard<std::mutex> lk(s->mu);
        DeviceGuard g(s->device);
        s->to_lagrange(n, out_points);
    });
}
API uint64_t b2p_srs_size(const b2p_srs* srs) { return srs ? reinterpret_cast<const SrsBase*>(srs)->size() : 0; }
API int b2p_srs_msm_params(const b2p_srs* srs, int* c, int* windows, uint64_t* buckets) {
    return guarded([&] {
        require(srs, "null argument");
        reinterpret_cast<const SrsBase*>(srs)->msm_params(c, windows, buckets);
    });
}
API void b2p_srs_free(b2p_srs* srs) {
    if (!srs) return;
    guarded([&] {
        DeviceGuard g(reinterpret_cast<SrsBase*>(srs)->device);
        delete reinterpret_cast<SrsBase*>(srs);
    });
}

API int b2p_msm_g1(b2p_srs* srs, int basis, const void* scalars, uint64_t n, void* out_affine) {
    return guarded([&] {
        require(srs && out_affine && (scalars || n == 0), "null argument");
        std::lock_guard<std::mutex> lk(reinterpret_cast<SrsBase*>(srs)->mu);
        DeviceGuard g(reinterpret_cast<SrsBase*>(srs)->device);
        reinterpret_cast<SrsBase*>(srs)->msm_g1(basis, scalars, n, out_affine, false);
    });
}
API int b2p_msm_g1_dev(b2p_srs* srs, int basis, const void* d_scalars, uint64_t n, void* out_affine) {
    return guarded([&] {
        require(srs && out_affine && (d_scalars || n == 0), "null argument");
        std::lock_guard<std::mutex> lk(reinterpret_cast<SrsBase*>(srs)->mu);
        DeviceGuard g(reinterpret_cast<SrsBase*>(srs)->device);
        reinterpret_cast<SrsBase*>(srs)->msm_g1(basis, d_scalars, n, out_affine, true);
    });
}
API int b2p_msm_g2(int curve, const void* g2_points, const void* scalars, uint64_t n, void* out_g2_affine) {
    return guarded([&] {
        require(out_g2_affine && ((g2_points && scalars) || n == 0), "null argument");
        ops_for(curve);
        require(n < (1ull << 26), "too many points");
        current_device();
        msm_g2(curve, g2_points, scalars, n, out_g2_affine);
    });
}
API int b2p_g1_sum(int curve, const void* points, uint64_t n, void* out_affine) {
    return guarded([&] {
        require(out_affine && (points || n == 0), "null argument");
        ops_for(curve)->g1_sum(points, n, out_affine);
    });
}
API int b2p_srs_set_commit_hook(b2p_srs* srs, b2p_commit_fn fn, void* ctx) {
    return guarded([&] {
        require(srs, "null argument");
        SrsBase* s = reinterpret_cast<SrsBase*>(srs);
        std::lock_guard<std::mutex> g(s->mu);
        s->set_commit_hook(fn, ctx);
    });
}
API int b2p_device_copy(void* d_dst, const void* d_src, uint64_t bytes) {
    return guarded([&] {
        require((d_dst && d_src) || bytes == 0, "null argument");
        if (bytes) {
            // cudaMemcpy does not wait for device-to-device copies: issue on the thread's stream and wait for it
            cudaError_t e = cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, cudaStreamPerThread);
            if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
            if (e != cudaSuccess) throw Error(B2P_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e));
        }
    });
}
API void* b2p_srs_stream(b2p_srs* srs) { return srs ? reinterpret_cast<SrsBase*>(srs)->stream_handle() : nullptr; }

API int b2p_ntt(int curve, void* data, uint64_t n, int flags) {
    return guarded([&] {
        require(data, "null argument");
        ops_for(curve)->ntt(data, n, flags);
    });
}

// ---- domain-sharded NTT (ntt_shard.cuh) and the peer memory its exchange runs over -------------------------
API int b2p_ntt_shard_create(int curve, uint64_t n, uint32_t world, uint32_t rank, b2p_ntt_shard** out) {
    return guarded([&] {
        require(out, "null argument");
        ops_for(curve);
        const int dev = current_device();
        NttShardBase* s = new_ntt_shard(curve, n, world, rank);
        s->device = dev;
        *out = reinterpret_cast<b2p_ntt_shard*>(s);
    });
}
API void b2p_ntt_shard_free(b2p_ntt_shard* s) {
    if (!s) return;
    guarded([&] {
        DeviceGuard g(reinterpret_cast<NttShardBase*>(s)->device);
        delete reinterpret_cast<NttShardBase*>(s);
    });
}
API uint64_t b2p_ntt_shard_local_size(const b2p_ntt_shard* s) {
    return s ? reinterpret_cast<const NttShardBase*>(s)->local_size() : 0;
}
API uint64_t b2p_ntt_shard_chunk_size(const b2p_ntt_shard* s) {
    return s ? reinterpret_cast<const NttShardBase*>(s)->chunk_size() : 0;
}
API int b2p_ntt_shard_forward_local(b2p_ntt_shard* s, const void* d_coeffs, uint64_t local_len, int flags, void* d_x,
                                    void* stream) {
    return guarded([&] {
        require(s && d_x && (d_coeffs || local_len == 0), "null argument");
        require((flags & ~B2P_NTT_COSET) == 0, "forward transform: only B2P_NTT_COSET is a valid flag");
        DeviceGuard g(reinterpret_cast<NttShardBase*>(s)->device);
        reinterpret_cast<NttShardBase*>(s)->forward_local(d_coeffs, local_len, flags, d_x, stream);
    });
}
API int b2p_ntt_shard_forward_combine(b2p_ntt_shard* s, const void* const* d_chunks, void* d_out, void* stream) {
    return guarded([&] {
        require(s && d_chunks && d_out, "null argument");
        DeviceGuard g(reinterpret_cast<NttShardBase*>(s)->device);
        reinterpret_cast<NttShardBase*>(s)->forward_combine(d_chunks, d_out, stream);
    });
}
API int b2p_ntt_shard_inverse_split(b2p_ntt_shard* s, const void* d_evals, void* const* d_chunks, void* stream) {
    return guarded([&] {
        require(s && d_chunks && d_evals, "null argument");
        DeviceGuard g(reinterpret_cast<NttShardBase*>(s)->device);
        reinterpret_cast<NttShardBase*>(s)->inverse_split(d_evals, d_chunks, stream);
    });
}
API int b2p_ntt_shard_inverse_local(b2p_ntt_shard* s, void* d_x, int flags, void* d_out, void* stream) {
    return guarded([&] {
        require(s && d_x, "null argument");
        require((flags & ~(B2P_NTT_COSET | B2P_NTT_INVERSE)) == 0, "unknown NTT flag");
        DeviceGuard g(reinterpret_cast<NttShardBase*>(s)->device);
        reinterpret_cast<NttShardBase*>(s)->inverse_local(d_x, flags, d_out, stream);
    });
}

static void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(B2P_ERR_CUDA, std::string("CUDA error: ") + what + ": " + cudaGetErrorString(e));
}
static_assert(sizeof(cudaIpcMemHandle_t) == B2P_IPC_HANDLE_BYTES, "IPC handle size");
API int b2p_peer_alloc(uint64_t bytes, void** d_ptr, void* ipc_handle) {
    return guarded([&] {
        require(d_ptr && bytes > 0, "null argument");
        void* p = nullptr;
        cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc");
        if (ipc_handle) {
            cudaIpcMemHandle_t h;
            cudaError_t e = cudaIpcGetMemHandle(&h, p);
            if (e != cudaSuccess) { cudaFree(p); cuda_ok(e, "cudaIpcGetMemHandle"); }
            memcpy(ipc_handle, &h, sizeof h);
        }
        *d_ptr = p;
    });
}
API int b2p_peer_open(const void* ipc_handle, void** d_ptr) {
    return guarded([&] {
        require(ipc_handle && d_ptr, "null argument");
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handle, sizeof h);
        cuda_ok(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    });
}
API int b2p_peer_close(void* d_ptr) {
    return guarded([&] {
        require(d_ptr, "null argument");
        cuda_ok(cudaIpcCloseMemHandle(d_ptr), "cudaIpcCloseMemHandle");
    });
}
API int b2p_peer_free(void* d_ptr) {
    return guarded([&] {
        if (d_ptr) cuda_ok(cudaFree(d_ptr), "cudaFree");
    });
}

// ---- one proof over the GPUs of a box: commitments sharded over the point set (shard_group.cuh) -------------
API int b2p_shard_group_create(int curve, uint32_t world, uint32_t rank, uint64_t total_points, b2p_srs* shard,
                               uint64_t ntt_rows, b2p_shard_group** out) {
    return guarded([&] {
        require(shard && out, "null argument");
        SrsBase* s = reinterpret_cast<SrsBase*>(shard);
        DeviceGuard g(s->device);
        ShardGroupBase* grp = ops_for(curve)->new_shard_group(world, rank, total_points, s, ntt_rows);
        grp->device = s->device;
        *out = reinterpret_cast<b2p_shard_group*>(grp);
    });
}
API int b2p_shard_group_attach(b2p_shard_group* g, b2p_srs* prover_srs, b2p_circuit* circuit) {
    return guarded([&] {
        require(g, "null argument");
        ShardGroupBase* grp = reinterpret_cast<ShardGroupBase*>(g);
        DeviceGuard dg(grp->device);
        if (prover_srs) {
            SrsBase* s = reinterpret_cast<SrsBase*>(prover_srs);
            std::lock_guard<std::mutex> lk(s->mu);
            grp->attach(s, reinterpret_cast<CircuitBase*>(circuit));
        } else {
            grp->attach(nullptr, nullptr);
        }
    });
}
API int b2p_shard_group_export(b2p_shard_group* g, void* ipc_handles_out) {
    return guarded([&] {
        require(g && ipc_handles_out, "null argument");
        DeviceGuard dg(reinterpret_cast<ShardGroupBase*>(g)->device);
        reinterpret_cast<ShardGroupBase*>(g)->ipc_handles(ipc_handles_out);
    });
}
API int b2p_shard_group_connect(b2p_shard_group* g, const void* all_handles) {
    return guarded([&] {
        require(g && all_handles, "null argument");
        DeviceGuard dg(reinterpret_cast<ShardGroupBase*>(g)->device);
        reinterpret_cast<ShardGroupBase*>(g)->connect(all_handles);
    });
}
API int b2p_shard_group_connect_local(b2p_shard_group* const* groups, uint32_t world) {
    return guarded([&] {
        require(groups && world >= 1 && world <= 8, "null argument");
        void* all[8 * SHARD_NPTR] = {nullptr};
        for (uint32_t i = 0; i < world; i++) {
            require(groups[i] != nullptr, "null group");
            const ShardGroupBase* gi = reinterpret_cast<const ShardGroupBase*>(groups[i]);
            require(gi->world == world && gi->rank == i, "groups must be listed by rank, all of the same world");
            gi->local_ptrs(all + (size_t)i * SHARD_NPTR);
        }
        for (uint32_t i = 0; i < world; i++) {
            ShardGroupBase* gi = reinterpret_cast<ShardGroupBase*>(groups[i]);
            DeviceGuard dg(gi->device);
            gi->connect_ptrs(all);
        }
    });
}
API int b2p_shard_group_serve_proof(b2p_shard_group* g, uint64_t n) {
    return guarded([&] {
        require(g, "null argument");
        DeviceGuard dg(reinterpret_cast<ShardGroupBase*>(g)->device);
        reinterpret_cast<ShardGroupBase*>(g)->serve_proof(n);
    });
}
API int b2p_shard_group_msm(b2p_shard_group* g, const void* d_scalars, uint64_t n, void* out_affine) {
    return guarded([&] {
        require(g && out_affine && (d_scalars || n == 0), "null argument");
        DeviceGuard dg(reinterpret_cast<ShardGroupBase*>(g)->device);
        reinterpret_cast<ShardGroupBase*>(g)->msm(d_scalars, n, out_affine);
    });
}
API int b2p_shard_group_serve_msm(b2p_shard_group* g, uint64_t n) {
    return guarded([&] {
        require(g, "null argument");
        DeviceGuard dg(reinterpret_cast<ShardGroupBase*>(g)->device);
        reinterpret_cast<ShardGroupBase*>(g)->serve_msm(n);
    });
}
API void b2p_shard_group_free(b2p_shard_group* g) {
    if (!g) return;
    guarded([&] {
        DeviceGuard dg(reinterpret_cast<ShardGroupBase*>(g)->device);
        delete reinterpret_cast<ShardGroupBase*>(g);
    });
}

API int b2p_circuit_load(b2p_srs* srs, uint64_t n, uint32_t nb_public, const void* ql, const void* qr, const void* qm,
                         const void* qo, const void* qk, const int64_t* perm, uint32_t k, const void* const* qcp,
                         const uint64_t* cidx, const void* vkb, uint64_t vkb_len, b2p_circuit** out) {
    return guarded([&] {
        require(srs && ql && qr && qm && qo && qk && perm && out, "null argument");
        require(k == 0 || (qcp && cidx), "BSB22 columns missing");
        SrsBase* s = reinterpret_cast<SrsBase*>(srs);
        std::lock_guard<std::mutex> lk(s->mu);
        DeviceGuard g(s->device);
        CircuitBase* c = ops_for(s->curve)->new_circuit();
        c->device = s->device;
        c->owner = s;
        try { c->load(s, n, nb_public, ql, qr, qm, qo, qk, perm, k, qcp, cidx, vkb, vkb_len); }
        catch (...) { delete c; throw; }
        *out = reinterpret_cast<b2p_circuit*>(c);
    });
}
API int b2p_circuit_vk_commitments(b2p_circuit* c, void* out_points) {
    return guarded([&] {
        require(c && out_points, "null argument");
        std::lock_guard<std::mutex> lk(reinterpret_cast<CircuitBase*>(c)->owner->mu);
        DeviceGuard g(reinterpret_cast<CircuitBase*>(c)->device);
        reinterpret_cast<CircuitBase*>(c)->vk_commitments(out_points);
    });
}
API void b2p_circuit_free(b2p_circuit* c) {
    if (!c) return;
    guarded([&] {
        DeviceGuard g(reinterpret_cast<CircuitBase*>(c)->device);
        delete reinterpret_cast<CircuitBase*>(c);
    });
}

API uint64_t b2p_proof_raw_size(int curve, uint32_t k) {
    const uint64_t pt = curve == B2P_BN254 ? 64 : 96;
    return 9 * pt + (7 + (uint64_t)k) * 32;
}

API int b2p_prove(b2p_circuit* c, const void* L, const void* R, const void* O, const void* const* pi2,
                  const void* bsb22, const void* blinding, void* out_raw) {
    return guarded([&] {
        require(c && L && R && O && blinding && out_raw, "null argument");
        std::lock_guard<std::mutex> lk(reinterpret_cast<CircuitBase*>(c)->owner->mu);
        DeviceGuard g(reinterpret_cast<CircuitBase*>(c)->device);
        reinterpret_cast<CircuitBase*>(c)->prove(L, R, O, pi2, bsb22, blinding, out_raw, false);
    });
}
API int b2p_prove_dev(b2p_circuit* c, const void* dL, const void* dR, const void* dO, const void* const* d_pi2,
                      const void* bsb22, const void* blinding, void* out_raw) {
    return guarded([&] {
        require(c && dL && dR && dO && blinding && out_raw, "null argument");
        std::lock_guard<std::mutex> lk(reinterpret_cast<CircuitBase*>(c)->owner->mu);
        DeviceGuard g(reinterpret_cast<CircuitBase*>(c)->device);
        reinterpret_cast<CircuitBase*>(c)->prove(dL, dR, dO, d_pi2, bsb22, blinding, out_raw, true);
    });
}
API void* b2p_circuit_stream(b2p_circuit* c) { return c ? reinterpret_cast<CircuitBase*>(c)->stream_handle() : nullptr; }

API uint64_t b2p_proof_marshal_size(int curve, uint32_t k) {
    return curve == B2P_BN254 ? (24 + 3 * (uint64_t)k) * 32 : (33 + 4 * (uint64_t)k) * 32;
}
API int b2p_marshal_proof(int curve, uint32_t k, const void* raw, const void* bsb22, void* out_bytes) {
    return guarded([&] {
        require(raw && out_bytes && (k == 0 || bsb22), "null argument");
        ops_for(curve)->marshal_proof(k, raw, bsb22, static_cast<uint8_t*>(out_bytes));
    });
}
API int b2p_marshal_public_inputs(int curve, const void* values, uint32_t nb_public, void* out_bytes) {
    return guarded([&] {
        require((values && out_bytes) || nb_public == 0, "null argument");
        ops_for(curve)->marshal_public_inputs(values, nb_public, static_cast<uint8_t*>(out_bytes));
    });
}

API int b2p_circuit_set_profiling(b2p_circuit* c, int enable) {
    return guarded([&] {
        require(c, "null argument");
        reinterpret_cast<CircuitBase*>(c)->set_profiling(enable != 0);
    });
}
API int b2p_circuit_stats(const b2p_circuit* c, double* out) {
    return guarded([&] {
        require(c && out, "null argument");
        memcpy(out, reinterpret_cast<const CircuitBase*>(c)->stats, sizeof(double) * B2P_STAT_COUNT);
    });
}

// ---- plonk.Verify on the host (verify.cu) ---------------------------------------------------------------
static void require_curve(int curve) {
    require(curve == B2P_BN254 || curve == B2P_BLS12_381, "unsupported curve id (B2P_BN254 = 0, B2P_BLS12_381 = 1)");
}
API int b2p_verify(int curve, uint64_t n, uint32_t nb_public, uint32_t k, const uint64_t* commitment_indexes,
                   const void* vk_points, const void* kzg_g1, const void* kzg_g2, const void* proof_bytes,
                   uint64_t proof_len, const void* public_inputs, uint64_t public_len) {
    return guarded([&] {
        require_curve(curve);
        require(vk_points && kzg_g1 && kzg_g2 && proof_bytes && (k == 0 || commitment_indexes) &&
                    (public_len == 0 || public_inputs), "null argument");
        require(k <= 64, "too many BSB22 commitments");
        HostVerifyKey vk{n, nb_public, k, commitment_indexes, vk_points, kzg_g1, kzg_g2};
        std::string why;
        static const uint8_t none = 0;
        if (!host_verify(curve, vk, proof_bytes, proof_len, public_inputs ? public_inputs : &none, public_len, &why))
            throw Error(B2P_ERR_VERIFY, "error verifying proof: " + why);
    });
}
API int b2p_verify_batch(int curve, uint64_t n, uint32_t nb_public, uint32_t k, const uint64_t* commitment_indexes,
                         const void* vk_points, const void* kzg_g1, const void* kzg_g2, const void* proofs,
                         uint64_t proof_len, const void* public_inputs, uint64_t public_len, uint64_t count,
                         uint64_t* first_bad) {
    return guarded([&] {
        require_curve(curve);
        require(vk_points && kzg_g1 && kzg_g2 && (count == 0 || proofs) && (k == 0 || commitment_indexes) &&
                    (public_len == 0 || count == 0 || public_inputs), "null argument");
        require(k <= 64, "too many BSB22 commitments");
        require(count <= (1u << 20), "too many proofs in one batch");
        HostVerifyKey vk{n, nb_public, k, commitment_indexes, vk_points, kzg_g1, kzg_g2};
        std::string why;
        uint64_t bad = count;
        static const uint8_t none = 0;
        const bool ok = host_verify_batch(curve, vk, proofs ? proofs : &none, proof_len,
                                          public_inputs ? public_inputs : &none, public_len, count, &bad, &why);
        if (first_bad) *first_bad = bad;
        if (!ok)
            throw Error(B2P_ERR_VERIFY, "error verifying proof" +
                                            (bad < count ? " " + std::to_string(bad) : std::string(" batch")) + ": " + why);
    });
}
API int b2p_verify_batch_dev(int curve, uint64_t n, uint32_t nb_public, uint32_t k, const uint64_t* commitment_indexes,
                             const void* vk_points, const void* kzg_g1, const void* kzg_g2, const void* proofs,
                             uint64_t proof_len, const void* public_inputs, uint64_t public_len, uint64_t count,
                             uint64_t* first_bad) {
    return guarded([&] {
        require_curve(curve);
        require(vk_points && kzg_g1 && kzg_g2 && (count == 0 || proofs) && (k == 0 || commitment_indexes) &&
                    (public_len == 0 || count == 0 || public_inputs), "null argument");
        require(k <= 64, "too many BSB22 commitments");
        require(count <= (1u << 20), "too many proofs in one batch");
        current_device();                          // fails loudly without a usable GPU: no fallback to the host batch
        HostVerifyKey vk{n, nb_public, k, commitment_indexes, vk_points, kzg_g1, kzg_g2};
        std::string why;
        uint64_t bad = count;
        static const uint8_t none = 0;
        const bool ok = device_verify_batch(curve, vk, proofs ? proofs : &none, proof_len,
                                            public_inputs ? public_inputs : &none, public_len, count, &bad, &why);
        if (first_bad) *first_bad = bad;
        if (!ok)
            throw Error(B2P_ERR_VERIFY, "error verifying proof" +
                                            (bad < count ? " " + std::to_string(bad) : std::string(" batch")) + ": " + why);
    });
}
API int b2p_pairing_check(int curve, const void* g1_points, const void* g2_points, uint64_t n, int* is_one) {
    return guarded([&] {
        require_curve(curve);
        require(is_one && (n == 0 || (g1_points && g2_points)), "null argument");
        require(n <= 1024, "too many pairs");
        std::string why;
        const bool ok = host_pairing_check(curve, g1_points, g2_points, n, &why);
        if (!ok && !why.empty()) throw Error(B2P_ERR_ARG, why);
        *is_one = ok ? 1 : 0;
    });
}
API int b2p_kzg_vk_load(int curve, const void* vk_bin, uint64_t len, void* out_g2, void* out_g1) {
    return guarded([&] {
        require_curve(curve);
        require(vk_bin && out_g2 && out_g1, "null argument");
        if (const char* e = host_kzg_vk_load(curve, vk_bin, len, out_g2, out_g1)) throw Error(B2P_ERR_ARG, e);
    });
}
// ---- witness solver ---------------------------------------------------------------------------------------
API int b2p_solver_create_hinted(int curve, uint64_t n, uint32_t nb_public, uint64_t nb_variables, const uint32_t* input_ids,
                                 uint32_t nb_inputs, const void* ql, const void* qr, const void* qm, const void* qo,
                                 const void* qk, const uint32_t* xa, const uint32_t* xb, const uint32_t* xc,
                                 const b2p_hint* hints, uint32_t n_hints, const uint8_t* unchecked_rows, b2p_solver** out) {
    return guarded([&] {
        require_curve(curve);
        require(ql && qr && qm && qo && qk && xa && xb && xc && out && (input_ids || nb_inputs == 0) && (hints || n_hints == 0),
                "null argument");
        const void* cols[5] = {ql, qr, qm, qo, qk};
        const int dev = current_device();
        SolverBase* s = new_solver(curve, n, nb_public, nb_variables, input_ids, nb_inputs, cols, xa, xb, xc, hints, n_hints,
                                   unchecked_rows);
        s->device = dev;
        *out = reinterpret_cast<b2p_solver*>(s);
    });
}
API int b2p_solver_create(int curve, uint64_t n, uint32_t nb_public, uint64_t nb_variables, const uint32_t* input_ids,
                          uint32_t nb_inputs, const void* ql, const void* qr, const void* qm, const void* qo,
                          const void* qk, const uint32_t* xa, const uint32_t* xb, const uint32_t* xc, b2p_solver** out) {
    return b2p_solver_create_hinted(curve, n, nb_public, nb_variables, input_ids, nb_inputs, ql, qr, qm, qo, qk, xa, xb, xc,
                                    nullptr, 0, nullptr, out);
}
API int b2p_solver_set_hint_fn(b2p_solver* s, b2p_hint_fn fn, void* ctx) {
    return guarded([&] {
        require(s, "null argument");
        SolverBase* b = reinterpret_cast<SolverBase*>(s);
        std::lock_guard<std::mutex> lk(b->mu);
        b->set_hint_fn(fn, ctx);
    });
}
API int b2p_solver_solve(b2p_solver* s, const void* inputs, int where, void* L, void* R, void* O) {
    return guarded([&] {
        require(s && L && R && O, "null argument");
        SolverBase* b = reinterpret_cast<SolverBase*>(s);
        std::lock_guard<std::mutex> lk(b->mu);
        DeviceGuard g(b->device);
        b->solve(inputs, where, L, R, O, false, nullptr);
    });
}
API int b2p_solver_solve_dev(b2p_solver* s, const void* inputs, int where, void** dL, void** dR, void** dO) {
    return guarded([&] {
        require(s && dL && dR && dO, "null argument");
        SolverBase* b = reinterpret_cast<SolverBase*>(s);
        std::lock_guard<std::mutex> lk(b->mu);
        DeviceGuard g(b->device);
        void* p[3] = {nullptr, nullptr, nullptr};
        b->solve(inputs, where, nullptr, nullptr, nullptr, true, p);
        *dL = p[0]; *dR = p[1]; *dO = p[2];
    });
}
API int b2p_solver_info(const b2p_solver* s, uint64_t* out) {
    return guarded([&] {
        require(s && out, "null argument");
        reinterpret_cast<const SolverBase*>(s)->info(out);
    });
}
API void b2p_solver_free(b2p_solver* s) {
    if (!s) return;
    guarded([&] {
        DeviceGuard g(reinterpret_cast<SolverBase*>(s)->device);
        delete reinterpret_cast<SolverBase*>(s);
    });
}

// ---- persisted keys ---------------------------------------------------------------------------------------
API int b2p_gnark_file_parse(const void* file, uint64_t len, b2p_gnark_file* out) {
    return guarded([&] {
        require(file && out, "null argument");
        if (const char* e = host_gnark_file_parse(file, len, out)) throw Error(B2P_ERR_ARG, e);
    });
}
API int b2p_gnark_vk_parse(int curve, const void* bytes, uint64_t len, b2p_gnark_vk* out) {
    return guarded([&] {
        require_curve(curve);
        require(bytes && out, "null argument");
        if (const char* e = host_gnark_vk_parse(curve, bytes, len, out)) throw Error(B2P_ERR_ARG, e);
    });
}
API int b2p_gnark_pk_parse(int curve, const void* bytes, uint64_t len, b2p_gnark_pk* out) {
    return guarded([&] {
        require_curve(curve);
        require(bytes && out, "null argument");
        if (const char* e = host_gnark_pk_parse(curve, bytes, len, out)) throw Error(B2P_ERR_ARG, e);
    });
}

// The library's own snapshot of a proving key's circuit half: a 64-byte header, then the arguments of
// b2p_circuit_load back to back (each section padded to 64 bytes so the mapped columns are aligned).
namespace {
struct SnapHeader {
    char magic[8];                 // "B2PKEY\0\1"
    uint32_t version, curve;
    uint64_t n;
    uint32_t nb_public, k;
    uint64_t vkb_len;
    uint64_t payload_len;          // bytes after the header
    uint64_t fnv;                  // FNV-1a (64-bit, 8 bytes at a time) of the payload
    uint64_t reserved;
};
static_assert(sizeof(SnapHeader) == 64, "snapshot header is 64 bytes");
const char SNAP_MAGIC[8] = {'B', '2', 'P', 'K', 'E', 'Y', 0, 1};
inline uint64_t pad64(uint64_t x) { return (x + 63) & ~63ull; }
inline uint64_t fnv1a_words(uint64_t h, const void* p, uint64_t bytes) {      // bytes is a multiple of 8
    const uint64_t* w = static_cast<const uint64_t*>(p);
    for (uint64_t i = 0; i < bytes / 8; i++) h = (h ^ w[i]) * 0x100000001b3ull;
    return h;
}
struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
};
struct Mapping {
    void* p = MAP_FAILED;
    size_t len = 0;
    int fd = -1;
    ~Mapping() {
        if (p != MAP_FAILED) munmap(p, len);
        if (fd >= 0) close(fd);
    }
};
}  // namespace

API int b2p_circuit_save(const char* path, int curve, uint64_t n, uint32_t nb_public, const void* ql, const void* qr,
                         const void* qm, const void* qo, const void* qk, const int64_t* perm, uint32_t k,
                         const void* const* qcp, const uint64_t* cidx, const void* vkb, uint64_t vkb_len) {
    return guarded([&] {
        require_curve(curve);
        require(path && ql && qr && qm && qo && qk && perm, "null argument");
        require(k == 0 || (qcp && cidx), "BSB22 columns missing");
        require(k <= B2P_MAX_COMMITMENTS, "too many BSB22 commitments");
        require(n >= 2 && (n & (n - 1)) == 0 && n <= (1ull << 32), "domain size must be a power of two >= 2");
        require(vkb || vkb_len == 0, "null argument");
        const uint64_t col = n * 32;                       // a multiple of 64
        std::vector<std::pair<const void*, uint64_t>> parts = {{ql, col}, {qr, col}, {qm, col}, {qo, col}, {qk, col},
                                                               {perm, 3 * n * 8}};
        for (uint32_t i = 0; i < k; i++) parts.push_back({qcp[i], col});
        if (k) parts.push_back({cidx, (uint64_t)k * 8});
        if (vkb_len) parts.push_back({vkb, vkb_len});
        SnapHeader h{};
        memcpy(h.magic, SNAP_MAGIC, 8);
        h.version = 1; h.curve = (uint32_t)curve; h.n = n; h.nb_public = nb_public; h.k = k; h.vkb_len = vkb_len;
        uint64_t fnv = 0xcbf29ce484222325ull;
        static const char zeros[64] = {0};
        std::vector<char> tailbuf;
        for (auto& pr : parts) {
            h.payload_len += pad64(pr.second);
            const uint64_t whole = pr.second & ~7ull;
            fnv = fnv1a_words(fnv, pr.first, whole);
            tailbuf.assign(pad64(pr.second) - whole, 0);
            memcpy(tailbuf.data(), static_cast<const char*>(pr.first) + whole, pr.second - whole);
            fnv = fnv1a_words(fnv, tailbuf.data(), tailbuf.size());
        }
        h.fnv = fnv;
        File f;
        f.f = fopen(path, "wb");
        if (!f.f) throw Error(B2P_ERR_ARG, std::string("cannot create ") + path);
        bool ok = fwrite(&h, sizeof h, 1, f.f) == 1;
        for (auto& pr : parts) {
            ok = ok && (pr.second == 0 || fwrite(pr.first, pr.second, 1, f.f) == 1);
            const uint64_t padn = pad64(pr.second) - pr.second;
            ok = ok && (padn == 0 || fwrite(zeros, padn, 1, f.f) == 1);
        }
        ok = ok && fflush(f.f) == 0;
        if (!ok) throw Error(B2P_ERR_INTERNAL, std::string("short write to ") + path);
    });
}

API int b2p_circuit_load_file(b2p_srs* srs, const char* path, b2p_circuit** out) {
    int rc = guarded([&] {
        require(srs && path && out, "null argument");
        Mapping m;
        m.fd = open(path, O_RDONLY);
        if (m.fd < 0) throw Error(B2P_ERR_ARG, std::string("cannot open ") + path);
        struct stat sb;
        if (fstat(m.fd, &sb) != 0 || (uint64_t)sb.st_size < sizeof(SnapHeader)) throw Error(B2P_ERR_ARG, "key snapshot: file too small");
        m.len = (size_t)sb.st_size;
        m.p = mmap(nullptr, m.len, PROT_READ, MAP_PRIVATE | MAP_POPULATE, m.fd, 0);
        if (m.p == MAP_FAILED) throw Error(B2P_ERR_INTERNAL, "key snapshot: mmap failed");
        const char* base = static_cast<const char*>(m.p);
        SnapHeader h;
        memcpy(&h, base, sizeof h);
        if (memcmp(h.magic, SNAP_MAGIC, 8) != 0 || h.version != 1) throw Error(B2P_ERR_ARG, "key snapshot: bad magic / version");
        SrsBase* s = reinterpret_cast<SrsBase*>(srs);
        if ((int)h.curve != s->curve) throw Error(B2P_ERR_ARG, "key snapshot: written for the other curve");
        if (h.n < 2 || (h.n & (h.n - 1)) || h.n > (1ull << 32) || h.k > B2P_MAX_COMMITMENTS)
            throw Error(B2P_ERR_ARG, "key snapshot: impossible sizes");
        const uint64_t col = h.n * 32;
        const uint64_t want = (5 + h.k) * col + pad64(3 * h.n * 8) + (h.k ? pad64((uint64_t)h.k * 8) : 0) + pad64(h.vkb_len);
        if (h.payload_len != want || m.len != sizeof h + want) throw Error(B2P_ERR_ARG, "key snapshot: length does not match its header");
        if (fnv1a_words(0xcbf29ce484222325ull, base + sizeof h, want) != h.fnv) throw Error(B2P_ERR_ARG, "key snapshot: checksum mismatch");
        const char* p = base + sizeof h;
        const void* cols[5];
        for (int i = 0; i < 5; i++) { cols[i] = p; p += col; }
        const int64_t* perm = reinterpret_cast<const int64_t*>(p);
        p += pad64(3 * h.n * 8);
        const void* qcp[B2P_MAX_COMMITMENTS] = {nullptr};
        for (uint32_t i = 0; i < h.k; i++) { qcp[i] = p; p += col; }
        const uint64_t* cidx = reinterpret_cast<const uint64_t*>(p);
        if (h.k) p += pad64((uint64_t)h.k * 8);
        const void* vkb = h.vkb_len ? p : nullptr;
        const int rc2 = b2p_circuit_load(srs, h.n, h.nb_public, cols[0], cols[1], cols[2], cols[3], cols[4], perm, h.k,
                                         h.k ? qcp : nullptr, h.k ? cidx : nullptr, vkb, h.vkb_len, out);
        if (rc2 != 0) throw Error(rc2, b2p_last_error());
    });
    return rc;
}

API int b2p_g2_generate_unsafe(int curve, const void* tau, void* out_g2) {
    return guarded([&] {
        require_curve(curve);
        require(tau && out_g2, "null argument");
        host_g2_unsafe(curve, tau, out_g2);
    });
}
