// Curve-erased interfaces between the C ABI (capi.cu) and the per-curve
// template instantiations (inst_prover_*.cu).  Keeping capi.cu free of the
// templates keeps its compile time in seconds.
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include "../../include/b200plonk.h"

namespace b2p {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct SrsBase {
    int curve = -1;
    int device = -1;      // CUDA device the handle lives on; every entry point switches to it
    // One stream, one MSM scratch area and one prover workspace per proving key: calls that use them are
    // serialised here ("re-entrant across handles, serialise calls on the same handle"); a circuit
    // handle locks the SRS handle it was loaded on.
    std::mutex mu;
    virtual ~SrsBase() {}
    virtual void load(const void* points, uint64_t n) = 0;
    virtual void load_compressed(const uint8_t* bytes, uint64_t n) = 0;
    virtual void generate_unsafe(const void* tau, uint64_t first, uint64_t stride, uint64_t n) = 0;
    virtual void get_points(uint64_t first, uint64_t count, void* out) const = 0;
    virtual uint64_t size() const = 0;
    virtual void to_lagrange(uint64_t n, void* out_points) = 0;
    virtual void msm_params(int* c, int* windows, uint64_t* buckets) const = 0;
    virtual void* stream_handle() = 0;
    virtual void set_commit_hook(int (*fn)(void*, const void*, uint64_t, void*), void* ctx) = 0;
    virtual void msm_g1(int basis, const void* scalars, uint64_t n, void* out_affine, bool device_scalars) = 0;
};

struct CircuitBase {
    int curve = -1;
    int device = -1;
    SrsBase* owner = nullptr;   // the SRS handle this circuit was loaded on (its lock serialises the calls)
    double stats[16] = {0};
    virtual ~CircuitBase() {}
    virtual void load(SrsBase* srs, uint64_t n, uint32_t nb_public, const void* ql, const void* qr, const void* qm,
                      const void* qo, const void* qk, const int64_t* perm, uint32_t k, const void* const* qcp,
                      const uint64_t* cidx, const void* vkb, uint64_t vkb_len) = 0;
    virtual void vk_commitments(void* out_points) = 0;
    virtual void prove(const void* L, const void* R, const void* O, const void* const* pi2, const void* bsb22,
                       const void* blinding, void* out_raw, bool device_inputs) = 0;
    virtual void* stream_handle() = 0;
    virtual void set_profiling(bool on) = 0;
    virtual void shard_buffers(void** out5, uint64_t* n) = 0;   // el er eo ez h (device), domain size
};

// One rank's part of an NTT whose domain is sharded over the GPUs of a box (ntt_shard.cuh).
struct NttShardBase {
    int curve = -1;
    int device = -1;
    virtual ~NttShardBase() {}
    virtual uint64_t local_size() const = 0;    // n / world
    virtual uint64_t chunk_size() const = 0;    // n / world^2
    virtual void forward_local(const void* d_coeffs, uint64_t local_len, int flags, void* d_x, void* stream) const = 0;
    virtual void forward_combine(const void* const* d_chunks, void* d_out, void* stream) const = 0;
    virtual void inverse_split(const void* d_evals, void* const* d_chunks, void* stream) const = 0;
    virtual void inverse_local(void* d_x, int flags, void* d_out, void* stream) const = 0;
};
NttShardBase* new_ntt_shard(int curve, uint64_t n, uint32_t world, uint32_t rank);   // inst_ntt.cu

// One rank of a shard group: the commitments (and, optionally, the big transforms) of one proof spread over the GPUs
// of a box (shard_group.cuh).  A rank shares SHARD_NPTR pieces of memory with its peers.
constexpr int SHARD_NPTR = 9;
struct ShardGroupBase {
    int curve = -1;
    int device = -1;
    uint32_t world = 1, rank = 0;
    uint64_t total = 0;                         // points of the whole SRS
    virtual ~ShardGroupBase() {}
    virtual void attach(SrsBase* prover_srs, CircuitBase* circuit) = 0;     // rank 0; nullptr detaches
    virtual void ipc_handles(void* out) const = 0;              // SHARD_NPTR x B2P_IPC_HANDLE_BYTES
    virtual void local_ptrs(void** out) const = 0;              // SHARD_NPTR pointers (ranks of one process)
    virtual void connect(const void* all_handles) = 0;          // world x SHARD_NPTR handles, rank-major
    virtual void connect_ptrs(void* const* all) = 0;            // world x SHARD_NPTR pointers, rank-major
    virtual void serve_proof(uint64_t n) = 0;                   // ranks > 0
    virtual void msm(const void* d_scalars, uint64_t n, void* out_affine) = 0;     // rank 0: one sharded commitment
    virtual void serve_msm(uint64_t n) = 0;                     // ranks > 0
};

// Level-parallel witness solver (solver.cuh, inst_solver.cu)
struct SolverBase {
    int curve = -1;
    int device = -1;
    std::mutex mu;
    virtual ~SolverBase() {}
    virtual void solve(const void* inputs, int where, void* L, void* R, void* O, bool device_out, void** dptrs) = 0;
    virtual void info(uint64_t* out8) const = 0;
    virtual void set_hint_fn(b2p_hint_fn fn, void* ctx) = 0;
};
SolverBase* new_solver(int curve, uint64_t n, uint32_t nb_public, uint64_t nb_variables, const uint32_t* input_ids,
                       uint32_t nb_inputs, const void* const cols[5], const uint32_t* xa, const uint32_t* xb,
                       const uint32_t* xc, const b2p_hint* hints, uint32_t n_hints, const uint8_t* unchecked);

struct CurveOps {
    virtual ~CurveOps() {}
    virtual ShardGroupBase* new_shard_group(uint32_t world, uint32_t rank, uint64_t total, SrsBase* shard,
                                            uint64_t ntt_rows) const = 0;
    virtual SrsBase* new_srs() const = 0;
    virtual CircuitBase* new_circuit() const = 0;
    virtual void ntt(void* data, uint64_t n, int flags) const = 0;
    virtual void g1_sum(const void* points, uint64_t n, void* out_affine) const = 0;
    virtual void marshal_proof(uint32_t k, const void* raw, const void* bsb22, uint8_t* out) const = 0;
    virtual void marshal_public_inputs(const void* values, uint32_t nb_public, uint8_t* out) const = 0;
};
const CurveOps* curve_ops_bn254();
const CurveOps* curve_ops_bls12381();

extern std::atomic<unsigned long long> g_launch_count;

// G2Affine.MultiExp, one shot (inst_msm_g2.cu): n G2Affine + n Fr (Montgomery) on the host -> one G2Affine
void msm_g2(int curve, const void* points, const void* scalars, uint64_t n, void* out);

// Host-side plonk.Verify (verify.cu): no device involved.
struct HostVerifyKey {
    uint64_t n;
    uint32_t nb_public, k;
    const uint64_t* commit_idx;   // k entries
    const void* vk_points;        // S1 S2 S3 Ql Qr Qm Qo Qk Qcp*, G1Affine memory
    const void* g1;               // Kzg.G1
    const void* g2;               // Kzg.G2[0], Kzg.G2[1], G2Affine memory
};
bool host_verify(int curve, const HostVerifyKey& vk, const void* proof, uint64_t proof_len, const void* pub,
                 uint64_t pub_len, std::string* why);
bool host_verify_batch(int curve, const HostVerifyKey& vk, const void* proofs, uint64_t proof_len, const void* pubs,
                       uint64_t pub_len, uint64_t count, uint64_t* first_bad, std::string* why);
// the same with the point combinations of the whole batch on the GPU (verify_batch.cuh, inst_verify_batch.cu)
bool device_verify_batch(int curve, const HostVerifyKey& vk, const void* proofs, uint64_t proof_len, const void* pubs,
                         uint64_t pub_len, uint64_t count, uint64_t* first_bad, std::string* why);
bool host_pairing_check(int curve, const void* g1s, const void* g2s, uint64_t n, std::string* why);
const char* host_kzg_vk_load(int curve, const void* vk_bin, uint64_t len, void* out_g2, void* out_g1);
void host_g2_unsafe(int curve, const void* tau_mont, void* out_two_g2);
// persisted keys (keyfile.hpp); nullptr = ok, else the reason
const char* host_gnark_file_parse(const void* file, uint64_t len, b2p_gnark_file* out);
const char* host_gnark_vk_parse(int curve, const void* bytes, uint64_t len, b2p_gnark_vk* out);
const char* host_gnark_pk_parse(int curve, const void* bytes, uint64_t len, b2p_gnark_pk* out);

}  // namespace b2p
