// One proof over the GPUs of a box, natively: the prover's kzg.Commit calls (9 per proof) sharded over
// the point set (BASELINE configs[2] "1->8 x B200 MSM shard over NVLink").  Replaces round 1's Python
// commit hook (a 32 n-byte NCCL broadcast + two host hops per commitment) with peer memory and flags:
//
//   rank 0 (runs b2p_prove)                               ranks g > 0 (b2p_shard_group_serve_proof)
//   ----------------------------------------------------  ------------------------------------------------------
//   commit(slot): D2D copy of the scalars into its          [kernels of the whole proof were queued in advance]
//     IPC-exported staging area, then ONE store per peer:   k_shard_wait  spins on ready[slot] in its own HBM
//     ready[slot] = proof number           --- NVLink -->   count / scatter kernels READ their 32 n/G bytes of the
//     Pippenger on its own block of the SRS                   scalars straight from rank 0's staging area (peer loads)
//                                                            Pippenger on their block of the SRS
//   fetch(first, cnt): reduction tails of its own slots     reduction tails, then k_shard_post: the partial sum (one
//     k_shard_wait_all on done[g][slot]     <-- NVLink ---    XYZZ point per slot) is STORED into rank 0's mailbox,
//     k_shard_sum: adds the G partial sums                    fence, done[g][slot] = proof number
//     D2H of cnt points, host inversion
//
// No collective library, no host hop between the local MSM and the exchange: the exchange is one 128/192-byte peer
// store and a flag per rank and slot.  The only host-level message is "a proof of n rows starts" (one per proof,
// sent by the host language; algoplonk_b200/shard_group.py), because the ranks queue the proof's fixed sequence of
// commitments (PROOF_COMMIT_SCHEDULE below -- the order Circuit::prove issues them in) ahead of time.
// Every wait has a timeout: a missing peer ends in an error code, not in a hung GPU.
#pragma once
#include "prover.cuh"

namespace b2p {

constexpr int SHARD_MAX_WORLD = 8;
constexpr uint64_t SHARD_MAIL_FLAG_BYTES = 4096;     // flags first, the partial sums behind them
constexpr unsigned long long SHARD_WAIT_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

struct ShardFlags {
    uint32_t ready[MSM_SLOTS];                        // on rank g > 0, written by rank 0: scalars of slot staged
    uint32_t done[SHARD_MAX_WORLD][MSM_SLOTS];        // on rank 0, written by rank g: partial sum of slot landed
    uint32_t error;                                   // a wait on this rank timed out
};
static_assert(sizeof(ShardFlags) <= SHARD_MAIL_FLAG_BYTES, "flag block too large");

// The commitments of one proof in the order Circuit::prove issues them: {extra scalars beyond n, result slot},
// a negative slot entry {first, -cnt} marks the fetch of `cnt` slots starting at `first`.
struct ShardStep { int a, b; };
static const ShardStep PROOF_COMMIT_SCHEDULE[] = {
    {2, 0}, {2, 1}, {2, 2}, {0, -3},          // [L] [R] [O], fetched together
    {3, 3}, {3, -1},                          // [Z]
    {2, 4}, {2, 5}, {2, 6}, {4, -3},          // [h0] [h1] [h2]
    {2, 8}, {8, -1},                          // W_{omega zeta}
    {2, 7}, {7, -1},                          // W_zeta
};

__device__ __forceinline__ unsigned long long shard_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool shard_reached(uint32_t seen, uint32_t want) { return (int32_t)(seen - want) >= 0; }

// one store per peer: *flags[g] = value, after everything this stream wrote before is visible system-wide
struct ShardPeerFlags { uint32_t* p[SHARD_MAX_WORLD]; };
static __global__ void k_shard_signal(ShardPeerFlags flags, int count, uint32_t value) {
    __threadfence_system();
    if ((int)threadIdx.x < count) *reinterpret_cast<volatile uint32_t*>(flags.p[threadIdx.x]) = value;
}
// thread t spins until flags[t * stride] reaches `value` (or the timeout sets *err)
static __global__ void k_shard_wait(const uint32_t* flags, int count, int stride, uint32_t value, uint32_t* err) {
    if ((int)threadIdx.x >= count) return;
    const volatile uint32_t* f = flags + (size_t)threadIdx.x * stride;
    const unsigned long long t0 = shard_now_ns();
    while (!shard_reached(*f, value)) {
        if (shard_now_ns() - t0 > SHARD_WAIT_TIMEOUT_NS) { *err = 1; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}
// rank g -> rank 0: partial sums of slots [first, first + cnt), then the flags
template <class Fp>
__global__ void k_shard_post(XYZZ<Fp>* __restrict__ remote_partials, const XYZZ<Fp>* __restrict__ local_results,
                             uint32_t* remote_done, int first, int cnt, uint32_t value) {
    const int t = threadIdx.x;
    if (t >= cnt) return;
    const uint4* src = reinterpret_cast<const uint4*>(local_results + first + t);
    uint4* dst = reinterpret_cast<uint4*>(remote_partials + first + t);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<Fp>) / 16); i++) dst[i] = src[i];
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t*>(remote_done + first + t) = value;
}
// rank 0: result[slot] += sum over the peers' partial sums
template <class Fp>
__global__ void k_shard_sum(XYZZ<Fp>* __restrict__ results, const XYZZ<Fp>* __restrict__ partials, int world, int first,
                            int cnt) {
    const int t = threadIdx.x;
    if (t >= cnt) return;
    XYZZ<Fp> acc = ld_xyzz_cg(results + first + t);
    for (int g = 1; g < world; g++) {
        const XYZZ<Fp> p = ld_xyzz_cg(partials + (size_t)g * MSM_SLOTS + first + t);   // written over NVLink: from L2
        acc.add(p);
    }
    st_xyzz(results + first + t, acc);
}

// [first, first + count) of rank `rank`: contiguous, balanced (sizes differ by at most one) -- the partition of
// algoplonk_b200/sharded.py:shard_range
inline void shard_block(uint64_t total, uint32_t rank, uint32_t world, uint64_t* first, uint64_t* count) {
    const uint64_t base = total / world, rem = total % world;
    *first = rank * base + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
}

template <class C>
struct ShardGroup : ShardGroupBase, CommitRouter {
    using Fr = typename C::Fr;
    using Fp = typename C::Fp;
    using Ext = XYZZ<Fp>;
    using Aff = Affine<Fp>;

    Srs<C>* shard = nullptr;              // this rank's block of the SRS (own table, own plan)
    Srs<C>* attached = nullptr;           // rank 0: the proving key whose commitments are routed here
    uint64_t first = 0, count = 0;        // this rank's block of the total points
    uint32_t proof_no = 0;
    bool connected = false, same_process = false;

    uint8_t* mail = nullptr;              // ShardFlags + partial sums, IPC-exported
    Fr* staging = nullptr;                // rank 0 only: MSM_SLOTS x total scalars, IPC-exported
    uint8_t* peer_mail[SHARD_MAX_WORLD] = {nullptr};
    Fr* peer_staging = nullptr;           // ranks > 0: rank 0's staging area, mapped
    uint32_t* h_err = nullptr;            // pinned

    ShardFlags* flags(uint8_t* m) const { return reinterpret_cast<ShardFlags*>(m); }
    Ext* partials(uint8_t* m) const { return reinterpret_cast<Ext*>(m + SHARD_MAIL_FLAG_BYTES); }
    static size_t mail_bytes() { return SHARD_MAIL_FLAG_BYTES + (size_t)SHARD_MAX_WORLD * MSM_SLOTS * sizeof(Ext); }
    // slots a proof uses: 0..8
    static constexpr int STAGE_SLOTS = 9;

    ShardGroup(uint32_t world_, uint32_t rank_, uint64_t total_, SrsBase* shard_) {
        curve = C::ID;
        world = world_; rank = rank_; total = total_;
        B2P_REQUIRE(world >= 1 && world <= SHARD_MAX_WORLD && rank < world, "shard group: bad rank / world (world <= 8)");
        B2P_REQUIRE(shard_ && shard_->curve == C::ID, "shard group: the SRS block is on another curve");
        shard = static_cast<Srs<C>*>(shard_);
        shard_block(total, rank, world, &first, &count);
        B2P_REQUIRE(shard->msm.npoints == count, "shard group: the SRS block does not hold this rank's share of the points");
        B2P_CUDA(cudaMalloc(&mail, mail_bytes()));
        B2P_CUDA(cudaMemset(mail, 0, mail_bytes()));
        if (rank == 0) B2P_CUDA(cudaMalloc(&staging, (size_t)STAGE_SLOTS * total * sizeof(Fr)));
        B2P_CUDA(cudaMallocHost(&h_err, sizeof(uint32_t)));
        *h_err = 0;
        // Load the four exchange kernels NOW (no-op launches).  CUDA loads a kernel at its first launch, and that
        // load can wait for kernels that are running -- such as a k_shard_wait spinning on a flag which only a
        // not-yet-loaded kernel would raise (seen with all ranks of a group in one process: a 20 s stall).
        {
            ShardPeerFlags none{};
            ShardFlags* f = flags(mail);
            B2P_LAUNCH(k_shard_signal, 1, 32, 0, 0, none, 0, 0u);
            B2P_LAUNCH(k_shard_wait, 1, 32, 0, 0, &f->ready[0], 0, 1, 0u, &f->error);
            B2P_LAUNCH((k_shard_post<Fp>), 1, 32, 0, 0, partials(mail), partials(mail), &f->done[0][0], 0, 0, 0u);
            B2P_LAUNCH((k_shard_sum<Fp>), 1, 32, 0, 0, partials(mail), partials(mail), 1, 0, 0);
        }
        B2P_CUDA(cudaDeviceSynchronize());
        peer_mail[rank] = mail;
    }
    ~ShardGroup() override {
        if (attached) attached->router = nullptr;
        if (connected && !same_process) {
            for (uint32_t g = 0; g < world; g++)
                if (g != rank && peer_mail[g]) cudaIpcCloseMemHandle(peer_mail[g]);
            if (peer_staging) cudaIpcCloseMemHandle(peer_staging);
        }
        if (mail) cudaFree(mail);
        if (staging) cudaFree(staging);
        if (h_err) cudaFreeHost(h_err);
    }

    void ipc_handles(void* out) const override {
        uint8_t* o = static_cast<uint8_t*>(out);
        memset(o, 0, 2 * B2P_IPC_HANDLE_BYTES);
        cudaIpcMemHandle_t h;
        B2P_CUDA(cudaIpcGetMemHandle(&h, mail));
        memcpy(o, &h, sizeof h);
        if (staging) {
            B2P_CUDA(cudaIpcGetMemHandle(&h, staging));
            memcpy(o + B2P_IPC_HANDLE_BYTES, &h, sizeof h);
        }
    }
    // all_handles: world x 2 IPC handles (mail, staging), rank-major.  Rank 0 maps every mailbox, the other ranks map
    // rank 0's mailbox and staging area.
    void connect(const void* all_handles) override {
        B2P_REQUIRE(!connected, "shard group: already connected");
        const uint8_t* hs = static_cast<const uint8_t*>(all_handles);
        auto open = [&](uint32_t g, int which) -> void* {
            cudaIpcMemHandle_t h;
            memcpy(&h, hs + ((size_t)g * 2 + which) * B2P_IPC_HANDLE_BYTES, sizeof h);
            void* p = nullptr;
            B2P_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            return p;
        };
        if (rank == 0) {
            for (uint32_t g = 1; g < world; g++) peer_mail[g] = static_cast<uint8_t*>(open(g, 0));
        } else {
            peer_mail[0] = static_cast<uint8_t*>(open(0, 0));
            peer_staging = static_cast<Fr*>(open(0, 1));
        }
        connected = true;
    }
    // the same wiring for ranks that live in ONE process on one device (tests: every rank a host thread)
    void connect_local(void* const* mails, void* staging0) override {
        B2P_REQUIRE(!connected, "shard group: already connected");
        for (uint32_t g = 0; g < world; g++)
            if (g != rank) peer_mail[g] = static_cast<uint8_t*>(mails[g]);
        if (rank != 0) peer_staging = static_cast<Fr*>(staging0);
        connected = same_process = true;
    }
    void* mail_ptr() const override { return mail; }
    void* staging_ptr() const override { return staging; }

    void attach(SrsBase* prover_srs) override {
        B2P_REQUIRE(rank == 0, "shard group: only rank 0 runs the prover");
        if (attached) attached->router = nullptr;
        attached = nullptr;
        if (!prover_srs) return;
        B2P_REQUIRE(connected || world == 1, "shard group: connect the ranks first");
        B2P_REQUIRE(prover_srs->curve == C::ID, "shard group: the proving key is on another curve");
        B2P_REQUIRE(prover_srs->device == shard->device, "shard group: the proving key lives on another device");
        attached = static_cast<Srs<C>*>(prover_srs);
        B2P_REQUIRE(attached->msm.npoints >= total, "shard group: the proving key's SRS is smaller than the sharded one");
        attached->router = this;
    }

    void slice(uint64_t n, uint64_t* lo, uint64_t* cnt) const {
        *lo = first < n ? first : n;
        const uint64_t hi = first + count < n ? first + count : n;
        *cnt = hi - *lo;
    }
    void check_err(const char* where, cudaStream_t st) {
        if (*h_err) {
            *h_err = 0;
            B2P_CUDA(cudaMemsetAsync(&flags(mail)->error, 0, sizeof(uint32_t), st));
            B2P_CUDA(cudaStreamSynchronize(st));
            throw Error(B2P_ERR_INTERNAL, std::string("shard group: timed out waiting for a peer (") + where + ")");
        }
    }

    // ---- rank 0: CommitRouter -------------------------------------------------------------------------------
    void begin_proof() override { proof_no++; }
    void commit(const void* d_scalars, uint64_t n, int slot, cudaStream_t st) override {
        B2P_REQUIRE(slot >= 0 && slot < STAGE_SLOTS, "shard group: result slot out of range");
        B2P_REQUIRE(n <= total, "shard group: more scalars than SRS points");
        const Fr* sc = static_cast<const Fr*>(d_scalars);
        if (world > 1) {
            Fr* stage = staging + (size_t)slot * total;
            if (n) B2P_CUDA(cudaMemcpyAsync(stage, sc, n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
            ShardPeerFlags pf;
            for (uint32_t g = 1; g < world; g++) pf.p[g - 1] = &flags(peer_mail[g])->ready[slot];
            B2P_LAUNCH(k_shard_signal, 1, 32, 0, st, pf, (int)world - 1, proof_no);
        }
        uint64_t lo, cnt;
        slice(n, &lo, &cnt);
        shard->msm.run_async(sc + lo, cnt, true, st, slot);
    }
    void fetch(int first_slot, int cnt, void* host_affine_out, cudaStream_t st) override {
        Ext h[MSM_SLOTS];
        B2P_REQUIRE(first_slot >= 0 && cnt >= 1 && first_slot + cnt <= STAGE_SLOTS, "shard group: result slot out of range");
        shard->msm.finish_async(first_slot, cnt, st);
        if (world > 1) {
            ShardFlags* f = flags(mail);
            for (uint32_t g = 1; g < world; g++)
                B2P_LAUNCH(k_shard_wait, 1, 32, 0, st, &f->done[g][first_slot], cnt, 1, proof_no, &f->error);
            B2P_LAUNCH((k_shard_sum<Fp>), 1, 32, 0, st, shard->msm.result.p, partials(mail), (int)world, first_slot, cnt);
            B2P_CUDA(cudaMemcpyAsync(h_err, &f->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        B2P_CUDA(cudaMemcpyAsync(h, shard->msm.result.p + first_slot, cnt * sizeof(Ext), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        check_err("partial sums of the other ranks", st);
        Aff* out = static_cast<Aff*>(host_affine_out);
        for (int i = 0; i < cnt; i++) out[i] = h[i].to_affine();
    }

    // ---- ranks > 0 ---------------------------------------------------------------------------------------------
    // Queues this rank's part of the 9 commitments of ONE proof of an n-row circuit and blocks until it is done.
    void serve_proof(uint64_t n) override {
        B2P_REQUIRE(rank != 0, "shard group: rank 0 proves, the other ranks serve");
        B2P_REQUIRE(connected, "shard group: connect the ranks first");
        B2P_REQUIRE(n + 3 <= total, "shard group: circuit too large for the sharded SRS");
        std::lock_guard<std::mutex> lk(shard->mu);
        cudaStream_t st = shard->stream;
        proof_no++;
        ShardFlags* mine = flags(mail);
        ShardFlags* root = flags(peer_mail[0]);
        for (const ShardStep& s : PROOF_COMMIT_SCHEDULE) {
            if (s.b >= 0) {
                const int slot = s.b;
                uint64_t lo, cnt;
                slice(n + s.a, &lo, &cnt);
                B2P_LAUNCH(k_shard_wait, 1, 32, 0, st, &mine->ready[slot], 1, 1, proof_no, &mine->error);
                shard->msm.run_async(peer_staging + (size_t)slot * total + lo, cnt, true, st, slot);
            } else {
                const int first_slot = s.a, cnt = -s.b;
                shard->msm.finish_async(first_slot, cnt, st);
                B2P_LAUNCH((k_shard_post<Fp>), 1, 32, 0, st, partials(peer_mail[0]) + (size_t)rank * MSM_SLOTS,
                           shard->msm.result.p, &root->done[rank][0], first_slot, cnt, proof_no);
            }
        }
        B2P_CUDA(cudaMemcpyAsync(h_err, &mine->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        check_err("the scalars of rank 0", st);
    }
};

template <class C>
ShardGroupBase* CurveOpsImpl<C>::new_shard_group(uint32_t world, uint32_t rank, uint64_t total, SrsBase* shard) const {
    return new ShardGroup<C>(world, rank, total, shard);
}

}  // namespace b2p
