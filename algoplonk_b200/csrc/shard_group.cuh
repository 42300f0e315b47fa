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
//
// With ntt_rows != 0 the five size-4n transforms of the proof (coset NTT of l, r, o, z; coset iNTT of the quotient)
// are spread over the same ranks as well (BASELINE configs[4] "NTT domain alltoall"; ntt_shard.cuh, world a power of
// two).  Per transform every rank: gathers its cyclic coefficients out of rank 0's staging area (peer loads, stride
// G), runs the local DIF passes into an exchange buffer that its peers have mapped, raises xready[rank] in every
// mailbox, waits for the others', and runs the combine kernel -- whose loads ARE the all-to-all (chunk `rank` of every
// peer's buffer over NVLink) and whose stores land directly in rank 0's evaluation buffer (el / er / eo / ez).  The
// inverse runs the mirror image on the quotient's evaluations and scatters the coefficients back into rank 0's h.
// The quotient kernel, the grand product and the openings stay on rank 0.
#pragma once
#include "prover.cuh"
#include "ntt_shard.cuh"
#include <cstdio>
#include <cstdlib>

namespace b2p {

constexpr int SHARD_MAX_WORLD = 8;
constexpr uint64_t SHARD_MAIL_FLAG_BYTES = 4096;     // flags first, the partial sums behind them
// how long a kernel spins on a flag before it gives up and raises the error flag (B2P_SHARD_TIMEOUT_MS overrides: tests)
inline unsigned long long shard_wait_timeout_ns() {
    const char* e = getenv("B2P_SHARD_TIMEOUT_MS");
    const long long ms = e ? atoll(e) : 0;
    return (ms > 0 ? (unsigned long long)ms : 20000ull) * 1000000ull;
}

struct ShardFlags {
    uint32_t ready[MSM_SLOTS];                        // on rank g > 0, written by rank 0: scalars of slot staged
    uint32_t done[SHARD_MAX_WORLD][MSM_SLOTS];        // on rank 0, written by rank g: partial sum of slot landed
    uint32_t error;                                   // a wait on this rank timed out
    uint32_t xready[SHARD_MAX_WORLD];                 // on every rank, written by rank g: its exchange buffer of transform # is complete
    uint32_t ntt_go;                                  // on rank g > 0, written by rank 0: the quotient of transform # is in rank 0's h
    uint32_t ntt_done[SHARD_MAX_WORLD];               // on rank 0, written by rank g: its part of transform # has landed on rank 0
};
static_assert(sizeof(ShardFlags) <= SHARD_MAIL_FLAG_BYTES, "flag block too large");

// The commitments of one proof in the order Circuit::prove issues them: {extra scalars beyond n, result slot},
// a negative slot entry {first, -cnt} marks the fetch of `cnt` slots starting at `first`.
// {SHARD_NTT, 0}: the proof's five big transforms happen here (only when the group shards them).
constexpr int SHARD_NTT = 1000;
struct ShardStep { int a, b; };
static const ShardStep PROOF_COMMIT_SCHEDULE[] = {
    {2, 0}, {2, 1}, {2, 2}, {0, -3},          // [L] [R] [O], fetched together
    {3, 3}, {3, -1},                          // [Z]
    {SHARD_NTT, 0},                           // l r o z -> 4n coset; quotient on rank 0; h back to coefficients
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
static __global__ void k_shard_wait(const uint32_t* flags, int count, int stride, uint32_t value, uint32_t* err,
                                    unsigned long long timeout_ns) {
    if ((int)threadIdx.x >= count) return;
    const volatile uint32_t* f = flags + (size_t)threadIdx.x * stride;
    const unsigned long long t0 = shard_now_ns();
    while (!shard_reached(*f, value)) {
        if (shard_now_ns() - t0 > timeout_ns) { *err = 1; break; }
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

// dst[j] = src[first + j stride] / dst[first + j stride] = src[j]: a rank's cyclic share of a vector that lives on rank 0
template <class Fr>
__global__ void k_shard_gather(Fr* __restrict__ dst, const Fr* __restrict__ src, uint64_t first, uint64_t stride, uint64_t cnt) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cnt) st_field(dst + j, ld_field(src + first + j * stride));
}
template <class Fr>
__global__ void k_shard_scatter(Fr* __restrict__ dst, const Fr* __restrict__ src, uint64_t first, uint64_t stride, uint64_t cnt) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cnt) st_field(dst + first + j * stride, ld_field(src + j));
}

// plain copy kernel: stands in for cudaMemcpyAsync when all ranks share one device (see ShardGroup::bulk_copy)
static __global__ void k_shard_copy(uint4* __restrict__ dst, const uint4* __restrict__ src, uint64_t n16) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// rank 0, after the inverse transform: h[j G + g] = stage[(g - 1) cap + j] for g = 1 .. G-1 (the other ranks' cyclic
// shares of the coefficients arrived as contiguous bulk copies; rank 0's own share is already in place)
template <class Fr>
__global__ void k_shard_interleave(Fr* __restrict__ h, const Fr* __restrict__ stage, uint32_t G, uint64_t cap, uint64_t out_len) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_len) return;
    const uint32_t g = (uint32_t)(idx % G);
    if (g) st_field(h + idx, ld_field(stage + (uint64_t)(g - 1) * cap + idx / G));
}

// [first, first + count) of rank `rank`: contiguous, balanced (sizes differ by at most one) -- the partition of
// algoplonk_b200/sharded.py:shard_range
inline void shard_block(uint64_t total, uint32_t rank, uint32_t world, uint64_t* first, uint64_t* count) {
    const uint64_t base = total / world, rem = total % world;
    *first = rank * base + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
}

// the memory a rank shares with its peers, in the order the handles / pointers are exchanged in
enum { SH_MAIL = 0, SH_STAGING, SH_XBUF, SH_EL, SH_ER, SH_EO, SH_EZ, SH_H, SH_HSTAGE, SH_NPTR };
static_assert(SH_NPTR == SHARD_NPTR, "iface.hpp and shard_group.cuh disagree on the shared pieces");

template <class C>
struct ShardGroup : ShardGroupBase, CommitRouter {
    using Fr = typename C::Fr;
    using Fp = typename C::Fp;
    using Ext = XYZZ<Fp>;
    using Aff = Affine<Fp>;

    Srs<C>* shard = nullptr;              // this rank's block of the SRS (own table, own plan)
    Srs<C>* attached = nullptr;           // rank 0: the proving key whose commitments are routed here
    uint64_t first = 0, count = 0;        // this rank's block of the total points
    uint32_t proof_no = 0, ntt_seq = 0;
    unsigned long long timeout_ns = shard_wait_timeout_ns();
    bool connected = false, same_process = false;

    uint8_t* mail = nullptr;              // ShardFlags + partial sums, shared
    Fr* staging = nullptr;                // rank 0 only: STAGE_SLOTS x total scalars, shared
    uint32_t* h_err = nullptr;            // pinned
    void* peer[SHARD_MAX_WORLD][SH_NPTR] = {{nullptr}};      // peer[g][what]: rank g's shared memory as seen from here

    // sharded transforms (ntt_rows != 0)
    uint64_t ntt_rows = 0;                // n of the circuits this group proves; the transforms have size 4n
    NttShard<Fr>* ntt = nullptr;
    Fr* xbuf = nullptr;                   // 2 exchange buffers of 4n/G elements, shared
    Fr* loc = nullptr;                    // this rank's cyclic coefficients before the local passes
    Fr* outbuf = nullptr;                 // ranks > 0: a block of evaluations on its way to / from rank 0 (bulk copies)
    Fr* hstage = nullptr;                 // rank 0: where the other ranks' shares of the quotient's coefficients land, shared
    uint64_t share_cap() const { return (3 * (ntt_rows + 2) + world - 1) / world; }
    void* circ_buf[5] = {nullptr};        // rank 0: the attached circuit's el er eo ez h

    // B2P_SHARD_TRACE=1: CUDA events around the steps of every sharded transform, printed per proof to stderr
    // (where the time of a transform goes on each rank: local passes / waiting for the peers / exchange kernel)
    struct Trace {
        bool on = false;
        std::vector<cudaEvent_t> ev;
        std::vector<const char*> what;
        void mark(const char* w, cudaStream_t st) {
            if (!on) return;
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            ev.push_back(e);
            what.push_back(w);
        }
        void dump(uint32_t rank) {
            if (!on || ev.empty()) return;
            cudaEventSynchronize(ev.back());
            std::string line = "[shard trace rank " + std::to_string(rank) + "]";
            for (size_t i = 1; i < ev.size(); i++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
                char buf[64];
                snprintf(buf, sizeof buf, " %s=%.3f", what[i], ms);
                line += buf;
            }
            fprintf(stderr, "%s\n", line.c_str());
            for (auto e : ev) cudaEventDestroy(e);
            ev.clear();
            what.clear();
        }
    } trace;
    ShardFlags* flags(void* m) const { return reinterpret_cast<ShardFlags*>(m); }
    Ext* partials(void* m) const { return reinterpret_cast<Ext*>(static_cast<uint8_t*>(m) + SHARD_MAIL_FLAG_BYTES); }
    static size_t mail_bytes() { return SHARD_MAIL_FLAG_BYTES + (size_t)SHARD_MAX_WORLD * MSM_SLOTS * sizeof(Ext); }
    static constexpr int STAGE_SLOTS = 9;         // slots a proof uses: 0..8

    ShardGroup(uint32_t world_, uint32_t rank_, uint64_t total_, SrsBase* shard_, uint64_t ntt_rows_) {
        curve = C::ID;
        world = world_; rank = rank_; total = total_; ntt_rows = ntt_rows_;
        { const char* e = getenv("B2P_SHARD_TRACE"); trace.on = e && atoi(e) != 0; }
        B2P_REQUIRE(world >= 1 && world <= SHARD_MAX_WORLD && rank < world, "shard group: bad rank / world (world <= 8)");
        B2P_REQUIRE(shard_ && shard_->curve == C::ID, "shard group: the SRS block is on another curve");
        shard = static_cast<Srs<C>*>(shard_);
        shard_block(total, rank, world, &first, &count);
        B2P_REQUIRE(shard->msm.npoints == count, "shard group: the SRS block does not hold this rank's share of the points");
        if (ntt_rows) {
            B2P_REQUIRE((world & (world - 1)) == 0, "shard group: sharded transforms need a power-of-two world");
            B2P_REQUIRE(ntt_rows >= 64 && (ntt_rows & (ntt_rows - 1)) == 0 && ntt_rows + 3 <= total,
                        "shard group: ntt_rows must be a power of two >= 64 with ntt_rows + 3 <= total points");
        }
        try {
            B2P_CUDA(cudaMalloc(&mail, mail_bytes()));
            B2P_CUDA(cudaMemset(mail, 0, mail_bytes()));
            if (rank == 0) B2P_CUDA(cudaMalloc(&staging, (size_t)STAGE_SLOTS * total * sizeof(Fr)));
            B2P_CUDA(cudaMallocHost(&h_err, sizeof(uint32_t)));
            *h_err = 0;
            if (ntt_rows) {
                ntt = new NttShard<Fr>();
                ntt->init(4 * ntt_rows, world, rank);
                B2P_CUDA(cudaMalloc(&xbuf, 2 * ntt->local_n * sizeof(Fr)));
                B2P_CUDA(cudaMalloc(&loc, ntt->local_n * sizeof(Fr)));
                if (rank) B2P_CUDA(cudaMalloc(&outbuf, ntt->local_n * sizeof(Fr)));
                else if (world > 1) B2P_CUDA(cudaMalloc(&hstage, (size_t)(world - 1) * share_cap() * sizeof(Fr)));
            }
            // Load the exchange kernels NOW (no-op launches).  CUDA loads a kernel at its first launch, and that load
            // can wait for kernels that are running -- such as a k_shard_wait spinning on a flag which only a
            // not-yet-loaded kernel would raise (seen with all ranks of a group in one process: a 20 s stall).
            ShardPeerFlags none{};
            ShardFlags* f = flags(mail);
            B2P_LAUNCH(k_shard_signal, 1, 32, 0, 0, none, 0, 0u);
            B2P_LAUNCH(k_shard_wait, 1, 32, 0, 0, &f->ready[0], 0, 1, 0u, &f->error, timeout_ns);
            B2P_LAUNCH((k_shard_post<Fp>), 1, 32, 0, 0, partials(mail), partials(mail), &f->done[0][0], 0, 0, 0u);
            B2P_LAUNCH((k_shard_sum<Fp>), 1, 32, 0, 0, partials(mail), partials(mail), 1, 0, 0);
            B2P_LAUNCH((k_shard_gather<Fr>), 1, 32, 0, 0, (Fr*)nullptr, (const Fr*)nullptr, 0, 1, 0);
            B2P_LAUNCH((k_shard_scatter<Fr>), 1, 32, 0, 0, (Fr*)nullptr, (const Fr*)nullptr, 0, 1, 0);
            B2P_LAUNCH((k_shard_interleave<Fr>), 1, 32, 0, 0, (Fr*)nullptr, (const Fr*)nullptr, 1u, 0, 0);
            B2P_LAUNCH(k_shard_copy, 1, 32, 0, 0, (uint4*)nullptr, (const uint4*)nullptr, (uint64_t)0);
            B2P_CUDA(cudaDeviceSynchronize());
        } catch (...) {
            release();
            throw;
        }
        peer[rank][SH_MAIL] = mail;
        peer[rank][SH_STAGING] = staging;
        peer[rank][SH_XBUF] = xbuf;
        peer[rank][SH_HSTAGE] = hstage;
    }
    void release() {
        if (mail) cudaFree(mail);
        if (staging) cudaFree(staging);
        if (xbuf) cudaFree(xbuf);
        if (loc) cudaFree(loc);
        if (outbuf) cudaFree(outbuf);
        if (hstage) cudaFree(hstage);
        if (h_err) cudaFreeHost(h_err);
        delete ntt;
        mail = nullptr; staging = nullptr; xbuf = nullptr; loc = nullptr; outbuf = nullptr; hstage = nullptr;
        h_err = nullptr; ntt = nullptr;
    }
    ~ShardGroup() override {
        if (attached) attached->router = nullptr;
        if (connected && !same_process)
            for (uint32_t g = 0; g < world; g++)
                for (int w = 0; w < SH_NPTR; w++)
                    if (g != rank && peer[g][w]) cudaIpcCloseMemHandle(peer[g][w]);
        release();
    }

    // ---- wiring --------------------------------------------------------------------------------------------------
    void attach(SrsBase* prover_srs, CircuitBase* circuit) override {
        B2P_REQUIRE(rank == 0, "shard group: only rank 0 runs the prover");
        if (attached) attached->router = nullptr;
        attached = nullptr;
        if (!prover_srs) return;
        B2P_REQUIRE(prover_srs->curve == C::ID, "shard group: the proving key is on another curve");
        B2P_REQUIRE(prover_srs->device == shard->device, "shard group: the proving key lives on another device");
        Srs<C>* s = static_cast<Srs<C>*>(prover_srs);
        B2P_REQUIRE(s->msm.npoints >= total, "shard group: the proving key's SRS is smaller than the sharded one");
        if (ntt_rows) {
            B2P_REQUIRE(circuit && circuit->owner == prover_srs, "shard group: sharded transforms need the circuit handle");
            uint64_t n = 0;
            void* bufs[5];
            circuit->shard_buffers(bufs, &n);
            B2P_REQUIRE(n == ntt_rows, "shard group: the circuit's domain differs from the group's ntt_rows");
            for (int i = 0; i < 5; i++) {
                // once the ranks are connected they have mapped the first circuit's buffers: only that one re-attaches
                B2P_REQUIRE(!connected || peer[0][SH_EL + i] == bufs[i],
                            "shard group: attach the circuit before the ranks are connected (its buffers are shared)");
                peer[0][SH_EL + i] = circ_buf[i] = bufs[i];
            }
        }
        attached = s;
        attached->router = this;
    }
    // this rank's shared memory: SH_NPTR pointers (same process) or IPC handles (zero where it has none)
    void local_ptrs(void** out) const override {
        for (int w = 0; w < SH_NPTR; w++) out[w] = peer[rank][w];
    }
    void ipc_handles(void* out) const override {
        uint8_t* o = static_cast<uint8_t*>(out);
        memset(o, 0, (size_t)SH_NPTR * B2P_IPC_HANDLE_BYTES);
        for (int w = 0; w < SH_NPTR; w++) {
            if (!peer[rank][w]) continue;
            cudaIpcMemHandle_t h;
            B2P_CUDA(cudaIpcGetMemHandle(&h, peer[rank][w]));
            memcpy(o + (size_t)w * B2P_IPC_HANDLE_BYTES, &h, sizeof h);
        }
    }
    // what this rank needs of rank g: rank 0 reads every mailbox; everybody needs rank 0's mailbox, staging area and
    // (sharded transforms) its evaluation buffers; with sharded transforms every rank needs every mailbox and buffer
    bool needs(uint32_t g, int w) const {
        if (g == rank) return false;
        if (w == SH_MAIL) return rank == 0 || g == 0 || ntt_rows;
        if (w == SH_XBUF) return ntt_rows != 0;
        if (w == SH_STAGING) return g == 0;
        return g == 0 && ntt_rows != 0;
    }
    void connect(const void* all_handles) override {
        B2P_REQUIRE(!connected, "shard group: already connected");
        const uint8_t* hs = static_cast<const uint8_t*>(all_handles);
        for (uint32_t g = 0; g < world; g++)
            for (int w = 0; w < SH_NPTR; w++) {
                if (!needs(g, w)) continue;
                cudaIpcMemHandle_t h;
                memcpy(&h, hs + ((size_t)g * SH_NPTR + w) * B2P_IPC_HANDLE_BYTES, sizeof h);
                B2P_CUDA(cudaIpcOpenMemHandle(&peer[g][w], h, cudaIpcMemLazyEnablePeerAccess));
            }
        connected = true;
    }
    void connect_ptrs(void* const* all) override {      // world x SH_NPTR pointers, ranks of one process
        B2P_REQUIRE(!connected, "shard group: already connected");
        for (uint32_t g = 0; g < world; g++)
            for (int w = 0; w < SH_NPTR; w++)
                if (needs(g, w)) peer[g][w] = all[(size_t)g * SH_NPTR + w];
        connected = same_process = true;
    }

    // Device-to-device copy on `st`.  Between GPUs: the copy engine (one bulk transfer over NVLink).  With every rank
    // in ONE process on one device (tests) the ranks' copies would share the device's copy-engine queues, where a copy
    // whose stream is still waiting on a flag blocks the copies queued behind it -- among them, possibly, the one that
    // would let the flag be raised; a copy kernel has no such queue.
    void bulk_copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
        if (!bytes) return;
        if (same_process) {
            B2P_LAUNCH(k_shard_copy, 592, 256, 0, st, static_cast<uint4*>(dst), static_cast<const uint4*>(src), (uint64_t)(bytes / 16));
        } else {
            B2P_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
        }
    }
    // Spin on `count` flags (stride apart) until they reach `value`.  With every rank in ONE process on one device
    // (tests) the host then waits for the kernel: nothing -- in particular no copy-engine operation, whose queues are
    // FIFO across streams -- is ever queued behind a kernel that is still spinning, so a blocked queue cannot sit in
    // front of the operation that would let the flag be raised.  One process per GPU: fully asynchronous.
    void wait_flags(const uint32_t* f, int count, int stride, uint32_t value, cudaStream_t st) {
        B2P_LAUNCH(k_shard_wait, 1, 32, 0, st, f, count, stride, value, &flags(mail)->error, timeout_ns);
        if (same_process) B2P_CUDA(cudaStreamSynchronize(st));
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
    Fr* stage_of(uint32_t g0_view_slot) const { return static_cast<Fr*>(peer[0][SH_STAGING]) + (size_t)g0_view_slot * total; }

    // ---- rank 0: CommitRouter -------------------------------------------------------------------------------
    void begin_proof() override {
        B2P_REQUIRE(connected || world == 1, "shard group: connect the ranks first");
        proof_no++;
    }
    void commit(const void* d_scalars, uint64_t n, int slot, cudaStream_t st) override {
        B2P_REQUIRE(slot >= 0 && slot < STAGE_SLOTS, "shard group: result slot out of range");
        B2P_REQUIRE(n <= total, "shard group: more scalars than SRS points");
        const Fr* sc = static_cast<const Fr*>(d_scalars);
        if (world > 1) {
            Fr* stage = staging + (size_t)slot * total;
            bulk_copy(stage, sc, n * sizeof(Fr), st);
            ShardPeerFlags pf;
            for (uint32_t g = 1; g < world; g++) pf.p[g - 1] = &flags(peer[g][SH_MAIL])->ready[slot];
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
                wait_flags(&f->done[g][first_slot], cnt, 1, proof_no, st);
            B2P_LAUNCH((k_shard_sum<Fp>), 1, 32, 0, st, shard->msm.result.p, partials(mail), (int)world, first_slot, cnt);
            B2P_CUDA(cudaMemcpyAsync(h_err, &f->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        B2P_CUDA(cudaMemcpyAsync(h, shard->msm.result.p + first_slot, cnt * sizeof(Ext), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        if (first_slot == 7) trace.dump(rank);           // the last fetch of a proof
        check_err("partial sums of the other ranks", st);
        Aff* out = static_cast<Aff*>(host_affine_out);
        for (int i = 0; i < cnt; i++) out[i] = h[i].to_affine();
    }

    // ---- the sharded transforms: the same code on every rank ---------------------------------------------------
    bool shards_ntt() const override { return ntt_rows != 0 && world > 1; }
    void xready_barrier(uint32_t seq, cudaStream_t st) {
        ShardPeerFlags pf;
        for (uint32_t g = 0; g < world; g++) pf.p[g] = &flags(peer[g][SH_MAIL])->xready[rank];
        B2P_LAUNCH(k_shard_signal, 1, 32, 0, st, pf, (int)world, seq);
        wait_flags(&flags(mail)->xready[0], (int)world, 1, seq, st);
    }
    void tell_rank0_done(uint32_t seq, cudaStream_t st) {
        if (rank == 0) return;
        ShardPeerFlags pf;
        pf.p[0] = &flags(peer[0][SH_MAIL])->ntt_done[rank];
        B2P_LAUNCH(k_shard_signal, 1, 32, 0, st, pf, 1, seq);
    }
    // coefficients (len of them, on rank 0: staging slot or `coeffs`) -> 4n coset evaluations in rank 0's buffer `which`
    void forward_step(const Fr* coeffs_as_seen_here, uint64_t len, int which, cudaStream_t st) {
        const uint32_t seq = ++ntt_seq, G = world;
        Fr* x = xbuf + (size_t)(seq & 1) * ntt->local_n;
        const uint64_t cnt = len > rank ? (len - rank + G - 1) / G : 0;
        trace.mark("fwd", st);
        if (cnt) B2P_LAUNCH((k_shard_gather<Fr>), div_up(cnt, 256), 256, 0, st, loc, coeffs_as_seen_here, (uint64_t)rank, (uint64_t)G, cnt);
        trace.mark("gather", st);
        ntt->forward_local(loc, cnt, B2P_NTT_COSET, x, st);
        trace.mark("local", st);
        xready_barrier(seq, st);
        trace.mark("barrier", st);
        const void* chunks[1 << NTT_SHARD_MAX_LOGG] = {nullptr};
        for (uint32_t g = 0; g < G; g++)
            chunks[g] = static_cast<Fr*>(peer[g][SH_XBUF]) + (size_t)(seq & 1) * ntt->local_n + (size_t)rank * ntt->chunk_len;
        // the exchange (peer LOADS of the chunks) writes this rank's block of the evaluations locally; from the other
        // ranks it then travels to rank 0's buffer as ONE bulk copy (copy engine over NVLink: ~3x the bandwidth the
        // combine kernel reached when its stores went to the peer directly, profiles/shard_trace_r2_*.txt)
        Fr* dst0 = static_cast<Fr*>(peer[0][which]) + (size_t)rank * ntt->local_n;
        ntt->forward_combine(chunks, rank ? outbuf : dst0, st);
        trace.mark("combine", st);
        if (rank) {
            bulk_copy(dst0, outbuf, ntt->local_n * sizeof(Fr), st);
            trace.mark("push", st);
        }
        tell_rank0_done(seq, st);
    }
    // 4n evaluations in rank 0's h -> the first out_len coefficients, back in rank 0's h
    void inverse_step(uint64_t out_len, cudaStream_t st) {
        const uint32_t seq = ++ntt_seq, G = world;
        Fr* x = xbuf + (size_t)(seq & 1) * ntt->local_n;
        Fr* h0 = static_cast<Fr*>(peer[0][SH_H]);
        if (rank == 0) {            // the quotient is complete (stream order): the other ranks may read it
            ShardPeerFlags pf;
            for (uint32_t g = 1; g < G; g++) pf.p[g - 1] = &flags(peer[g][SH_MAIL])->ntt_go;
            B2P_LAUNCH(k_shard_signal, 1, 32, 0, st, pf, (int)G - 1, seq);
        } else {
            wait_flags(&flags(mail)->ntt_go, 1, 1, seq, st);
        }
        void* chunks[1 << NTT_SHARD_MAX_LOGG] = {nullptr};
        for (uint32_t g = 0; g < G; g++)
            chunks[g] = static_cast<Fr*>(peer[g][SH_XBUF]) + (size_t)(seq & 1) * ntt->local_n + (size_t)rank * ntt->chunk_len;
        trace.mark("inv-go", st);
        const Fr* block = h0 + (size_t)rank * ntt->local_n;
        if (rank) {                 // pull this rank's block of the quotient's evaluations with one bulk copy
            bulk_copy(outbuf, block, ntt->local_n * sizeof(Fr), st);
            block = outbuf;
            trace.mark("pull", st);
        }
        ntt->inverse_split(block, chunks, st);
        trace.mark("split", st);
        xready_barrier(seq, st);     // every rank has read its block of h and delivered its chunks
        trace.mark("barrier", st);
        ntt->inverse_local(x, B2P_NTT_INVERSE | B2P_NTT_COSET, nullptr, st);
        trace.mark("local", st);
        const uint64_t cnt = out_len > rank ? (out_len - rank + G - 1) / G : 0;
        B2P_REQUIRE(cnt <= share_cap(), "shard group: more quotient coefficients than the staging area holds");
        if (rank == 0) {            // own share: in place (local strided stores)
            if (cnt) B2P_LAUNCH((k_shard_scatter<Fr>), div_up(cnt, 256), 256, 0, st, h0, x, (uint64_t)0, (uint64_t)G, cnt);
        } else if (cnt) {           // the others': one contiguous bulk copy into rank 0's staging area
            Fr* slot = static_cast<Fr*>(peer[0][SH_HSTAGE]) + (size_t)(rank - 1) * share_cap();
            bulk_copy(slot, x, cnt * sizeof(Fr), st);
        }
        trace.mark("scatter", st);
        tell_rank0_done(seq, st);
    }
    void wait_ntt_done(cudaStream_t st) {      // rank 0: every rank's part of transform ntt_seq has landed here
        ShardFlags* f = flags(mail);
        if (world > 1) wait_flags(&f->ntt_done[1], (int)world - 1, 1, ntt_seq, st);
        trace.mark("wait-done", st);
    }
    // CommitRouter (rank 0)
    void ntt_forward(const void* d_coeffs, uint64_t len, int which, cudaStream_t st) override {
        forward_step(static_cast<const Fr*>(d_coeffs), len, SH_EL + which, st);
    }
    void ntt_forward_wait(cudaStream_t st) override { wait_ntt_done(st); }
    void ntt_inverse(uint64_t out_len, cudaStream_t st) override {
        inverse_step(out_len, st);
        wait_ntt_done(st);
        if (world > 1)
            B2P_LAUNCH((k_shard_interleave<Fr>), div_up(out_len, 256), 256, 0, st, static_cast<Fr*>(peer[0][SH_H]), hstage,
                       world, share_cap(), out_len);
        trace.mark("interleave", st);
    }

    // ---- one stand-alone commitment over the group (kzg.Commit / G1Affine.MultiExp, multi-GPU form) ----------------
    // rank 0: n device-resident scalars (Montgomery) -> the affine commitment on the host
    void msm(const void* d_scalars, uint64_t n, void* out_affine) override {
        B2P_REQUIRE(rank == 0, "shard group: rank 0 commits, the other ranks serve");
        B2P_REQUIRE(!attached, "shard group: detach the proving key before stand-alone commitments (one sequence of proof numbers)");
        std::lock_guard<std::mutex> lk(shard->mu);
        begin_proof();
        commit(d_scalars, n, 0, shard->stream);
        fetch(0, 1, out_affine, shard->stream);
    }
    void serve_msm(uint64_t n) override {
        B2P_REQUIRE(rank != 0 && connected, "shard group: the other ranks serve, after connect");
        B2P_REQUIRE(n <= total, "shard group: more scalars than SRS points");
        std::lock_guard<std::mutex> lk(shard->mu);
        cudaStream_t st = shard->stream;
        proof_no++;
        ShardFlags* mine = flags(mail);
        uint64_t lo, cnt;
        slice(n, &lo, &cnt);
        wait_flags(&mine->ready[0], 1, 1, proof_no, st);
        shard->msm.run_async(stage_of(0) + lo, cnt, true, st, 0);
        shard->msm.finish_async(0, 1, st);
        B2P_LAUNCH((k_shard_post<Fp>), 1, 32, 0, st, partials(peer[0][SH_MAIL]) + (size_t)rank * MSM_SLOTS,
                   shard->msm.result.p, &flags(peer[0][SH_MAIL])->done[rank][0], 0, 1, proof_no);
        B2P_CUDA(cudaMemcpyAsync(h_err, &mine->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        check_err("rank 0", st);
    }

    // ---- ranks > 0 ---------------------------------------------------------------------------------------------
    // Queues this rank's part of ONE proof of an n-row circuit and blocks until it is done.
    void serve_proof(uint64_t n) override {
        B2P_REQUIRE(rank != 0, "shard group: rank 0 proves, the other ranks serve");
        B2P_REQUIRE(connected, "shard group: connect the ranks first");
        B2P_REQUIRE(n + 3 <= total, "shard group: circuit too large for the sharded SRS");
        B2P_REQUIRE(!ntt_rows || n == ntt_rows, "shard group: this group shards the transforms of another circuit size");
        std::lock_guard<std::mutex> lk(shard->mu);
        cudaStream_t st = shard->stream;
        proof_no++;
        ShardFlags* mine = flags(mail);
        ShardFlags* root = flags(peer[0][SH_MAIL]);
        for (const ShardStep& s : PROOF_COMMIT_SCHEDULE) {
            if (s.a == SHARD_NTT) {
                if (!shards_ntt()) continue;
                for (int w = 0; w < 4; w++) forward_step(stage_of(w), n + (w == 3 ? 3 : 2), SH_EL + w, st);
                inverse_step(3 * (n + 2), st);
            } else if (s.b >= 0) {
                const int slot = s.b;
                uint64_t lo, cnt;
                slice(n + s.a, &lo, &cnt);
                wait_flags(&mine->ready[slot], 1, 1, proof_no, st);
                shard->msm.run_async(stage_of(slot) + lo, cnt, true, st, slot);
            } else {
                const int first_slot = s.a, cnt = -s.b;
                shard->msm.finish_async(first_slot, cnt, st);
                B2P_LAUNCH((k_shard_post<Fp>), 1, 32, 0, st, partials(peer[0][SH_MAIL]) + (size_t)rank * MSM_SLOTS,
                           shard->msm.result.p, &root->done[rank][0], first_slot, cnt, proof_no);
            }
        }
        B2P_CUDA(cudaMemcpyAsync(h_err, &mine->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        trace.dump(rank);
        check_err("rank 0 or a peer", st);
    }
};

template <class C>
ShardGroupBase* CurveOpsImpl<C>::new_shard_group(uint32_t world, uint32_t rank, uint64_t total, SrsBase* shard,
                                                 uint64_t ntt_rows) const {
    return new ShardGroup<C>(world, rank, total, shard, ntt_rows);
}

}  // namespace b2p
