/* b200plonk.h -- C ABI of the B200-native PLONK prover.
 *
 * This is the drop-in boundary for the one hot call AlgoPlonk makes:
 *
 *     proof, err := plonk.Prove(cc.Ccs, cc.Pk, witness)     reference algoplonk.go:89
 *                                                           (and testutils/testutils.go:47)
 *
 * The reference has no FFI of its own (pure Go over gnark v0.15.0 / gnark-crypto
 * v0.20.1, go.mod:8-9); the entry points below are what a cgo shim with the
 * signature of plonk.Prove binds (go/gpuplonk/prove.go).  Every
 * buffer uses gnark-crypto's in-memory layout so the shim passes
 * unsafe.Pointer(&slice[0]) with no conversion:
 *
 *   Fr        32 bytes : 4 x u64 little-endian limbs, Montgomery form (R = 2^256)
 *   G1Affine  BN254 64 bytes / BLS12-381 96 bytes : X || Y, each Fp in Montgomery
 *             form (R = 2^256 / 2^384), little-endian limbs; (0,0) = infinity.
 *
 * All functions return 0 on success and a negative code on failure; the message
 * is available from b2p_last_error() (thread local).  Nothing aborts, nothing
 * throws across the boundary, no caller pointer is retained after a call returns
 * (cgo rule).  Handles may be used from any thread; a handle is bound to the CUDA device that was current when
 * it was created and every entry point switches to that device for the duration of the call (goroutines migrate
 * between OS threads).  A proving key (an SRS handle and the circuit handles loaded on it) owns one stream and one
 * workspace: concurrent b2p_prove / b2p_msm_g1 / b2p_circuit_load calls on the same key are serialised by the
 * library (a lock per SRS handle), calls on different keys run concurrently -- several proofs can be in flight on
 * one GPU, one proving key each.  Freeing a handle while another thread uses it is the caller's error.
 *
 * The verification entry points (b2p_verify, b2p_verify_batch, b2p_pairing_check, b2p_kzg_vk_load,
 * b2p_g2_generate_unsafe: plonk.Verify, algoplonk.go:93) take no handle and touch no device: host arithmetic, as in
 * gnark, re-entrant from any number of threads.  G2Affine = X.A0 X.A1 Y.A0 Y.A1, each an Fp as above.
 */
#ifndef B200PLONK_H
#define B200PLONK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2P_BN254      0   /* ecc.BN254      (algoplonk.go:39) */
#define B2P_BLS12_381  1   /* ecc.BLS12_381  (algoplonk.go:39) */

#define B2P_OK             0
#define B2P_ERR_ARG       -1   /* bad argument / unsupported size */
#define B2P_ERR_CUDA      -2   /* CUDA runtime failure (no device, OOM, launch error) */
#define B2P_ERR_INTERNAL  -3
#define B2P_ERR_VERIFY    -4   /* b2p_verify: the proof was rejected (plonk.Verify's error) */

#define B2P_BASIS_CANONICAL 0  /* pk.Kzg          : [tau^j]_1            */
#define B2P_BASIS_LAGRANGE  1  /* pk.KzgLagrange  : [L_j(tau)]_1, size n */

typedef struct b2p_srs b2p_srs;
typedef struct b2p_circuit b2p_circuit;

/* ---- process -------------------------------------------------------------- */

/* Selects the CUDA device for the calling thread's subsequent handle creation
 * (-1 keeps the current device).  Fails with B2P_ERR_CUDA when no GPU is usable:
 * there is no CPU fallback inside this library. */
int b2p_init(int device);
const char* b2p_last_error(void);
const char* b2p_version(void);
/* Kernel launches issued by this library since process start (for bench.py's gpu_launches). */
uint64_t b2p_launch_count(void);

/* Page-locked host memory for the wire columns.  b2p_prove accepts any host pointer, but only pinned buffers
 * upload at PCIe speed and overlap with compute (pageable memory is staged by the driver at a fraction of
 * that); the Go shim copies gnark's solver output into buffers from here instead of into fresh slices. */
int b2p_host_alloc(uint64_t bytes, void** out);
void b2p_host_free(void* p);

/* ---- SRS  (replaces: kzg.SRS as loaded by setup/setup.go:165-193) ---------- */

/* g1_canonical: n_can points [tau^j]_1 (what srs.Pk.ReadFrom produced, setup.go:173,189).
 * g1_lagrange : optional (may be NULL, n_lag = 0).  Lagrange-basis commitments are
 * computed as iNTT + canonical MSM, which is the same group element as
 * MSM(kzg.ToLagrangeG1(...)) (setup.go:124,138); the argument exists so the shim
 * can pass the whole gnark ProvingKey.  Builds the windowed point table in HBM. */
int b2p_srs_load(int curve, const void* g1_canonical, uint64_t n_can,
                 const void* g1_lagrange, uint64_t n_lag, b2p_srs** out);

/* The embedded trusted setups as they sit on disk (setup/<name>/pk.bin, gnark kzg.ProvingKey.WriteTo):
 * u32 big-endian point count, then compressed G1 points (BN254 32 B with 2 flag bits, BLS12-381 48 B with
 * 3 flag bits).  Replaces loadTrustedSetupBytes + srs.Pk.ReadFrom (setup.go:165-228) for the first `count`
 * points: the square roots of the decompression run on the GPU.  Fails with "pk.bin too small for N
 * elements" like setup.go:219-223.  gnark's decoder also checks subgroup membership of BLS12-381 points;
 * this loader, like gnark's UnsafeReadFrom, does not (the files are audited ceremony outputs). */
int b2p_srs_load_compressed(int curve, const void* pk_bin, uint64_t len, uint64_t count, b2p_srs** out);

/* unsafekzg.NewSRS (setup.go:102-108, the TestOnly setups): [tau^j]_1 for j < n_can,
 * generated on the device from a known tau (Fr, Montgomery form). */
int b2p_srs_generate_unsafe(int curve, const void* tau, uint64_t n_can, b2p_srs** out);

/* One rank's shard of the same SRS: [tau^(first+j)]_1 for j < count (multi-GPU MSM, the point set is
 * split across the GPUs of a box). */
int b2p_srs_generate_unsafe_range(int curve, const void* tau, uint64_t first, uint64_t count, b2p_srs** out);

/* A cyclic shard: [tau^(first + j*stride)]_1 for j < count.  With first = rank and stride = world this is the
 * SRS shard that matches the cyclic coefficient distribution of the domain-sharded NTT (b2p_ntt_shard_*), so a
 * polynomial that leaves a sharded inverse transform is committed with no redistribution. */
int b2p_srs_generate_unsafe_strided(int curve, const void* tau, uint64_t first, uint64_t stride, uint64_t count,
                                    b2p_srs** out);

/* kzg.ToLagrangeG1(srs.Pk.G1[:n]) (setup.go:124,138: pk.KzgLagrange of gnark's ProvingKey): the n Lagrange-basis
 * points [L_j(tau)]_1 = (1/n) sum_i omega^(-ij) [tau^i]_1, computed on the GPU as an inverse transform over the group
 * (XYZZ points, one launch per stage, point-by-twiddle multiplications in the butterflies) and written to the host as n
 * G1Affine.  n a power of two <= the SRS size.  b2p_prove does not need it (see b2p_srs_load); it is for callers that
 * keep gnark's own ProvingKey complete, e.g. the shim's fallback to plonk.Prove. */
int b2p_srs_to_lagrange(b2p_srs* srs, uint64_t n, void* out_points);

/* Copies canonical points [first, first+count) back to the host (G1Affine layout). */
int b2p_srs_get_points(const b2p_srs* srs, uint64_t first, uint64_t count, void* out);
uint64_t b2p_srs_size(const b2p_srs* srs);
/* window bits / number of windows / buckets chosen for this SRS (reporting) */
int b2p_srs_msm_params(const b2p_srs* srs, int* c, int* windows, uint64_t* buckets);
void b2p_srs_free(b2p_srs* srs);

/* ---- primitives (replace gnark-crypto G1Affine.MultiExp / fft.Domain.FFT) -- */

/* out_affine = sum_i scalars[i] * basis[i], n <= SRS size (Lagrange: n a power of two).
 * Used by the shim for the BSB22 commitment hint (kzg.Commit on pk.KzgLagrange). */
int b2p_msm_g1(b2p_srs* srs, int basis, const void* scalars, uint64_t n, void* out_affine);

/* G2Affine.MultiExp: out = sum_i scalars[i] * g2_points[i].  One shot: n G2Affine (X.A0 X.A1 Y.A0 Y.A1, Montgomery)
 * and n Fr (Montgomery) on the host, the windowed table is built for this call and dropped.  The same Pippenger
 * kernels as G1, instantiated over Fp2 coordinates.  Not on plonk.Prove's path (a PLONK key holds two G2 points,
 * setup/setup.go:124,216-225); it completes "MSM over G1/G2" for callers that build KZG / Groth16-style keys. */
int b2p_msm_g2(int curve, const void* g2_points, const void* scalars, uint64_t n, void* out_g2_affine);

/* Same with the scalars already resident in device memory (device pointer, Montgomery form). */
int b2p_msm_g1_dev(b2p_srs* srs, int basis, const void* d_scalars, uint64_t n, void* out_affine);
/* out_affine = sum of n affine points, computed on the host: the local add that follows the all_gather of
 * the per-GPU partial sums of a point-set-sharded MSM (NCCL has no reduction over group elements). */
int b2p_g1_sum(int curve, const void* points, uint64_t n, void* out_affine);
/* The cudaStream_t every launch of this SRS handle is issued on. */
void* b2p_srs_stream(b2p_srs* srs);
/* Commit hook -- how one proof is spread over several GPUs (point-set-sharded MSMs).  While a hook
 * is set, every kzg.Commit made on this handle (the 9 MSMs of b2p_prove on its circuits, b2p_msm_g1) is
 * delegated: the handle's stream is synchronised, then fn(ctx, d_scalars, n, out) is called on the calling
 * thread with the DEVICE pointer of the n scalars (Fr, Montgomery; valid until fn returns) and must write the
 * commitment sum_j scalars[j] [tau^j]_1 as one G1Affine (Montgomery) to the HOST buffer `out`; non-zero return
 * aborts the call with B2P_ERR_INTERNAL.  fn must not call into THIS handle (its lock is held); it typically
 * broadcasts the scalars to the other ranks, runs b2p_msm_g1_dev on each rank's own SRS shard and adds the
 * partial sums (b2p_g1_sum): algoplonk_b200/sharded_prover.py.  fn = NULL removes the hook. */
typedef int (*b2p_commit_fn)(void* ctx, const void* d_scalars, uint64_t n, void* out_affine);
int b2p_srs_set_commit_hook(b2p_srs* srs, b2p_commit_fn fn, void* ctx);
/* Device-to-device copy on the current device, complete on return: lets a commit hook written in a language without
 * CUDA bindings stage the scalars it was handed into a buffer of its own (e.g. the tensor it broadcasts). */
int b2p_device_copy(void* d_dst, const void* d_src, uint64_t bytes);

/* ---- one proof over the GPUs of one box: commitments and transforms sharded -----------------------------------
 * The native multi-GPU form of plonk.Prove's two hot primitives (BASELINE configs[2]: "2^20 BN254, 1->8 x B200 MSM
 * shard over NVLink", configs[4]: "2^21 BLS12-381, 8 x B200, NTT domain alltoall"; the reference is single-process and
 * has no counterpart).  One process per GPU.  Rank g holds the block [first_g, first_g + count_g) of the SRS (balanced
 * contiguous partition of total_points) as its own b2p_srs `shard` (b2p_srs_generate_unsafe_range, or b2p_srs_load of
 * its slice of pk.Kzg.G1).  Rank 0 runs b2p_prove on an ordinary proving key attached to the group:
 *   commitments (9 kzg.Commit per proof): the scalars are staged in a shared buffer on rank 0, the other ranks'
 *     Pippenger kernels read their slice of it over NVLink (peer loads), store their partial sum into rank 0's
 *     mailbox (peer stores) and raise a flag; rank 0 adds the partial sums on the device;
 *   transforms (ntt_rows = n != 0, world a power of two, circuits without BSB22 commitments and <= 8 public inputs):
 *     the four coset NTTs and the coset iNTT of size 4n run as ntt-shard transforms over all ranks -- local passes,
 *     flags, then a combine / split kernel whose loads / stores over NVLink are the all-to-all; evaluations are written
 *     straight into (read straight from) rank 0's buffers.  Quotient, grand product and openings stay on rank 0.
 * No collective library and no host hop per commitment or transform: the host language exchanges the shared-memory
 * handles once and tells the other ranks "a proof of n rows starts" (they call b2p_shard_group_serve_proof(n), which
 * queues their part of the proof's fixed sequence and returns when it is done).  Proof bytes equal the single-GPU
 * proof's.  Every wait times out (B2P_ERR_INTERNAL) instead of hanging when a rank is missing.
 * Order of calls: create (all ranks) -> attach (rank 0, after b2p_circuit_load: the key's own setup commitments are
 * not sharded; `circuit` is required when ntt_rows != 0) -> export + exchange of the handles by the host language ->
 * connect (all ranks) -> b2p_prove on rank 0 / serve_proof elsewhere.  While attached the SRS handle serves
 * b2p_prove only (b2p_msm_g1 on it fails); attach(NULL, NULL) detaches; free the group before the handles it uses. */
typedef struct b2p_shard_group b2p_shard_group;
#define B2P_SHARD_HANDLES 9       /* shared pieces per rank: mailbox, staging, exchange buffer, el er eo ez h, h staging */
int b2p_shard_group_create(int curve, uint32_t world, uint32_t rank, uint64_t total_points, b2p_srs* shard,
                           uint64_t ntt_rows, b2p_shard_group** out);
int b2p_shard_group_attach(b2p_shard_group* g, b2p_srs* prover_srs, b2p_circuit* circuit);
/* ipc_handles_out: B2P_SHARD_HANDLES x B2P_IPC_HANDLE_BYTES for this rank (zero where it shares nothing) */
int b2p_shard_group_export(b2p_shard_group* g, void* ipc_handles_out);
/* all_handles: world x B2P_SHARD_HANDLES x B2P_IPC_HANDLE_BYTES, rank-major (gathered by the host language) */
int b2p_shard_group_connect(b2p_shard_group* g, const void* all_handles);
/* The same wiring for `world` groups that live in ONE process on one device (plain pointers instead of IPC
 * handles: CUDA cannot map its own allocations); groups[i] must be rank i.  For tests on a single GPU. */
int b2p_shard_group_connect_local(b2p_shard_group* const* groups, uint32_t world);
int b2p_shard_group_serve_proof(b2p_shard_group* g, uint64_t n);
/* One stand-alone commitment over the group -- kzg.Commit / G1Affine.MultiExp in multi-GPU form: rank 0 passes n
 * DEVICE-resident scalars (Fr, Montgomery) and receives the G1Affine sum_j scalars[j] [tau^j]_1; every other rank calls
 * serve_msm(n) for it.  Same exchange as a proof's commitments (peer loads of the scalars, peer store of the partial
 * sum, flags).  No proving key may be attached while the group is used this way. */
int b2p_shard_group_msm(b2p_shard_group* g, const void* d_scalars, uint64_t n, void* out_affine);
int b2p_shard_group_serve_msm(b2p_shard_group* g, uint64_t n);
void b2p_shard_group_free(b2p_shard_group* g);

#define B2P_NTT_INVERSE   1   /* FFTInverse (includes the 1/n scaling) */
#define B2P_NTT_COSET     2   /* on the coset FrMultiplicativeGen * <omega> */
/* In-place, natural order in and out (gnark: FFT(DIF)+BitReverse / BitReverse+FFTInverse(DIT)). */
int b2p_ntt(int curve, void* data, uint64_t n, int flags);

/* ---- NTT with the domain sharded over the GPUs of one box ------------------------------------------
 * The multi-GPU form of fft.Domain.FFT / FFTInverse (BASELINE config 5).  The reference
 * is single-process and has no counterpart; single-GPU semantics are b2p_ntt's.  world = 1, 2, 4 or 8 ranks,
 * one process per GPU, n >= world^2.  Distribution:
 *   coefficients : cyclic      -- rank r holds a[j*world + r] at local index j            (n/world Fr)
 *   evaluations  : bit-reversed order cut in world blocks -- rank r holds A(omega^brev(p)) for
 *                  p in [r*n/world, (r+1)*n/world), i.e. b2p_ntt's DIF output before its bit reversal.
 * A transform is two launches around ONE exchange of world chunks of n/world^2 Fr per rank:
 *   forward :  forward_local (DIF passes on the shard -> d_x)   | exchange |  forward_combine (-> d_out)
 *   inverse :  inverse_split (d_evals -> chunks)                | exchange |  inverse_local (DIT passes in place)
 * d_chunks[r] is where the chunk exchanged with rank r lives: rank r's own buffer + my_rank*chunk when peer
 * memory is mapped (b2p_peer_*: the combine / split kernel then IS the exchange, loads / stores over NVLink,
 * and the caller only orders the ranks with a barrier), or slot r of the receive / send buffer of an
 * all_to_all.  All pointers are device pointers; launches go to `stream` (a cudaStream_t, NULL = default). */
typedef struct b2p_ntt_shard b2p_ntt_shard;
int b2p_ntt_shard_create(int curve, uint64_t n, uint32_t world, uint32_t rank, b2p_ntt_shard** out);
void b2p_ntt_shard_free(b2p_ntt_shard* s);
uint64_t b2p_ntt_shard_local_size(const b2p_ntt_shard* s);   /* n / world   */
uint64_t b2p_ntt_shard_chunk_size(const b2p_ntt_shard* s);   /* n / world^2 */
/* d_x (n/world Fr) <- DIF(local_len <= n/world coefficients of this rank, zero padded); flags: 0 or
 * B2P_NTT_COSET.  Chunk d of d_x (Fr [d*chunk, (d+1)*chunk)) is the part rank d combines. */
int b2p_ntt_shard_forward_local(b2p_ntt_shard* s, const void* d_coeffs, uint64_t local_len, int flags, void* d_x,
                                void* stream);
/* d_out (n/world Fr) <- this rank's block of evaluations from the world chunks d_chunks[r] (chunk my_rank of
 * rank r's d_x). */
int b2p_ntt_shard_forward_combine(b2p_ntt_shard* s, const void* const* d_chunks, void* d_out, void* stream);
/* d_chunks[r] (n/world^2 Fr each) <- the part of rank r's coefficients that this rank's block of
 * evaluations d_evals determines; it belongs at Fr [my_rank*chunk, (my_rank+1)*chunk) of rank r's d_x. */
int b2p_ntt_shard_inverse_split(b2p_ntt_shard* s, const void* d_evals, void* const* d_chunks, void* stream);
/* d_x (n/world Fr, transformed in place): the exchanged chunks -> this rank's coefficients; the 1/n
 * (B2P_NTT_COSET: g^-i / n) of the whole transform is applied here.  flags: B2P_NTT_INVERSE, optionally
 * | B2P_NTT_COSET.  d_out: NULL or d_x leaves the result in d_x; otherwise it is also copied to d_out (d_x is
 * exchange memory the next transform overwrites). */
int b2p_ntt_shard_inverse_local(b2p_ntt_shard* s, void* d_x, int flags, void* d_out, void* stream);

/* Device memory other processes on the box can map (CUDA IPC): the exchange buffers of the sharded NTT.
 * b2p_peer_alloc returns the pointer and, when ipc_handle != NULL, the B2P_IPC_HANDLE_BYTES-byte handle to
 * send to the other ranks; they map it with b2p_peer_open (not valid in the allocating process itself). */
#define B2P_IPC_HANDLE_BYTES 64
int b2p_peer_alloc(uint64_t bytes, void** d_ptr, void* ipc_handle);
int b2p_peer_open(const void* ipc_handle, void** d_ptr);
int b2p_peer_close(void* d_ptr);
int b2p_peer_free(void* d_ptr);

/* ---- circuit (replaces: the trace + domains inside gnark's plonk.ProvingKey) */

/* ql..qk: Lagrange-form selector columns, n Fr each (pk trace; qk WITHOUT public inputs).
 * perm  : gnark trace.S, 3n entries.
 * qcp / commitment_constraint_idx: k BSB22 selector columns and vk.CommitmentConstraintIndexes.
 * vk_transcript: the bytes gnark binds into the gamma challenge before the public inputs
 *   (S1 S2 S3 Ql Qr Qm Qo Qk Qcp*, G1Affine.Marshal() each; reference
 *   verifier/templateLogicSigBN254.go:131-132).  NULL: the library commits to the
 *   columns itself (what plonk.Setup does, setup.go:149) and derives the bytes. */
int b2p_circuit_load(b2p_srs* srs, uint64_t n, uint32_t nb_public,
                     const void* ql, const void* qr, const void* qm, const void* qo, const void* qk,
                     const int64_t* perm, uint32_t k, const void* const* qcp,
                     const uint64_t* commitment_constraint_idx,
                     const void* vk_transcript, uint64_t vk_transcript_len,
                     b2p_circuit** out);
/* The verifying key's commitments: 8+k G1Affine in the order S1 S2 S3 Ql Qr Qm Qo Qk Qcp*. */
int b2p_circuit_vk_commitments(b2p_circuit* c, void* out_points);
void b2p_circuit_free(b2p_circuit* c);

/* ---- prove (replaces: plonk.Prove, algoplonk.go:89) ------------------------ */

/* Raw proof = 9 G1Affine (LRO[0..2], Z, H[0..2], BatchedProof.H, ZShiftedOpening.H)
 * followed by 7+k Fr (BatchedProof.ClaimedValues[0..5+k], ZShiftedOpening.ClaimedValue),
 * all in gnark's in-memory layout: the shim memcpy's them into *plonk_bn254.Proof. */
uint64_t b2p_proof_raw_size(int curve, uint32_t k);

/* L, R, O  : solved wire columns (Lagrange, n Fr each) from gnark's solver.
 * pi2      : k committed-value columns (Lagrange, n Fr each), NULL when k = 0.
 * bsb22    : k G1Affine commitments (proof.Bsb22Commitments), NULL when k = 0.
 * blinding : 9 Fr  bl0 bl1 br0 br1 bo0 bo1 bz0 bz1 bz2  -- the coefficients gnark draws
 *            with fr.SetRandom for the blinding polynomials b_L, b_R, b_O (order 1) and
 *            b_Z (order 2).  They are an INPUT so that proofs are reproducible
 *            bit for bit; the shim draws them from crypto/rand. */
int b2p_prove(b2p_circuit* c, const void* L, const void* R, const void* O,
              const void* const* pi2, const void* bsb22,
              const void* blinding, void* out_proof_raw);

/* Same as b2p_prove with L, R, O and the pi2 columns already resident in device memory
 * (device pointers, Montgomery form); bsb22 / blinding / out_proof_raw stay host pointers.
 * Used by a prover service that keeps witnesses in HBM, and by bench.py's resident-input
 * measurement.  Not reachable from the reference's API (gnark's solver output is host memory). */
int b2p_prove_dev(b2p_circuit* c, const void* dL, const void* dR, const void* dO,
                  const void* const* d_pi2, const void* bsb22,
                  const void* blinding, void* out_proof_raw);
/* The cudaStream_t every launch of this circuit (and its SRS) is issued on, for callers
 * that bracket calls with their own CUDA events. */
void* b2p_circuit_stream(b2p_circuit* c);

/* ---- marshalling (replaces: MarshalProof / MarshalPublicInputs, helper.go:13-110) */

uint64_t b2p_proof_marshal_size(int curve, uint32_t k);      /* (24+3k)*32 / (33+4k)*32 */
int b2p_marshal_proof(int curve, uint32_t k, const void* proof_raw, const void* bsb22, void* out_bytes);
/* public inputs: nb_public Fr (Montgomery) -> nb_public * 32 bytes big-endian */
int b2p_marshal_public_inputs(int curve, const void* values, uint32_t nb_public, void* out_bytes);

/* ---- verification (replaces: plonk.Verify, algoplonk.go:93; testutils/testutils.go:51) ----
 *
 * The self-check (*CompiledCircuit).Verify runs right after plonk.Prove, on the marshalled proof
 * (b2p_marshal_proof) and public inputs (b2p_marshal_public_inputs).  Host arithmetic only -- gnark verifies
 * on the CPU too; these three calls need no GPU and no b2p_init.  The acceptance condition is the one the
 * reference's generated verifiers implement (verifier/templateLogicSigBN254.go:126-397,
 * templateLogicSigBLS12_381.go:144-420) with a real pairing check against the setup's G2 points.
 *
 *   n, nb_public        vk.Size, vk.NbPublicVariables
 *   k, commitment_indexes   len(vk.Qcp), vk.CommitmentConstraintIndexes
 *   vk_points           8+k G1Affine: S[0] S[1] S[2] Ql Qr Qm Qo Qk Qcp[0..k)   (what b2p_circuit_vk_commitments writes)
 *   kzg_g1              vk.Kzg.G1                    (1 G1Affine, = SRS point 0)
 *   kzg_g2              vk.Kzg.G2[0], vk.Kzg.G2[1]   (2 G2Affine: X.A0 X.A1 Y.A0 Y.A1, Montgomery limbs)
 * Returns B2P_OK when the proof is accepted, B2P_ERR_VERIFY when it is rejected (b2p_last_error says at which
 * check: wrong length, value not reduced, point off the curve, pairing), B2P_ERR_ARG for null arguments. */
int b2p_verify(int curve, uint64_t n, uint32_t nb_public, uint32_t k, const uint64_t* commitment_indexes,
               const void* vk_points, const void* kzg_g1, const void* kzg_g2,
               const void* proof_bytes, uint64_t proof_len,
               const void* public_inputs, uint64_t public_len);
/* `count` proofs of the SAME circuit in one call (a prover service checking what it emitted): proof i at
 * proofs + i*proof_len, its public inputs at public_inputs + i*public_len.  Every proof goes through all the
 * checks of b2p_verify up to the pairing; the `count` pairing equations are then folded with 128-bit
 * coefficients hashed from the whole batch into ONE pairing check (soundness error 2^-128; the folding
 * kzg.BatchVerifyMultiPoints applies to the two openings of a single proof).  B2P_OK: all accepted.
 * B2P_ERR_VERIFY: *first_bad = index of the first proof rejected before the pairing, or `count` when only the
 * folded pairing check failed (at least one proof is invalid; b2p_verify one by one finds it).  Batches of 4 or
 * more spread their per-proof work over host threads: env B2P_VERIFY_THREADS, default min(cores, 8). */
int b2p_verify_batch(int curve, uint64_t n, uint32_t nb_public, uint32_t k, const uint64_t* commitment_indexes,
                     const void* vk_points, const void* kzg_g1, const void* kzg_g2,
                     const void* proofs, uint64_t proof_len,
                     const void* public_inputs, uint64_t public_len, uint64_t count, uint64_t* first_bad);
/* The same batch with the group arithmetic on the GPU (needs b2p_init): the three point combinations of every proof
 * -- [Lin], the folded digest, the pairing pair, 9+k / 5+k / 7 points with full-width scalars -- and BLS12-381's
 * r-torsion tests run as one warp per combination, three launches per batch; the SHA-256 transcript and the scalar
 * work stay on host threads; one host pairing check at the end.  Same verdicts and error reporting as
 * b2p_verify_batch (a point outside the r-torsion subgroup is reported as such instead of as a parse failure). */
int b2p_verify_batch_dev(int curve, uint64_t n, uint32_t nb_public, uint32_t k, const uint64_t* commitment_indexes,
                         const void* vk_points, const void* kzg_g1, const void* kzg_g2,
                         const void* proofs, uint64_t proof_len,
                         const void* public_inputs, uint64_t public_len, uint64_t count, uint64_t* first_bad);
/* prod_i e(g1_points[i], g2_points[i]) == 1 ?  (replaces the curve package's PairingCheck, the last line of
 * kzg.BatchVerifyMultiPoints; setup/trusted_setup_test.go checks its setups with the same equation).
 * Points off their curve give B2P_ERR_ARG. */
int b2p_pairing_check(int curve, const void* g1_points, const void* g2_points, uint64_t n, int* is_one);
/* srs.Vk.ReadFrom on an embedded setup/<name>/vk.bin (setup/setup.go:174,190,204-211): 2 compressed G2 + 1
 * compressed G1 (gnark's flag bits: BN254 10/11/01, BLS12-381 100/101/110 on the first byte; G2 as X.A1 || X.A0)
 * -> vk.Kzg.G2 as 2 G2Affine and vk.Kzg.G1 as 1 G1Affine (Montgomery): the kzg_g2 / kzg_g1 arguments of b2p_verify.
 * B2P_ERR_ARG: wrong length, bad flag, coordinate not reduced, x not on the curve / twist. */
int b2p_kzg_vk_load(int curve, const void* vk_bin, uint64_t len, void* out_g2, void* out_g1);
/* [1]_2, [tau]_2 : the G2 half of the TestOnly setups' unsafekzg.NewSRS (setup/setup.go:124);
 * tau as in b2p_srs_generate_unsafe; writes 2 G2Affine. */
int b2p_g2_generate_unsafe(int curve, const void* tau, void* out_g2);

/* ---- persisted keys (replaces: utils.DeserializeCompiledCircuit, utils/utils.go:124-157) ----
 *
 * utils.SerializeCompiledCircuit (utils/utils.go:97-121) writes a gob stream of
 * CompiledCircuitBytes{Ccs, Pk, Vk []byte; Curve ecc.ID}: Pk = plonk.ProvingKey.WriteTo (VerifyingKey, then
 * Kzg and KzgLagrange as uint32 count + compressed G1 -- the format of the embedded setup/<name>/pk.bin),
 * Vk = plonk.VerifyingKey.WriteTo, Ccs = gnark's CBOR constraint system.  A prover service that warm-starts
 * from such a file hands the FILE BYTES to these calls instead of waiting for gnark to decompress 2n+3 points on
 * the CPU:  b2p_gnark_file_parse -> byte ranges;  b2p_gnark_pk_parse -> key fields + where the Kzg points sit;
 * b2p_srs_load_compressed(curve, pk + kzg_off, pk_len - kzg_off, kzg_count, &srs) decompresses them on the GPU.
 * The Ccs range goes back to gnark (the solver needs it on the CPU either way).  Host code: no GPU, no b2p_init.
 * Field order of the VerifyingKey is recalled from gnark v0.15.0, not read from a gnark-written file (there is
 * none here): the parser validates what it decodes and fails with B2P_ERR_ARG rather than return a wrong key. */
#define B2P_MAX_COMMITMENTS 8
typedef struct b2p_gnark_file {
    int32_t  curve;                    /* B2P_BN254 / B2P_BLS12_381 */
    uint32_t ecc_id;                   /* gnark-crypto's ecc.ID as stored (1 = BN254, 3 = BLS12-381) */
    uint64_t ccs_off, ccs_len;         /* byte ranges inside the file */
    uint64_t pk_off, pk_len;
    uint64_t vk_off, vk_len;
} b2p_gnark_file;
typedef struct b2p_gnark_vk {
    uint64_t size, nb_public;          /* vk.Size, vk.NbPublicVariables */
    uint32_t k, has_lines;             /* len(vk.Qcp); whether Kzg.Lines was present (skipped) */
    uint64_t encoded_len;              /* bytes the VerifyingKey occupies */
    uint64_t commitment_indexes[B2P_MAX_COMMITMENTS];
    uint8_t  size_inv[32], generator[32], coset_shift[32];      /* fr.Element memory (Montgomery) */
    uint8_t  points[(8 + B2P_MAX_COMMITMENTS) * 96];           /* S1 S2 S3 Ql Qr Qm Qo Qk Qcp*: G1Affine memory, packed
                                                                  at 64 (BN254) / 96 (BLS12-381) bytes: b2p_verify's vk_points */
    uint8_t  kzg_g1[96];               /* vk.Kzg.G1 */
    uint8_t  kzg_g2[2 * 192];          /* vk.Kzg.G2[0], [1]: G2Affine memory, packed at 128 / 192 bytes */
} b2p_gnark_vk;
typedef struct b2p_gnark_pk {
    b2p_gnark_vk vk;
    uint64_t kzg_off, kzg_count;            /* pk.Kzg: offset of its uint32 header inside the Pk bytes, points */
    uint64_t lagrange_off, lagrange_count;  /* pk.KzgLagrange */
} b2p_gnark_pk;
int b2p_gnark_file_parse(const void* file, uint64_t len, b2p_gnark_file* out);
int b2p_gnark_vk_parse(int curve, const void* vk_bytes, uint64_t len, b2p_gnark_vk* out);
int b2p_gnark_pk_parse(int curve, const void* pk_bytes, uint64_t len, b2p_gnark_pk* out);

/* The library's own key snapshot: everything b2p_circuit_load takes (selector columns, permutation, BSB22 columns,
 * transcript bytes) in one file, so a restarted prover reloads a proving key without gnark rebuilding the trace
 * (NewTrace walks the constraint system: seconds at 2^20).  b2p_circuit_save writes it; b2p_circuit_load_file maps
 * it and uploads straight from the page cache.  Little-endian, Montgomery limbs as in memory; header checked
 * (magic, version, curve, sizes, FNV-1a of the payload). */
int b2p_circuit_save(const char* path, int curve, uint64_t n, uint32_t nb_public,
                     const void* ql, const void* qr, const void* qm, const void* qo, const void* qk,
                     const int64_t* perm, uint32_t k, const void* const* qcp,
                     const uint64_t* commitment_constraint_idx,
                     const void* vk_transcript, uint64_t vk_transcript_len);
int b2p_circuit_load_file(b2p_srs* srs, const char* path, b2p_circuit** out);

/* ---- witness solver (replaces: spr.Solve inside plonk.Prove, algoplonk.go:81-89) ----
 *
 * Every row  ql*a + qr*b + qm*a*b + qo*c + qk = 0  with exactly one unassigned wire determines that wire (gnark's
 * solver rule).  b2p_solver_create analyses the rows once (which wire each row solves, dependency levels -- the
 * levels gnark computes at compile time for its goroutines); b2p_solver_solve runs level after level on the GPU
 * (one thread per row, runs of narrow levels inside one block) or, for deep narrow circuits where a dependency
 * chain is the whole job, on one host thread; B2P_SOLVE_AUTO picks by a cost model of the level structure.
 * Columns are the padded trace of b2p_circuit_load (n rows, public rows first, qk WITHOUT public inputs); xa xb xc
 * give the variable of each row's L R O wire (padding rows and unused wires: variable 0, as gnark pads).
 * input_ids: the variables assigned by the caller (public, then secret), values in the same order in `inputs`.
 * Rows that need a hint make b2p_solver_create fail with B2P_ERR_ARG: use b2p_solver_create_hinted.
 * solve: B2P_ERR_VERIFY "constraint #i is not satisfied" when the assignment breaks a row (every row is checked).
 * L R O: n Fr each (Montgomery), what b2p_prove takes; the _dev form leaves them in HBM (valid until the next
 * solve on this handle) for b2p_prove_dev.  One call at a time per handle. */
#define B2P_SOLVE_AUTO   0
#define B2P_SOLVE_HOST   1
#define B2P_SOLVE_DEVICE 2
typedef struct b2p_solver b2p_solver;
int b2p_solver_create(int curve, uint64_t n, uint32_t nb_public, uint64_t nb_variables,
                      const uint32_t* input_ids, uint32_t nb_inputs,
                      const void* ql, const void* qr, const void* qm, const void* qo, const void* qk,
                      const uint32_t* xa, const uint32_t* xb, const uint32_t* xc, b2p_solver** out);
/* Hints -- gnark's hint functions (bit decompositions, inverse-or-zero, the BSB22 commitment hint, ...): variables a
 * function of the CALLER's computes from other variables.  create_hinted is told which variables each hint reads and
 * produces; the solver places a hint where its inputs are known and calls `fn(ctx, id, inputs, n_in, outputs, n_out)`
 * there during a solve (values as Fr in Montgomery form; return 0, anything else fails the solve with B2P_ERR_INTERNAL).
 * On the device path a hint is a synchronisation point (inputs down, outputs up).  unchecked_rows (n bytes, may be NULL):
 * rows left out of the final gate check -- BSB22's committed rows and commitment row, whose qcp * pi2 term and hash
 * the prover adds (b2p_prove's bsb22 arguments).  The shim passes gnark's own hint functions through a cgo callback. */
typedef struct b2p_hint {
    uint32_t id;                  /* the caller's name for the function (gnark: solver.HintID) */
    uint32_t n_in, n_out;
    const uint32_t* in_vars;      /* variables read */
    const uint32_t* out_vars;     /* variables produced (not inputs, not another hint's outputs) */
} b2p_hint;
typedef int (*b2p_hint_fn)(void* ctx, uint32_t id, const void* inputs, uint32_t n_in, void* outputs, uint32_t n_out);
int b2p_solver_create_hinted(int curve, uint64_t n, uint32_t nb_public, uint64_t nb_variables,
                             const uint32_t* input_ids, uint32_t nb_inputs,
                             const void* ql, const void* qr, const void* qm, const void* qo, const void* qk,
                             const uint32_t* xa, const uint32_t* xb, const uint32_t* xc,
                             const b2p_hint* hints, uint32_t n_hints, const uint8_t* unchecked_rows, b2p_solver** out);
int b2p_solver_set_hint_fn(b2p_solver* s, b2p_hint_fn fn, void* ctx);
int b2p_solver_solve(b2p_solver* s, const void* inputs, int where, void* L, void* R, void* O);
int b2p_solver_solve_dev(b2p_solver* s, const void* inputs, int where, void** dL, void** dR, void** dO);
/* out[8]: levels, widest level, solved rows, launches per solve, estimated host us, estimated device us,
 * last solve in us (wall), where the last solve ran (B2P_SOLVE_HOST / B2P_SOLVE_DEVICE) */
int b2p_solver_info(const b2p_solver* s, uint64_t* out);
void b2p_solver_free(b2p_solver* s);

/* ---- instrumentation -------------------------------------------------------- */

#define B2P_STAT_TOTAL_MS        0   /* b2p_prove wall time incl. H2D/D2H               */
#define B2P_STAT_MSM_MS          1   /* device time in MSM launches (CUDA events)        */
#define B2P_STAT_MSM_ACCUM_MS    2   /* ... of which bucket accumulation                 */
#define B2P_STAT_NTT_MS          3
#define B2P_STAT_QUOTIENT_MS     4
#define B2P_STAT_MSM_CALLS       5
#define B2P_STAT_MSM_ACCUM_ADDS  6   /* mixed additions performed by the accumulation    */
#define B2P_STAT_H2D_BYTES       7
#define B2P_STAT_D2H_BYTES       8
#define B2P_STAT_LAUNCHES        9
#define B2P_STAT_COUNT          16
/* enable!=0: time phases of subsequent b2p_prove calls with CUDA events (adds syncs). */
int b2p_circuit_set_profiling(b2p_circuit* c, int enable);
int b2p_circuit_stats(const b2p_circuit* c, double* out /* B2P_STAT_COUNT doubles */);

#ifdef __cplusplus
}
#endif
#endif /* B200PLONK_H */
