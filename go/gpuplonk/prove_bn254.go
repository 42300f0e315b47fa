// Package gpuplonk is the reference-side binding of libb200plonk.so: a drop-in for
// gnark's plonk.Prove at the one call site AlgoPlonk has
//
//	proof, err := plonk.Prove(cc.Ccs, cc.Pk, witness)        // algoplonk.go:89, testutils/testutils.go:47
//
// becoming
//
//	proof, err := gpuplonk.Prove(cc.Ccs, cc.Pk, witness)
//
// The witness solver stays gnark's (CPU); the five proving rounds run on the GPU.
// NOT COMPILED in the build container (no Go toolchain, no gnark module cache): this file is
// the binding a maintainer adds, kept mechanical on purpose.  It needs gnark v0.15.0 /
// gnark-crypto v0.20.1 (go.mod:8-9) and cgo with -lb200plonk.
//
// BN254 only in this file; prove_bls12381.go is the same text with the bls12-381 packages.
package gpuplonk

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../algoplonk_b200 -lb200plonk
#include <stdlib.h>
#include "b200plonk.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	fft_bn254 "github.com/consensys/gnark-crypto/ecc/bn254/fr/fft"
	"github.com/consensys/gnark/backend"
	"github.com/consensys/gnark/backend/plonk"
	plonk_bn254 "github.com/consensys/gnark/backend/plonk/bn254"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
	"github.com/consensys/gnark/constraint/solver"
)

// BlindingSource lets tests fix the 9 blinding scalars (bl0 bl1 br0 br1 bo0 bo1 bz0 bz1 bz2);
// nil = crypto/rand, like gnark's fr.SetRandom.
var BlindingSource func() [9]fr.Element

// upload builds the device-resident key once per proving key: SRS table + selector / permutation columns.
func upload(spr *cs_bn254.SparseR1CS, pk *plonk_bn254.ProvingKey) (*gpuKey, error) {
	if k := lookup(pk); k != nil {
		return k, nil
	}
	if err := call(func() C.int { return C.b2p_init(-1) }); err != nil {
		return nil, err
	}
	k := &gpuKey{}
	var err error
	g1 := pk.Kzg.G1 // canonical SRS, n+3 points, gnark in-memory layout == library layout
	if err := call(func() C.int {
		return C.b2p_srs_load(C.B2P_BN254, unsafe.Pointer(&g1[0]), C.uint64_t(len(g1)), nil, 0, &k.srs)
	}); err != nil {
		return nil, err
	}
	if k.circuit, err = loadCircuitBN254(k.srs, spr, pk.Vk, ""); err != nil {
		C.b2p_srs_free(k.srs)
		return nil, err
	}
	k.allocColumns(int(pk.Vk.Size))
	remember(pk, k) // may evict the least recently used key (MaxResidentKeys)
	return k, nil
}

// loadCircuitBN254 builds the trace of a constraint system (gnark's NewTrace) and makes it resident on the SRS handle's
// GPU; with a non-empty snapshotPath it also writes the library's own snapshot of it (b2p_circuit_save), which
// persist.go's warm start reads back with b2p_circuit_load_file instead of rebuilding the trace.
func loadCircuitBN254(srs *C.b2p_srs, spr *cs_bn254.SparseR1CS, vk *plonk_bn254.VerifyingKey, snapshotPath string) (*C.b2p_circuit, error) {
	// gnark v0.15: NewTrace(spr *cs.SparseR1CS, domain *fft.Domain) -- Lagrange-form ql qr qm qo qk, S, qcp
	trace := plonk_bn254.NewTrace(spr, fft_bn254.NewDomain(vk.Size))
	n := C.uint64_t(vk.Size)
	col := func(p interface{ Coefficients() []fr.Element }) unsafe.Pointer {
		return unsafe.Pointer(&p.Coefficients()[0])
	}
	nq := len(trace.Qcp)
	qcpPtrs := make([]unsafe.Pointer, nq)
	for i := range trace.Qcp {
		qcpPtrs[i] = col(trace.Qcp[i])
	}
	qcp, unpin := pointerArray(qcpPtrs) // Go pointers inside an array handed to C: pinned for the call
	defer unpin()
	var cidx *C.uint64_t
	if nq > 0 {
		cidx = (*C.uint64_t)(unsafe.Pointer(&vk.CommitmentConstraintIndexes[0]))
	}
	// the VK digests gnark binds into gamma: S1 S2 S3 Ql Qr Qm Qo Qk Qcp*, Marshal() each
	var vkb []byte
	for _, p := range append(append([]bn254.G1Affine{}, vk.S[:]...), vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk) {
		vkb = append(vkb, p.Marshal()...)
	}
	for _, p := range vk.Qcp {
		vkb = append(vkb, p.Marshal()...)
	}
	if snapshotPath != "" {
		cs := C.CString(snapshotPath)
		defer C.free(unsafe.Pointer(cs))
		if err := call(func() C.int {
			return C.b2p_circuit_save(cs, C.B2P_BN254, n, C.uint32_t(vk.NbPublicVariables),
				col(trace.Ql), col(trace.Qr), col(trace.Qm), col(trace.Qo), col(trace.Qk),
				(*C.int64_t)(unsafe.Pointer(&trace.S[0])), C.uint32_t(nq), (*unsafe.Pointer)(unsafe.Pointer(qcp)), cidx,
				unsafe.Pointer(&vkb[0]), C.uint64_t(len(vkb)))
		}); err != nil {
			return nil, err
		}
	}
	var c *C.b2p_circuit
	if err := call(func() C.int {
		return C.b2p_circuit_load(srs, n, C.uint32_t(vk.NbPublicVariables),
			col(trace.Ql), col(trace.Qr), col(trace.Qm), col(trace.Qo), col(trace.Qk),
			(*C.int64_t)(unsafe.Pointer(&trace.S[0])), C.uint32_t(nq), (*unsafe.Pointer)(unsafe.Pointer(qcp)), cidx,
			unsafe.Pointer(&vkb[0]), C.uint64_t(len(vkb)), &c)
	}); err != nil {
		return nil, err
	}
	return c, nil
}

// Prove has plonk.Prove's signature.  Any failure of the GPU path falls back to gnark's CPU prover
// (the fallback lives HERE, in the caller's language; the library itself has none).
func Prove(ccs constraint.ConstraintSystem, pk plonk.ProvingKey, fullWitness witness.Witness,
	opts ...backend.ProverOption) (plonk.Proof, error) {
	spr, ok1 := ccs.(*cs_bn254.SparseR1CS)
	bpk, ok2 := pk.(*plonk_bn254.ProvingKey)
	if !ok1 || !ok2 {
		return proveOtherCurves(ccs, pk, fullWitness, opts...) // prove_bls12381.go, else plonk.Prove
	}
	proof, err := proveBN254(spr, bpk, fullWitness, opts...)
	if err != nil {
		var unsat *solver.UnsatisfiedConstraintError
		if errors.As(err, &unsat) {
			return nil, err // same error gnark would return
		}
		return plonk.Prove(ccs, pk, fullWitness, opts...)
	}
	return proof, nil
}

func proveBN254(spr *cs_bn254.SparseR1CS, pk *plonk_bn254.ProvingKey, fullWitness witness.Witness,
	opts ...backend.ProverOption) (*plonk_bn254.Proof, error) {
	key, err := upload(spr, pk)
	if err != nil {
		return nil, err
	}
	// One proof at a time per device-resident key.  (A service that wants several proofs in flight per GPU
	// uploads the key more than once -- bench.py does exactly that -- the library is re-entrant across handles.)
	key.mu.Lock()
	defer key.mu.Unlock()
	// Not required for correctness (every entry point switches to the handle's CUDA device itself), but it
	// keeps the blocking cgo call from being counted against GOMAXPROCS scheduling.
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()

	popt, err := backend.NewProverConfig(opts...)
	if err != nil {
		return nil, err
	}
	proof := &plonk_bn254.Proof{}
	n := int(pk.Vk.Size)
	k := len(pk.Vk.CommitmentConstraintIndexes)
	pi2 := make([][]fr.Element, k)
	proof.Bsb22Commitments = make([]bn254.G1Affine, k)

	// BSB22 hint override (gnark prove.go bsb22ComputeCommitmentHint): commit the committed wires on
	// the Lagrange basis with b2p_msm_g1, hash the point to the field.
	if k > 0 {
		popt.SolverOpts = append(popt.SolverOpts, bsb22Hints(spr, key, pi2, proof.Bsb22Commitments, n)...)
	}
	w, okw := fullWitness.Vector().(fr.Vector)
	if !okw {
		return nil, witness.ErrInvalidWitness
	}
	sol, err := spr.Solve(w, popt.SolverOpts...)
	if err != nil {
		return nil, err
	}
	s := sol.(*cs_bn254.SparseR1CSSolution)
	// the key's own page-locked columns (allocated once at upload; key.mu is held): no per-proof allocation
	L, R, O := pad(key, 0, s.L, n), pad(key, 1, s.R, n), pad(key, 2, s.O, n)

	var blinding [9]fr.Element
	if BlindingSource != nil {
		blinding = BlindingSource()
	} else {
		for i := range blinding {
			if _, err := blinding[i].SetRandom(); err != nil {
				return nil, err
			}
		}
	}

	raw := make([]byte, int(C.b2p_proof_raw_size(C.B2P_BN254, C.uint32_t(k))))
	pi2Ptrs := make([]unsafe.Pointer, k)
	for i := range pi2 {
		pi2Ptrs[i] = unsafe.Pointer(&pi2[i][0])
	}
	pi2p, unpin := pointerArray(pi2Ptrs)
	defer unpin()
	var bsbp unsafe.Pointer
	if k > 0 {
		bsbp = unsafe.Pointer(&proof.Bsb22Commitments[0])
	}
	if err := call(func() C.int {
		return C.b2p_prove(key.circuit, unsafe.Pointer(&L[0]), unsafe.Pointer(&R[0]), unsafe.Pointer(&O[0]),
			(*unsafe.Pointer)(unsafe.Pointer(pi2p)), bsbp, unsafe.Pointer(&blinding[0]), unsafe.Pointer(&raw[0]))
	}); err != nil {
		return nil, err
	}
	// raw = 9 G1Affine then 7+k fr.Element, gnark memory layout: copy into the gnark struct
	pts := unsafe.Slice((*bn254.G1Affine)(unsafe.Pointer(&raw[0])), 9)
	frs := unsafe.Slice((*fr.Element)(unsafe.Pointer(&raw[9*64])), 7+k)
	copy(proof.LRO[:], pts[0:3])
	proof.Z = pts[3]
	copy(proof.H[:], pts[4:7])
	proof.BatchedProof.H = pts[7]
	proof.ZShiftedOpening.H = pts[8]
	proof.BatchedProof.ClaimedValues = append([]fr.Element{}, frs[:6+k]...)
	proof.ZShiftedOpening.ClaimedValue = frs[6+k]
	return proof, nil
}

// pad copies a solver column into column `which` of the key's page-locked set (zero padded to n elements);
// without pinned memory (allocation failed at upload) it falls back to a pageable slice: correct, slower upload.
func pad(key *gpuKey, which int, v []fr.Element, n int) []fr.Element {
	var out []fr.Element
	if p := key.cols[which]; p != nil && key.n == n {
		out = unsafe.Slice((*fr.Element)(p), n)
	} else {
		out = make([]fr.Element, n)
	}
	k := copy(out, v)
	for i := k; i < n; i++ {
		out[i] = fr.Element{}
	}
	return out
}

