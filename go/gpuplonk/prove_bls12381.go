// prove_bls12381.go -- the BLS12-381 twin of prove_bn254.go (same text, bls12-381 packages, 96-byte points):
// AlgoPlonk supports exactly these two curves (algoplonk.go:39-41).  Generated mechanically from
// prove_bn254.go + hints.go; NOT COMPILED in the build container (no Go toolchain).
package gpuplonk

/*
#include <stdlib.h>
#include "b200plonk.h"
*/
import "C"

import (
	"errors"
	"math/big"
	"runtime"
	"unsafe"

	bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381"
	"github.com/consensys/gnark-crypto/ecc/bls12-381/fr"
	fft_bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381/fr/fft"
	"github.com/consensys/gnark-crypto/ecc/bls12-381/fr/hash_to_field"
	"github.com/consensys/gnark/backend"
	"github.com/consensys/gnark/backend/plonk"
	plonk_bls12381 "github.com/consensys/gnark/backend/plonk/bls12-381"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
	cs_bls12381 "github.com/consensys/gnark/constraint/bls12-381"
	"github.com/consensys/gnark/constraint/solver"
)

// BlindingSourceBls: see BlindingSource (prove_bn254.go).
var BlindingSourceBls func() [9]fr.Element

// proveOtherCurves: BLS12-381 runs on the GPU like BN254; anything else is not an AlgoPlonk curve and stays on
// gnark, as does any failure of the GPU path other than an unsatisfied constraint system.
func proveOtherCurves(ccs constraint.ConstraintSystem, pk plonk.ProvingKey, w witness.Witness,
	opts ...backend.ProverOption) (plonk.Proof, error) {
	spr, ok1 := ccs.(*cs_bls12381.SparseR1CS)
	bpk, ok2 := pk.(*plonk_bls12381.ProvingKey)
	if ok1 && ok2 {
		proof, err := proveBLS12381(spr, bpk, w, opts...)
		if err == nil {
			return proof, nil
		}
		var unsat *solver.UnsatisfiedConstraintError
		if errors.As(err, &unsat) {
			return nil, err
		}
	}
	return plonk.Prove(ccs, pk, w, opts...)
}

// upload builds the device-resident key once per proving key: SRS table + selector / permutation columns.
func uploadBls(spr *cs_bls12381.SparseR1CS, pk *plonk_bls12381.ProvingKey) (*gpuKey, error) {
	if k := lookup(pk); k != nil {
		return k, nil
	}
	if err := call(func() C.int { return C.b2p_init(-1) }); err != nil {
		return nil, err
	}
	k := &gpuKey{}
	g1 := pk.Kzg.G1 // canonical SRS, n+3 points, gnark in-memory layout == library layout
	if err := call(func() C.int {
		return C.b2p_srs_load(C.B2P_BLS12_381, unsafe.Pointer(&g1[0]), C.uint64_t(len(g1)), nil, 0, &k.srs)
	}); err != nil {
		return nil, err
	}
	var err error
	if k.circuit, err = loadCircuitBLS12381(k.srs, spr, pk.Vk, ""); err != nil {
		C.b2p_srs_free(k.srs)
		return nil, err
	}
	k.allocColumns(int(pk.Vk.Size))
	remember(pk, k) // may evict the least recently used key (MaxResidentKeys)
	return k, nil
}

// loadCircuitBLS12381: prove_bn254.go's loadCircuitBN254 with the bls12-381 packages.
func loadCircuitBLS12381(srs *C.b2p_srs, spr *cs_bls12381.SparseR1CS, vk *plonk_bls12381.VerifyingKey, snapshotPath string) (*C.b2p_circuit, error) {
	trace := plonk_bls12381.NewTrace(spr, fft_bls12381.NewDomain(vk.Size))
	n := C.uint64_t(vk.Size)
	col := func(p interface{ Coefficients() []fr.Element }) unsafe.Pointer {
		return unsafe.Pointer(&p.Coefficients()[0])
	}
	nq := len(trace.Qcp)
	qcpPtrs := make([]unsafe.Pointer, nq)
	for i := range trace.Qcp {
		qcpPtrs[i] = col(trace.Qcp[i])
	}
	qcp, unpin := pointerArray(qcpPtrs)
	defer unpin()
	var cidx *C.uint64_t
	if nq > 0 {
		cidx = (*C.uint64_t)(unsafe.Pointer(&vk.CommitmentConstraintIndexes[0]))
	}
	var vkb []byte
	for _, p := range append(append([]bls12381.G1Affine{}, vk.S[:]...), vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk) {
		vkb = append(vkb, p.Marshal()...)
	}
	for _, p := range vk.Qcp {
		vkb = append(vkb, p.Marshal()...)
	}
	if snapshotPath != "" {
		cs := C.CString(snapshotPath)
		defer C.free(unsafe.Pointer(cs))
		if err := call(func() C.int {
			return C.b2p_circuit_save(cs, C.B2P_BLS12_381, n, C.uint32_t(vk.NbPublicVariables),
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

// loadCircuitOtherCurves is persist.go's dispatch for keys that are not BN254.
func loadCircuitOtherCurves(w *WarmKey, snapshotPath string) (*C.b2p_circuit, error) {
	spr, ok1 := w.Ccs.(*cs_bls12381.SparseR1CS)
	vk, ok2 := w.Vk.(*plonk_bls12381.VerifyingKey)
	if !ok1 || !ok2 {
		return nil, errors.New("gpuplonk: compiled circuit is on a curve AlgoPlonk does not support")
	}
	return loadCircuitBLS12381(w.srs, spr, vk, snapshotPath)
}

func proveBLS12381(spr *cs_bls12381.SparseR1CS, pk *plonk_bls12381.ProvingKey, fullWitness witness.Witness,
	opts ...backend.ProverOption) (*plonk_bls12381.Proof, error) {
	key, err := uploadBls(spr, pk)
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
	proof := &plonk_bls12381.Proof{}
	n := int(pk.Vk.Size)
	k := len(pk.Vk.CommitmentConstraintIndexes)
	pi2 := make([][]fr.Element, k)
	proof.Bsb22Commitments = make([]bls12381.G1Affine, k)

	// BSB22 hint override (gnark prove.go bsb22ComputeCommitmentHint): commit the committed wires on
	// the Lagrange basis with b2p_msm_g1, hash the point to the field.
	if k > 0 {
		popt.SolverOpts = append(popt.SolverOpts, bsb22HintsBls(spr, key, pi2, proof.Bsb22Commitments, n)...)
	}
	w, okw := fullWitness.Vector().(fr.Vector)
	if !okw {
		return nil, witness.ErrInvalidWitness
	}
	sol, err := spr.Solve(w, popt.SolverOpts...)
	if err != nil {
		return nil, err
	}
	s := sol.(*cs_bls12381.SparseR1CSSolution)
	// the key's own page-locked columns (allocated once at upload; key.mu is held): no per-proof allocation
	L, R, O := padBls(key, 0, s.L, n), padBls(key, 1, s.R, n), padBls(key, 2, s.O, n)

	var blinding [9]fr.Element
	if BlindingSourceBls != nil {
		blinding = BlindingSourceBls()
	} else {
		for i := range blinding {
			if _, err := blinding[i].SetRandom(); err != nil {
				return nil, err
			}
		}
	}

	raw := make([]byte, int(C.b2p_proof_raw_size(C.B2P_BLS12_381, C.uint32_t(k))))
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
	pts := unsafe.Slice((*bls12381.G1Affine)(unsafe.Pointer(&raw[0])), 9)
	frs := unsafe.Slice((*fr.Element)(unsafe.Pointer(&raw[9*96])), 7+k)
	copy(proof.LRO[:], pts[0:3])
	proof.Z = pts[3]
	copy(proof.H[:], pts[4:7])
	proof.BatchedProof.H = pts[7]
	proof.ZShiftedOpening.H = pts[8]
	proof.BatchedProof.ClaimedValues = append([]fr.Element{}, frs[:6+k]...)
	proof.ZShiftedOpening.ClaimedValue = frs[6+k]
	return proof, nil
}

// pad copies a solver column into a page-locked buffer of n elements (zero padded): pinned memory uploads at
// PCIe speed and overlaps with the first transforms; the buffer is returned to the pool after the proof.
// padBls copies a solver column into column `which` of the key's page-locked set (zero padded to n elements);
// without pinned memory (allocation failed at upload) it falls back to a pageable slice: correct, slower upload.
func padBls(key *gpuKey, which int, v []fr.Element, n int) []fr.Element {
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

// bsb22Hints mirrors gnark's bsb22ComputeCommitmentHint (backend/plonk/bls12-381/prove.go): for commitment
// i the solver hands over the committed wire values; they are written into a Lagrange column that is
// zero elsewhere, two slots (the commitment's own row and the last constraint row) get random values,
// the column is committed on the Lagrange SRS -- here through b2p_msm_g1 -- and the point is hashed to
// the scalar field with DST "BSB22-Plonk" (verifier/templateLogicSigBLS12_381.go:404-420).
func bsb22HintsBls(spr *cs_bls12381.SparseR1CS, key *gpuKey, pi2 [][]fr.Element, coms []bls12381.G1Affine, n int) []solver.Option {
	infos := spr.CommitmentInfo.(constraint.PlonkCommitments)
	out := make([]solver.Option, 0, len(infos))
	for i := range infos {
		i := i
		out = append(out, solver.OverrideHint(infos[i].HintID, func(_ *big.Int, ins, outs []*big.Int) error {
			col := make([]fr.Element, n)
			offset := spr.GetNbPublicVariables()
			for j, row := range infos[i].Committed {
				col[offset+row].SetBigInt(ins[j])
			}
			if _, err := col[offset+infos[i].CommitmentIndex].SetRandom(); err != nil {
				return err
			}
			if _, err := col[offset+spr.GetNbConstraints()-1].SetRandom(); err != nil {
				return err
			}
			pi2[i] = col
			if err := call(func() C.int {
				return C.b2p_msm_g1(key.srs, C.B2P_BASIS_LAGRANGE, unsafe.Pointer(&col[0]), C.uint64_t(n),
					unsafe.Pointer(&coms[i]))
			}); err != nil {
				return err
			}
			h := hash_to_field.New([]byte("BSB22-Plonk"))
			h.Write(coms[i].Marshal())
			var res fr.Element
			res.SetBytes(h.Sum(nil))
			res.BigInt(outs[0])
			return nil
		}))
	}
	return out
}
