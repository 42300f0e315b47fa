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

var keysBls = map[*plonk_bls12381.ProvingKey]*gpuKey{}

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
	mu.Lock()
	defer mu.Unlock()
	if k, ok := keysBls[pk]; ok {
		return k, nil
	}
	if rc := C.b2p_init(-1); rc != 0 {
		return nil, lastErr(rc)
	}
	k := &gpuKey{}
	g1 := pk.Kzg.G1 // canonical SRS, n+3 points, gnark in-memory layout == library layout
	if rc := C.b2p_srs_load(C.B2P_BLS12_381, unsafe.Pointer(&g1[0]), C.uint64_t(len(g1)), nil, 0, &k.srs); rc != 0 {
		return nil, lastErr(rc)
	}
	trace := plonk_bls12381.NewTrace(spr, pk.Vk.Size) // Lagrange-form ql qr qm qo qk, S, qcp
	n := C.uint64_t(pk.Vk.Size)
	col := func(p interface{ Coefficients() []fr.Element }) unsafe.Pointer {
		return unsafe.Pointer(&p.Coefficients()[0])
	}
	nq := len(trace.Qcp)
	var qcp *unsafe.Pointer
	var cidx *C.uint64_t
	if nq > 0 {
		ptrs := (*[1 << 10]unsafe.Pointer)(C.malloc(C.size_t(nq) * C.size_t(unsafe.Sizeof(uintptr(0)))))
		defer C.free(unsafe.Pointer(ptrs))
		for i := range trace.Qcp {
			ptrs[i] = col(trace.Qcp[i])
		}
		qcp = &ptrs[0]
		cidx = (*C.uint64_t)(unsafe.Pointer(&pk.Vk.CommitmentConstraintIndexes[0]))
	}
	// the VK digests gnark binds into gamma: S1 S2 S3 Ql Qr Qm Qo Qk Qcp*, Marshal() each
	var vkb []byte
	for _, p := range append(append([]bls12381.G1Affine{}, pk.Vk.S[:]...), pk.Vk.Ql, pk.Vk.Qr, pk.Vk.Qm, pk.Vk.Qo, pk.Vk.Qk) {
		vkb = append(vkb, p.Marshal()...)
	}
	for _, p := range pk.Vk.Qcp {
		vkb = append(vkb, p.Marshal()...)
	}
	rc := C.b2p_circuit_load(k.srs, n, C.uint32_t(pk.Vk.NbPublicVariables),
		col(trace.Ql), col(trace.Qr), col(trace.Qm), col(trace.Qo), col(trace.Qk),
		(*C.int64_t)(unsafe.Pointer(&trace.S[0])), C.uint32_t(nq), qcp, cidx,
		unsafe.Pointer(&vkb[0]), C.uint64_t(len(vkb)), &k.circuit)
	if rc != 0 {
		C.b2p_srs_free(k.srs)
		return nil, lastErr(rc)
	}
	keysBls[pk] = k
	return k, nil
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
	L, R, O := padBls(s.L, n), padBls(s.R, n), padBls(s.O, n)
	defer releaseBls(L, R, O)

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
	var pi2p *unsafe.Pointer
	var bsbp unsafe.Pointer
	if k > 0 {
		ptrs := (*[1 << 10]unsafe.Pointer)(C.malloc(C.size_t(k) * C.size_t(unsafe.Sizeof(uintptr(0)))))
		defer C.free(unsafe.Pointer(ptrs))
		for i := range pi2 {
			ptrs[i] = unsafe.Pointer(&pi2[i][0])
		}
		pi2p = &ptrs[0]
		bsbp = unsafe.Pointer(&proof.Bsb22Commitments[0])
	}
	rc := C.b2p_prove(key.circuit, unsafe.Pointer(&L[0]), unsafe.Pointer(&R[0]), unsafe.Pointer(&O[0]),
		pi2p, bsbp, unsafe.Pointer(&blinding[0]), unsafe.Pointer(&raw[0]))
	if rc != 0 {
		return nil, lastErr(rc)
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
func padBls(v []fr.Element, n int) []fr.Element {
	var p unsafe.Pointer
	if rc := C.b2p_host_alloc(C.uint64_t(n)*C.uint64_t(unsafe.Sizeof(fr.Element{})), &p); rc != 0 {
		out := make([]fr.Element, n) // pageable fallback: correct, slower upload
		copy(out, v)
		return out
	}
	out := unsafe.Slice((*fr.Element)(p), n)
	k := copy(out, v)
	for i := k; i < n; i++ {
		out[i] = fr.Element{}
	}
	pinned.Store(p, struct{}{})
	return out
}

func releaseBls(cols ...[]fr.Element) {
	for _, c := range cols {
		if len(c) == 0 {
			continue
		}
		p := unsafe.Pointer(&c[0])
		if _, ok := pinned.LoadAndDelete(p); ok {
			C.b2p_host_free(p)
		}
	}
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
			if rc := C.b2p_msm_g1(key.srs, C.B2P_BASIS_LAGRANGE, unsafe.Pointer(&col[0]), C.uint64_t(n),
				unsafe.Pointer(&coms[i])); rc != 0 {
				return lastErr(rc)
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
