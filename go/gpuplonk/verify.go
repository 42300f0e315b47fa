// verify.go -- the reference-side binding of b2p_verify: a drop-in for gnark's plonk.Verify at the call sites
// AlgoPlonk has
//
//	err = plonk.Verify(proof, cc.Vk, publicWitness)        // algoplonk.go:93, testutils/testutils.go:51
//
// becoming
//
//	err = gpuplonk.Verify(proof, cc.Vk, publicWitness)
//
// The library checks the marshalled proof (the bytes ExportProofAndPublicInputs writes, helper.go:13-110) against
// the verifying key with its own host arithmetic and pairing check; nothing is uploaded to the GPU.
// NOT COMPILED in the build container (no Go toolchain), like the rest of this package.
package gpuplonk

/*
#include "b200plonk.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"

	bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381"
	fr_bls12381 "github.com/consensys/gnark-crypto/ecc/bls12-381/fr"
	"github.com/consensys/gnark-crypto/ecc/bn254"
	fr_bn254 "github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark/backend"
	"github.com/consensys/gnark/backend/plonk"
	plonk_bls12381 "github.com/consensys/gnark/backend/plonk/bls12-381"
	plonk_bn254 "github.com/consensys/gnark/backend/plonk/bn254"
	"github.com/consensys/gnark/backend/witness"
)

// ErrInvalidProof is returned when the library rejects the proof; the message names the failed check.
var ErrInvalidProof = errors.New("error verifying proof")

// Verify has plonk.Verify's signature.  Other curves (and verifier options, which change the hash functions) go to gnark.
func Verify(proof plonk.Proof, vk plonk.VerifyingKey, publicWitness witness.Witness, opts ...backend.VerifierOption) error {
	if len(opts) != 0 {
		return plonk.Verify(proof, vk, publicWitness, opts...)
	}
	switch p := proof.(type) {
	case *plonk_bn254.Proof:
		if v, ok := vk.(*plonk_bn254.VerifyingKey); ok {
			return verifyBN254(p, v, publicWitness)
		}
	case *plonk_bls12381.Proof:
		if v, ok := vk.(*plonk_bls12381.VerifyingKey); ok {
			return verifyBLS12381(p, v, publicWitness)
		}
	}
	return plonk.Verify(proof, vk, publicWitness)
}

// verdict runs one verification call and reads its message on the same OS thread (b2p_last_error is thread local).
func verdict(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	switch rc := f(); rc {
	case C.B2P_OK:
		return nil
	case C.B2P_ERR_VERIFY:
		return errors.Join(ErrInvalidProof, errors.New(C.GoString(C.b2p_last_error())))
	default:
		return fmt.Errorf("b200plonk error %d: %s", int(rc), C.GoString(C.b2p_last_error()))
	}
}

func verifyBN254(p *plonk_bn254.Proof, vk *plonk_bn254.VerifyingKey, w witness.Witness) error {
	pub, ok := w.Vector().(fr_bn254.Vector)
	if !ok {
		return errors.New("witness is not over the BN254 scalar field")
	}
	pubBytes := make([]byte, 0, 32*len(pub))
	for i := range pub {
		b := pub[i].Bytes()
		pubBytes = append(pubBytes, b[:]...)
	}
	blob := p.MarshalSolidity() // the layout helper.go:16-17 exports for BN254
	points := append([]bn254.G1Affine{vk.S[0], vk.S[1], vk.S[2], vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk}, vk.Qcp...)
	var cidx *C.uint64_t
	if len(vk.CommitmentConstraintIndexes) > 0 {
		cidx = (*C.uint64_t)(unsafe.Pointer(&vk.CommitmentConstraintIndexes[0]))
	}
	var pubPtr unsafe.Pointer
	if len(pubBytes) > 0 {
		pubPtr = unsafe.Pointer(&pubBytes[0])
	}
	return verdict(func() C.int {
		return C.b2p_verify(C.B2P_BN254, C.uint64_t(vk.Size), C.uint32_t(vk.NbPublicVariables),
		C.uint32_t(len(vk.Qcp)), cidx, unsafe.Pointer(&points[0]), unsafe.Pointer(&vk.Kzg.G1),
		unsafe.Pointer(&vk.Kzg.G2[0]), unsafe.Pointer(&blob[0]), C.uint64_t(len(blob)), pubPtr, C.uint64_t(len(pubBytes)))
	})
}

func verifyBLS12381(p *plonk_bls12381.Proof, vk *plonk_bls12381.VerifyingKey, w witness.Witness) error {
	pub, ok := w.Vector().(fr_bls12381.Vector)
	if !ok {
		return errors.New("witness is not over the BLS12-381 scalar field")
	}
	pubBytes := make([]byte, 0, 32*len(pub))
	for i := range pub {
		b := pub[i].Bytes()
		pubBytes = append(pubBytes, b[:]...)
	}
	// helper.go:27-88: L R O | H0 H1 H2 | l r o s1 s2 | Z | z(omega zeta) | W_zeta | W_{omega zeta} | qcp_i | Bsb22_i
	var blob []byte
	pt := func(a *bls12381.G1Affine) { b := a.RawBytes(); blob = append(blob, b[:]...) }
	sc := func(e *fr_bls12381.Element) { b := e.Bytes(); blob = append(blob, b[:]...) }
	for i := 0; i < 3; i++ {
		pt(&p.LRO[i])
	}
	for i := 0; i < 3; i++ {
		pt(&p.H[i])
	}
	cv := p.BatchedProof.ClaimedValues // [lin, l, r, o, s1, s2, qcp...]: the first is not serialised
	for i := 1; i < 6; i++ {
		sc(&cv[i])
	}
	pt(&p.Z)
	sc(&p.ZShiftedOpening.ClaimedValue)
	pt(&p.BatchedProof.H)
	pt(&p.ZShiftedOpening.H)
	for i := range vk.Qcp {
		sc(&cv[6+i])
	}
	for i := range p.Bsb22Commitments {
		pt(&p.Bsb22Commitments[i])
	}
	points := append([]bls12381.G1Affine{vk.S[0], vk.S[1], vk.S[2], vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk}, vk.Qcp...)
	var cidx *C.uint64_t
	if len(vk.CommitmentConstraintIndexes) > 0 {
		cidx = (*C.uint64_t)(unsafe.Pointer(&vk.CommitmentConstraintIndexes[0]))
	}
	var pubPtr unsafe.Pointer
	if len(pubBytes) > 0 {
		pubPtr = unsafe.Pointer(&pubBytes[0])
	}
	return verdict(func() C.int {
		return C.b2p_verify(C.B2P_BLS12_381, C.uint64_t(vk.Size), C.uint32_t(vk.NbPublicVariables),
		C.uint32_t(len(vk.Qcp)), cidx, unsafe.Pointer(&points[0]), unsafe.Pointer(&vk.Kzg.G1),
		unsafe.Pointer(&vk.Kzg.G2[0]), unsafe.Pointer(&blob[0]), C.uint64_t(len(blob)), pubPtr, C.uint64_t(len(pubBytes)))
	})
}
