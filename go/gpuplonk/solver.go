// solver.go -- the witness solver on the GPU (SURVEY 8f rank 4; spr.Solve inside plonk.Prove,
// /root/reference/algoplonk.go:81-89) for hint-free circuits: rows and wire ids come from the compiled
// SparseR1CS once per key (b2p_solver_create), the assigned inputs per proof (b2p_solver_solve_dev), and L, R, O stay
// in HBM for b2p_prove_dev.  Circuits with hints (BSB22 commitments, std gadgets that call hint functions) keep
// gnark's solver: b2p_solver_create refuses them.  NOT COMPILED in the build container (no Go toolchain).
package gpuplonk

/*
#include "b200plonk.h"
*/
import "C"

import (
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
)

// deviceSolver is created lazily per resident key.
type deviceSolver struct {
	h        *C.b2p_solver
	nbInputs int
}

// newDeviceSolverBN254 flattens the constraint system into the library's row form: the five coefficient columns of
// the padded trace (what NewTrace produces) and the variable id of every row's L, R, O wire.
func newDeviceSolverBN254(spr *cs_bn254.SparseR1CS, ql, qr, qm, qo, qk []fr.Element, n int) (*deviceSolver, error) {
	nbPublic := spr.GetNbPublicVariables()
	nbInputs := nbPublic + spr.GetNbSecretVariables()
	xa, xb, xc := make([]uint32, n), make([]uint32, n), make([]uint32, n)
	for i := 0; i < nbPublic; i++ {
		xa[i] = uint32(i)
	}
	row := nbPublic
	for _, c := range spr.GetSparseR1Cs() { // one row per constraint, in the order NewTrace lays them out
		xa[row], xb[row], xc[row] = c.XA, c.XB, c.XC
		row++
	}
	ids := make([]uint32, nbInputs) // gnark numbers public, then secret, then internal variables
	for i := range ids {
		ids[i] = uint32(i)
	}
	s := &deviceSolver{nbInputs: nbInputs}
	err := call(func() C.int {
		return C.b2p_solver_create(C.B2P_BN254, C.uint64_t(n), C.uint32_t(nbPublic),
			C.uint64_t(spr.GetNbInternalVariables()+nbInputs), (*C.uint32_t)(unsafe.Pointer(&ids[0])), C.uint32_t(nbInputs),
			unsafe.Pointer(&ql[0]), unsafe.Pointer(&qr[0]), unsafe.Pointer(&qm[0]), unsafe.Pointer(&qo[0]), unsafe.Pointer(&qk[0]),
			(*C.uint32_t)(unsafe.Pointer(&xa[0])), (*C.uint32_t)(unsafe.Pointer(&xb[0])), (*C.uint32_t)(unsafe.Pointer(&xc[0])), &s.h)
	})
	if err != nil {
		return nil, err // hints, or a row the rule cannot solve: the caller keeps spr.Solve
	}
	return s, nil
}

// solve assigns every internal variable from the witness vector (public, then secret) and returns device pointers
// to L, R, O -- valid until the next solve on this key (the key's mutex is held by the caller).
func (s *deviceSolver) solve(inputs []fr.Element) (l, r, o unsafe.Pointer, err error) {
	err = call(func() C.int {
		return C.b2p_solver_solve_dev(s.h, unsafe.Pointer(&inputs[0]), C.B2P_SOLVE_AUTO, &l, &r, &o)
	})
	return
}
