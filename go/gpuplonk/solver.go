// solver.go -- the witness solver on the GPU (SURVEY 8f rank 4; spr.Solve inside plonk.Prove,
// /root/reference/algoplonk.go:81-89) for hint-free circuits: rows and wire ids come from the compiled
// SparseR1CS once per key (b2p_solver_create), the assigned inputs per proof (b2p_solver_solve_dev), and L, R, O stay
// in HBM for b2p_prove_dev.  Hints stay gnark's: newDeviceSolverBN254 hands b2p_solver_create_hinted the (inputs,
// outputs) of every hint instruction of the constraint system, and goHintTrampoline -- the b2p_hint_fn the library calls
// during a solve -- runs the registered Go hint function (solver.GetRegisteredHint, the BSB22 override of hints.go
// included) on big.Int copies of the values.  NOT COMPILED in the build container (no Go toolchain).
package gpuplonk

/*
#include "b200plonk.h"
extern int goHintTrampoline(void* ctx, uint32_t id, void* inputs, uint32_t n_in, void* outputs, uint32_t n_out);
*/
import "C"

import (
	"math/big"
	"runtime/cgo"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark/constraint"
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

// ---- hints ---------------------------------------------------------------------------------------------------------
//
// gnark's solver executes hint instructions between constraint instructions; the library needs to know only which
// variables each one reads and writes (b2p_hint) and calls back for the values.  The callback below is exported to C
// (//export) and registered with b2p_solver_set_hint_fn; ctx is a cgo.Handle of the *hintTable of the key.

// hintTable: per key, the hint functions by the id given to the library (the index of the hint instruction).
type hintTable struct {
	fns []solverHint
}
type solverHint struct {
	fn      func(mod *big.Int, inputs, outputs []*big.Int) error // solver.Hint
	nIn     int
	nOut    int
}

//export goHintTrampoline
func goHintTrampoline(ctx unsafe.Pointer, id C.uint32_t, inputs unsafe.Pointer, nIn C.uint32_t, outputs unsafe.Pointer, nOut C.uint32_t) C.int {
	tbl := cgo.Handle(uintptr(ctx)).Value().(*hintTable)
	if int(id) >= len(tbl.fns) {
		return 1
	}
	h := tbl.fns[id]
	in := unsafe.Slice((*fr.Element)(inputs), int(nIn))
	out := unsafe.Slice((*fr.Element)(outputs), int(nOut))
	bi := make([]*big.Int, len(in))
	for i := range in {
		bi[i] = in[i].BigInt(new(big.Int)) // Montgomery -> canonical
	}
	bo := make([]*big.Int, len(out))
	for i := range bo {
		bo[i] = new(big.Int)
	}
	if err := h.fn(fr.Modulus(), bi, bo); err != nil {
		return 2
	}
	for i := range out {
		out[i].SetBigInt(bo[i])
	}
	return 0
}

// hintRecords walks the instruction stream of a compiled constraint system and returns, per hint instruction, what the
// library needs (the variables it reads and the range it writes) and what the trampoline needs (the linear
// expressions gnark evaluates before calling the hint function).  gnark v0.15.0: hint instructions carry a blueprint
// implementing constraint.BlueprintHint whose DecompressHint fills a constraint.HintMapping{HintID, Inputs
// []LinearExpression, OutputRange} [UPSTREAM-RECALL: constraint/core.go, constraint/solver.go -- not compiled here].
type hintRecord struct {
	mapping constraint.HintMapping
	inVars  []uint32 // distinct variables of mapping.Inputs, in first-use order: b2p_hint.in_vars
	outVars []uint32 // mapping.OutputRange.Start .. End: b2p_hint.out_vars
}

func hintRecords(sys *constraint.System) []hintRecord {
	var out []hintRecord
	for i := range sys.Instructions {
		inst := sys.Instructions[i]
		bh, ok := sys.Blueprints[inst.BlueprintID].(constraint.BlueprintHint)
		if !ok {
			continue
		}
		var r hintRecord
		bh.DecompressHint(&r.mapping, inst.Unpack(sys))
		seen := map[uint32]bool{}
		for _, le := range r.mapping.Inputs {
			for _, t := range le {
				if v := uint32(t.WireID()); !seen[v] {
					seen[v] = true
					r.inVars = append(r.inVars, v)
				}
			}
		}
		for v := r.mapping.OutputRange.Start; v < r.mapping.OutputRange.End; v++ {
			r.outVars = append(r.outVars, v)
		}
		out = append(out, r)
	}
	return out
}
