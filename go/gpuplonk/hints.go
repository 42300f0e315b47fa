// BSB22 commitment hint for prove_bn254.go (the BLS12-381 twin lives in prove_bls12381.go).  NOT COMPILED here (no Go).
package gpuplonk

/*
#include "b200plonk.h"
*/
import "C"

import (
	"math/big"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc/bn254"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr/hash_to_field"
	"github.com/consensys/gnark/constraint"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
	"github.com/consensys/gnark/constraint/solver"
)

// bsb22Hints mirrors gnark's bsb22ComputeCommitmentHint (backend/plonk/bn254/prove.go): for commitment
// i the solver hands over the committed wire values; they are written into a Lagrange column that is
// zero elsewhere, two slots (the commitment's own row and the last constraint row) get random values,
// the column is committed on the Lagrange SRS -- here through b2p_msm_g1 -- and the point is hashed to
// the scalar field with DST "BSB22-Plonk" (verifier/templateLogicSigBN254.go:386-397).
func bsb22Hints(spr *cs_bn254.SparseR1CS, key *gpuKey, pi2 [][]fr.Element, coms []bn254.G1Affine, n int) []solver.Option {
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
