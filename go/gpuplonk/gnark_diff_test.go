// gnark_diff_test.go -- closes the "parity unpinned" gap on a box that has Go 1.25 + a module cache
// (SURVEY 8c): proves the same witness with gnark's CPU prover and with the GPU library using the SAME
// blinding scalars, and requires the marshalled proofs to be byte-identical.  gnark draws its blinding
// scalars with fr.SetRandom from crypto/rand.Reader; with a constant-byte reader every draw yields the
// same element regardless of goroutine order, which is what BlindingSource replays.
// NOT RUN in the build container (no Go toolchain).
package gpuplonk

import (
	"bytes"
	"crypto/rand"
	"testing"

	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark/backend/plonk"
	"github.com/consensys/gnark/frontend"
	ap "github.com/giuliop/algoplonk"
	"github.com/giuliop/algoplonk/setup"
)

type constReader struct{ b byte }

func (c constReader) Read(p []byte) (int, error) {
	for i := range p {
		p[i] = c.b
	}
	return len(p), nil
}

type basicCircuit struct { // examples/basic/logicsigVerifier/main.go:30-43
	A, B frontend.Variable `gnark:",public"`
	C    frontend.Variable
}

func (c *basicCircuit) Define(api frontend.API) error {
	aa, bb, cc := api.Mul(c.A, c.A), api.Mul(c.B, c.B), api.Mul(c.C, c.C)
	api.AssertIsEqual(api.Add(aa, bb), cc)
	return nil
}

func TestByteIdenticalToGnark(t *testing.T) {
	cc, err := ap.Compile(&basicCircuit{}, ecc.BN254, setup.TestOnlySetup(ecc.BN254))
	if err != nil {
		t.Fatal(err)
	}
	w, _ := frontend.NewWitness(&basicCircuit{A: 3, B: 4, C: 5}, ecc.BN254.ScalarField())

	old := rand.Reader
	rand.Reader = constReader{0x42}
	var one fr.Element
	one.SetRandom() // what every SetRandom returns under the constant reader
	ref, err := plonk.Prove(cc.Ccs, cc.Pk, w)
	rand.Reader = old
	if err != nil {
		t.Fatal(err)
	}
	BlindingSource = func() (b [9]fr.Element) {
		for i := range b {
			b[i] = one
		}
		return
	}
	got, err := Prove(cc.Ccs, cc.Pk, w)
	if err != nil {
		t.Fatal(err)
	}
	if !bytes.Equal(ap.MarshalProof(ref), ap.MarshalProof(got)) {
		t.Fatal("GPU proof differs from gnark's on the same witness and blinding scalars")
	}
	pub, _ := w.Public()
	if err := plonk.Verify(got, cc.Vk, pub); err != nil {
		t.Fatal(err)
	}
}

// The library's verifier gives gnark's verdicts: gnark's own proof is accepted, the same proof with another public
// witness is rejected by both.
func TestVerifyMatchesGnark(t *testing.T) {
	for _, curve := range []ecc.ID{ecc.BN254, ecc.BLS12_381} {
		cc, err := ap.Compile(&basicCircuit{}, curve, setup.TestOnlySetup(curve))
		if err != nil {
			t.Fatal(err)
		}
		w, _ := frontend.NewWitness(&basicCircuit{A: 3, B: 4, C: 5}, curve.ScalarField())
		proof, err := plonk.Prove(cc.Ccs, cc.Pk, w)
		if err != nil {
			t.Fatal(err)
		}
		pub, _ := w.Public()
		if err := Verify(proof, cc.Vk, pub); err != nil {
			t.Fatalf("%v: gnark's proof rejected: %v", curve, err)
		}
		other, _ := frontend.NewWitness(&basicCircuit{A: 4, B: 3, C: 5}, curve.ScalarField())
		otherPub, _ := other.Public()
		if plonk.Verify(proof, cc.Vk, otherPub) == nil || Verify(proof, cc.Vk, otherPub) == nil {
			t.Fatalf("%v: proof accepted for another public witness", curve)
		}
	}
}
