// persist.go -- the persisted-key fast path (SURVEY 8f rank 2): utils.DeserializeCompiledCircuit
// (/root/reference/utils/utils.go:124-157) for a prover that keeps its keys on the GPU.  gnark's ProvingKey.ReadFrom
// decompresses 2n+3 G1 points on the CPU (n+3 + n square roots in Fp: seconds at 2^20); here the file's bytes go to
// the library, which finds the Kzg section (b2p_gnark_file_parse / b2p_gnark_pk_parse) and decompresses it on the GPU
// (b2p_srs_load_compressed) while gnark decodes the constraint system on a goroutine.  NOT COMPILED in the build
// container (no Go toolchain); the C entry points it binds are exercised by tests/test_keyfile.py and
// tests/test_gpu_keyfile.py through ctypes.
package gpuplonk

/*
#include <stdlib.h>
#include "b200plonk.h"
*/
import "C"

import (
	"bytes"
	"fmt"
	"os"
	"unsafe"

	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark/backend/plonk"
	plonk_bn254 "github.com/consensys/gnark/backend/plonk/bn254"
	"github.com/consensys/gnark/constraint"
	cs_bn254 "github.com/consensys/gnark/constraint/bn254"
)

// WarmKey is what LoadCompiledCircuit returns: the constraint system (for gnark's solver), the verifying key, and
// a proving key whose point data already sits in HBM.
type WarmKey struct {
	Ccs   constraint.ConstraintSystem
	Vk    plonk.VerifyingKey
	Curve ecc.ID
	srs   *C.b2p_srs
	info  C.b2p_gnark_pk
}

// LoadCompiledCircuit reads a file written by utils.SerializeCompiledCircuit.
func LoadCompiledCircuit(path string) (*WarmKey, error) {
	data, err := os.ReadFile(path)
	if err != nil {
		return nil, fmt.Errorf("error reading compiled circuit file: %v", err)
	}
	var f C.b2p_gnark_file
	if err := call(func() C.int {
		return C.b2p_gnark_file_parse(unsafe.Pointer(&data[0]), C.uint64_t(len(data)), &f)
	}); err != nil {
		return nil, fmt.Errorf("error decoding compiled circuit: %v", err)
	}
	w := &WarmKey{Curve: ecc.ID(f.ecc_id)}
	pk := data[f.pk_off : f.pk_off+f.pk_len]
	if err := call(func() C.int {
		return C.b2p_gnark_pk_parse(f.curve, unsafe.Pointer(&pk[0]), C.uint64_t(len(pk)), &w.info)
	}); err != nil {
		return nil, fmt.Errorf("error reading PK data: %v", err)
	}
	// the constraint system is gnark's to decode (CBOR); it runs beside the GPU's point decompression
	ccsDone := make(chan error, 1)
	go func() {
		w.Ccs = plonk.NewCS(w.Curve)
		_, e := w.Ccs.ReadFrom(bytes.NewReader(data[f.ccs_off : f.ccs_off+f.ccs_len]))
		ccsDone <- e
	}()
	kzg := pk[w.info.kzg_off:w.info.lagrange_off]
	if err := call(func() C.int {
		return C.b2p_srs_load_compressed(f.curve, unsafe.Pointer(&kzg[0]), C.uint64_t(len(kzg)), w.info.kzg_count, &w.srs)
	}); err != nil {
		<-ccsDone
		return nil, fmt.Errorf("error reading PK data: %v", err)
	}
	if err := <-ccsDone; err != nil {
		C.b2p_srs_free(w.srs)
		return nil, fmt.Errorf("error reading CCS data: %v", err)
	}
	w.Vk = plonk.NewVerifyingKey(w.Curve)
	if _, err := w.Vk.ReadFrom(bytes.NewReader(data[f.vk_off : f.vk_off+f.vk_len])); err != nil {
		C.b2p_srs_free(w.srs)
		return nil, fmt.Errorf("error reading VK data: %v", err)
	}
	return w, nil
}

// snapshotPath is where the circuit half of a key (the arguments of b2p_circuit_load) is cached next to its file.
func snapshotPath(keyPath string) string { return keyPath + ".b2pk" }

// Circuit returns the device-resident circuit of the key: from the library's snapshot when it is at least as new
// as the key file (utils.ShouldRecompile's rule, utils/utils.go:68-86), else by building the trace once
// (uploadTrace in prove_<curve>.go) and writing the snapshot for the next start.
func (w *WarmKey) Circuit(keyPath string) (*C.b2p_circuit, error) {
	var c *C.b2p_circuit
	snap := snapshotPath(keyPath)
	if !shouldRecompile(snap, keyPath) {
		cs := C.CString(snap)
		defer C.free(unsafe.Pointer(cs))
		if err := call(func() C.int { return C.b2p_circuit_load_file(w.srs, cs, &c) }); err == nil {
			return c, nil
		}
	}
	// no usable snapshot: build the trace once, write the snapshot for the next start, make the circuit resident
	switch spr := w.Ccs.(type) {
	case *cs_bn254.SparseR1CS:
		return loadCircuitBN254(w.srs, spr, w.Vk.(*plonk_bn254.VerifyingKey), snap)
	default:
		return loadCircuitOtherCurves(w, snap) // prove_bls12381.go
	}
}

func shouldRecompile(target string, sources ...string) bool {
	t, err := os.Stat(target)
	if err != nil {
		return true
	}
	for _, s := range sources {
		si, err := os.Stat(s)
		if err != nil || si.ModTime().After(t.ModTime()) {
			return true
		}
	}
	return false
}
