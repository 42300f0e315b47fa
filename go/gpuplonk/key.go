// key.go -- what the two curve files share: the device-resident key cache (with eviction), the cgo call wrapper
// that keeps a return code and its thread-local message on ONE OS thread, the page-locked column pool and the
// pointer-array helper that respects the cgo pointer rules.  NOT COMPILED in the build container (no Go toolchain).
package gpuplonk

/*
#include <stdlib.h>
#include "b200plonk.h"
*/
import "C"

import (
	"container/list"
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// gpuKey is one proving key resident in HBM: SRS table, selector / permutation columns, prover workspace, and three
// page-locked host columns that every proof of this key reuses (a proof holds key.mu, so one set is enough).
type gpuKey struct {
	mu      sync.Mutex // a handle takes one call at a time (b200plonk.h); concurrent provers of one key queue here
	srs     *C.b2p_srs
	circuit *C.b2p_circuit
	cols    [3]unsafe.Pointer // pinned, n * 32 bytes each (b2p_host_alloc); nil = pageable fallback
	n       int
	elem    *list.Element // position in the LRU list
	owner   any           // the gnark proving key (map key), for eviction
}

// MaxResidentKeys bounds how many proving keys stay resident in HBM (a 2^20-row BN254 key is ~5.5 GB: SRS table,
// coset evaluations of the selectors, workspace).  The least recently used key is freed when the bound is
// exceeded; Free(pk) drops one explicitly.
var MaxResidentKeys = 8

var (
	mu   sync.Mutex
	keys = map[any]*gpuKey{} // *plonk_bn254.ProvingKey / *plonk_bls12381.ProvingKey -> resident key
	lru  = list.New()        // front = most recently used
)

// call runs one library call and, if it failed, reads b2p_last_error() on the SAME OS thread: the message is
// thread local on the C side and goroutines migrate between threads.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != 0 {
		return fmt.Errorf("b200plonk error %d: %s", int(rc), C.GoString(C.b2p_last_error()))
	}
	return nil
}

// lookup returns the resident key of pk (and marks it most recently used), or nil.
func lookup(pk any) *gpuKey {
	mu.Lock()
	defer mu.Unlock()
	k := keys[pk]
	if k != nil {
		lru.MoveToFront(k.elem)
	}
	return k
}

// remember registers a freshly uploaded key and evicts the least recently used ones beyond MaxResidentKeys.
func remember(pk any, k *gpuKey) {
	mu.Lock()
	k.owner = pk
	k.elem = lru.PushFront(k)
	keys[pk] = k
	var evict []*gpuKey
	for lru.Len() > MaxResidentKeys && MaxResidentKeys > 0 {
		old := lru.Remove(lru.Back()).(*gpuKey)
		delete(keys, old.owner)
		evict = append(evict, old)
	}
	mu.Unlock()
	for _, old := range evict {
		old.free()
	}
}

// Free releases the device-resident copy of a proving key (plonk.ProvingKey of either curve); the next Prove with
// it uploads it again.
func Free(pk any) {
	mu.Lock()
	k := keys[pk]
	if k != nil {
		lru.Remove(k.elem)
		delete(keys, pk)
	}
	mu.Unlock()
	if k != nil {
		k.free()
	}
}

func (k *gpuKey) free() {
	k.mu.Lock() // wait for a proof in flight on this key
	defer k.mu.Unlock()
	if k.circuit != nil {
		C.b2p_circuit_free(k.circuit)
		k.circuit = nil
	}
	if k.srs != nil {
		C.b2p_srs_free(k.srs)
		k.srs = nil
	}
	for i, p := range k.cols {
		if p != nil {
			C.b2p_host_free(p)
			k.cols[i] = nil
		}
	}
}

// allocColumns gives the key its three page-locked columns (n elements of 32 bytes).  Pinned memory uploads at
// PCIe speed and overlaps with the first transforms; if the allocation fails the proofs fall back to pageable
// slices (correct, slower upload).
func (k *gpuKey) allocColumns(n int) {
	k.n = n
	for i := range k.cols {
		var p unsafe.Pointer
		if C.b2p_host_alloc(C.uint64_t(n)*32, &p) != 0 {
			p = nil
		}
		k.cols[i] = p
	}
}

// pointerArray passes Go pointers to C the way the cgo rules allow: every element is pinned for the duration of
// the call (runtime.Pinner, Go >= 1.21) and the array itself is Go memory handed over for that call only.  The
// returned function unpins; call it after the C call returned.
func pointerArray(ptrs []unsafe.Pointer) (**byte, func()) {
	if len(ptrs) == 0 {
		return nil, func() {}
	}
	var pin runtime.Pinner
	for _, p := range ptrs {
		pin.Pin(p)
	}
	arr := make([]unsafe.Pointer, len(ptrs))
	copy(arr, ptrs)
	pin.Pin(&arr[0])
	return (**byte)(unsafe.Pointer(&arr[0])), pin.Unpin
}
