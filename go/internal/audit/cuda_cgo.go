//go:build cgo && cuda

package audit

/*
#cgo LDFLAGS: -lafcrypto
#include "afcrypto.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// CUDA keeps the log state on one GPU (afc_merkle).  For a multi-GPU box each device appends its contiguous,
// 2^k-aligned leaf range and the 32-byte subtree roots are exchanged with afc_comm_allgather_roots (NCCL over NVLink)
// and folded with afc_merkle_append_hashes — SURVEY.md §8e.
type CUDA struct {
	m *C.afc_merkle
}

func NewCUDA(ctx unsafe.Pointer) (*CUDA, error) {
	var m *C.afc_merkle
	if rc := C.afc_merkle_new((*C.afc_ctx)(ctx), &m); rc != 0 {
		return nil, fmt.Errorf("afc_merkle_new: %s", C.GoString(C.afc_strerror(rc)))
	}
	return &CUDA{m: m}, nil
}

func (a *CUDA) Close() { C.afc_merkle_free(a.m) }

func (a *CUDA) Append(leaves [][]byte) (root [32]byte, size uint64, err error) {
	off := make([]C.uint64_t, len(leaves)+1)
	total := 0
	for i, l := range leaves {
		off[i] = C.uint64_t(total)
		total += len(l)
	}
	off[len(leaves)] = C.uint64_t(total)
	buf := C.afc_alloc_pinned(C.size_t(total + 1))
	defer C.afc_free_pinned(buf)
	dst := unsafe.Slice((*byte)(buf), total+1)
	for i, l := range leaves {
		copy(dst[off[i]:], l)
	}
	var sz C.uint64_t
	if rc := C.afc_merkle_append(a.m, (*C.uint8_t)(buf), &off[0], C.uint32_t(len(leaves)), (*C.uint8_t)(unsafe.Pointer(&root[0])), &sz); rc != 0 {
		return root, 0, fmt.Errorf("afc_merkle_append: %s", C.GoString(C.afc_strerror(rc)))
	}
	return root, uint64(sz), nil
}

func (a *CUDA) Root() (root [32]byte, size uint64, err error) {
	var sz C.uint64_t
	if rc := C.afc_merkle_root(a.m, (*C.uint8_t)(unsafe.Pointer(&root[0])), &sz); rc != 0 {
		return root, 0, fmt.Errorf("afc_merkle_root: %s", C.GoString(C.afc_strerror(rc)))
	}
	return root, uint64(sz), nil
}

// Tree materialises every level over a fixed set of leaf hashes (afc_merkle_tree) so that audit paths and
// old-root -> new-root consistency proofs can be read out for an export bundle (internal/handlers/ui/did.go:533-865).
type Tree struct {
	t     *C.afc_merkle_tree
	depth uint32
}

func NewTree(ctx unsafe.Pointer, leafHashes [][32]byte) (*Tree, error) {
	var t *C.afc_merkle_tree
	var p *C.uint8_t
	if len(leafHashes) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&leafHashes[0]))
	}
	if rc := C.afc_merkle_tree_build((*C.afc_ctx)(ctx), p, C.uint64_t(len(leafHashes)), &t); rc != 0 {
		return nil, fmt.Errorf("afc_merkle_tree_build: %s", C.GoString(C.afc_strerror(rc)))
	}
	var depth C.uint32_t
	C.afc_merkle_tree_root(t, nil, nil, &depth)
	return &Tree{t: t, depth: uint32(depth)}, nil
}

func (t *Tree) Close() { C.afc_merkle_tree_free(t.t) }

// ConsistencyProof returns RFC 6962 §2.1.2 PROOF(first, D[n]).
func (t *Tree) ConsistencyProof(first uint64) ([][32]byte, error) {
	buf := make([][32]byte, 2*t.depth+2)
	var n C.uint32_t
	if rc := C.afc_merkle_tree_consistency_proof(t.t, C.uint64_t(first), (*C.uint8_t)(unsafe.Pointer(&buf[0])), &n); rc != 0 {
		return nil, fmt.Errorf("afc_merkle_tree_consistency_proof: %s", C.GoString(C.afc_strerror(rc)))
	}
	return buf[:n], nil
}

// VerifyConsistency checks many earlier checkpoints against one later (size, root) on the device (RFC 9162 §2.1.4.2);
// this is what `af vc verify` calls in place of the checkChainIntegrity stub (internal/cli/vc_verification_enhanced.go:531-534).
func VerifyConsistency(ctx unsafe.Pointer, firstSizes []uint64, firstRoots [][32]byte, size uint64, root [32]byte, proofs [][][32]byte) ([]bool, error) {
	m := len(firstSizes)
	if m == 0 {
		return nil, nil
	}
	off := make([]C.uint32_t, m+1)
	var flat [][32]byte
	for i, p := range proofs {
		off[i] = C.uint32_t(len(flat))
		flat = append(flat, p...)
	}
	off[m] = C.uint32_t(len(flat))
	if len(flat) == 0 {
		flat = make([][32]byte, 1)
	}
	ok := make([]byte, m)
	rc := C.afc_merkle_verify_consistency_batch((*C.afc_ctx)(ctx), (*C.uint64_t)(unsafe.Pointer(&firstSizes[0])), (*C.uint8_t)(unsafe.Pointer(&firstRoots[0])),
		C.uint64_t(size), (*C.uint8_t)(unsafe.Pointer(&root[0])), (*C.uint8_t)(unsafe.Pointer(&flat[0])), &off[0], C.uint32_t(m), (*C.uint8_t)(unsafe.Pointer(&ok[0])))
	if rc != 0 {
		return nil, fmt.Errorf("afc_merkle_verify_consistency_batch: %s", C.GoString(C.afc_strerror(rc)))
	}
	out := make([]bool, m)
	for i := range ok {
		out[i] = ok[i] != 0
	}
	return out, nil
}
