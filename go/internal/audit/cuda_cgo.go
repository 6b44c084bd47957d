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
