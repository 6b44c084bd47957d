//go:build cgo && cuda

package crypto

/*
#cgo LDFLAGS: -lafcrypto
#include <stdlib.h>
#include "afcrypto.h"
*/
import "C"

import (
	"fmt"
	"sync/atomic"
	"unsafe"
)

// CUDA drives one afc_ctx per GPU and deals batches to them round-robin (SURVEY.md §8e: independent units, no
// collective).  On any non-zero return code the batch is recomputed with the stdlib backend and FallbackTotal is
// incremented — external behaviour never changes (mirrors vc_service.go:242-289 turning failures into Valid:false).
type CUDA struct {
	ctxs          []*C.afc_ctx
	next          atomic.Uint64
	FallbackTotal atomic.Uint64
	fallback      Stdlib
}

func NewCUDA(devices []int) (*CUDA, error) {
	c := &CUDA{}
	for _, d := range devices {
		var ctx *C.afc_ctx
		if rc := C.afc_init(C.int(d), &ctx); rc != 0 {
			c.Close()
			return nil, fmt.Errorf("afc_init(%d): %s", d, C.GoString(C.afc_strerror(rc)))
		}
		c.ctxs = append(c.ctxs, ctx)
	}
	if len(c.ctxs) == 0 {
		return nil, fmt.Errorf("crypto: no CUDA devices configured")
	}
	return c, nil
}

func (c *CUDA) Name() string { return "cuda" }
func (c *CUDA) Close() error {
	for _, x := range c.ctxs {
		C.afc_destroy(x)
	}
	c.ctxs = nil
	return nil
}
func (c *CUDA) pick() *C.afc_ctx { return c.ctxs[int(c.next.Add(1))%len(c.ctxs)] }

// pack copies [][]byte into one C buffer + offsets (cgo cannot pass Go pointers to Go pointers).  The buffers come
// from afc_alloc_pinned so the library's H2D copies are asynchronous.
func pack(msgs [][]byte) (buf unsafe.Pointer, off []C.uint64_t) {
	off = make([]C.uint64_t, len(msgs)+1)
	total := 0
	for i, m := range msgs {
		off[i] = C.uint64_t(total)
		total += len(m)
	}
	off[len(msgs)] = C.uint64_t(total)
	buf = C.afc_alloc_pinned(C.size_t(total + 1))
	dst := unsafe.Slice((*byte)(buf), total+1)
	for i, m := range msgs {
		copy(dst[off[i]:], m)
	}
	return
}

func (c *CUDA) VerifyBatch(pks [][32]byte, msgs [][]byte, sigs [][64]byte) ([]bool, error) {
	n := len(msgs)
	if len(pks) != n || len(sigs) != n {
		return nil, fmt.Errorf("crypto: mismatched batch lengths")
	}
	if n == 0 {
		return nil, nil
	}
	buf, off := pack(msgs)
	defer C.afc_free_pinned(buf)
	ok := make([]byte, n)
	rc := C.afc_ed25519_verify_batch(c.pick(), (*C.uint8_t)(unsafe.Pointer(&pks[0])), (*C.uint8_t)(unsafe.Pointer(&sigs[0])),
		(*C.uint8_t)(buf), &off[0], C.uint32_t(n), (*C.uint8_t)(unsafe.Pointer(&ok[0])))
	if rc != 0 {
		c.FallbackTotal.Add(1)
		return c.fallback.VerifyBatch(pks, msgs, sigs)
	}
	out := make([]bool, n)
	for i := range ok {
		out[i] = ok[i] == 1
	}
	return out, nil
}

func (c *CUDA) SignBatch(seeds [][32]byte, msgs [][]byte) ([][64]byte, error) {
	n := len(msgs)
	if len(seeds) != n {
		return nil, fmt.Errorf("crypto: %d seeds for %d messages", len(seeds), n)
	}
	if n == 0 {
		return nil, nil
	}
	buf, off := pack(msgs)
	defer C.afc_free_pinned(buf)
	out := make([][64]byte, n)
	rc := C.afc_ed25519_sign_batch(c.pick(), (*C.uint8_t)(unsafe.Pointer(&seeds[0])), (*C.uint8_t)(buf), &off[0], C.uint32_t(n),
		(*C.uint8_t)(unsafe.Pointer(&out[0])))
	if rc != 0 {
		c.FallbackTotal.Add(1)
		return c.fallback.SignBatch(seeds, msgs)
	}
	return out, nil
}

func (c *CUDA) HMACSHA256Batch(keys, msgs [][]byte) ([][32]byte, error) {
	n := len(msgs)
	if len(keys) != n {
		return nil, fmt.Errorf("crypto: mismatched batch lengths")
	}
	if n == 0 {
		return nil, nil
	}
	buf, off := pack(msgs)
	defer C.afc_free_pinned(buf)
	kbuf, koff64 := pack(keys)
	defer C.afc_free_pinned(kbuf)
	koff := make([]C.uint32_t, n+1)
	for i := range koff64 {
		koff[i] = C.uint32_t(koff64[i])
	}
	out := make([][32]byte, n)
	rc := C.afc_hmac_sha256_batch(c.pick(), (*C.uint8_t)(kbuf), &koff[0], (*C.uint8_t)(buf), &off[0], C.uint32_t(n),
		(*C.uint8_t)(unsafe.Pointer(&out[0])))
	if rc != 0 {
		c.FallbackTotal.Add(1)
		return c.fallback.HMACSHA256Batch(keys, msgs)
	}
	return out, nil
}

func (c *CUDA) SHA256Batch(msgs [][]byte) ([][32]byte, error) {
	n := len(msgs)
	if n == 0 {
		return nil, nil
	}
	buf, off := pack(msgs)
	defer C.afc_free_pinned(buf)
	out := make([][32]byte, n)
	rc := C.afc_sha256_batch(c.pick(), (*C.uint8_t)(buf), &off[0], C.uint32_t(n), (*C.uint8_t)(unsafe.Pointer(&out[0])))
	if rc != 0 {
		c.FallbackTotal.Add(1)
		return c.fallback.SHA256Batch(msgs)
	}
	return out, nil
}
