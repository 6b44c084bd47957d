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
	"sync"
	"sync/atomic"
	"time"
	"unsafe"
)

// CUDA drives one afc_ctx per GPU and deals batches to them round-robin (SURVEY.md §8e: independent units, no
// collective).  On any non-zero return code the batch is recomputed with the stdlib backend and the fallback counter is
// incremented — external behaviour never changes (mirrors vc_service.go:242-289 turning failures into Valid:false).
type CUDA struct {
	devs     []*device
	next     atomic.Uint64
	fallback Stdlib
}

// device = one context + its pool of pinned packing buffers.  The buffers come from afc_alloc_pinned_for, i.e. they sit on
// the NUMA node of THAT GPU whichever goroutine's OS thread asks (one process, eight contexts: half of the GPUs hang off the
// other socket), and they are reused: cudaHostAlloc / cudaFreeHost cost hundreds of microseconds and serialise the process.
type device struct {
	ctx  *C.afc_ctx
	mu   sync.Mutex
	free map[int][]unsafe.Pointer // size class (power of two, bytes) -> idle buffers
	held int                      // bytes parked in `free`
}

const (
	minClass     = 1 << 16  // 64 KiB: smaller requests share this class
	maxPoolBytes = 1 << 30  // idle pinned memory kept per device; beyond it buffers are freed on return
)

func sizeClass(n int) int {
	c := minClass
	for c < n {
		c <<= 1
	}
	return c
}

func (d *device) get(n int) (unsafe.Pointer, int) {
	c := sizeClass(n)
	d.mu.Lock()
	if l := d.free[c]; len(l) > 0 {
		p := l[len(l)-1]
		d.free[c] = l[:len(l)-1]
		d.held -= c
		d.mu.Unlock()
		return p, c
	}
	d.mu.Unlock()
	return C.afc_alloc_pinned_for(d.ctx, C.size_t(c)), c
}

func (d *device) put(p unsafe.Pointer, c int) {
	if p == nil {
		return
	}
	d.mu.Lock()
	if d.held+c <= maxPoolBytes {
		d.free[c] = append(d.free[c], p)
		d.held += c
		d.mu.Unlock()
		return
	}
	d.mu.Unlock()
	C.afc_free_pinned(p)
}

// NewCUDA opens one context per device.  signConstantTime selects the constant-time fixed-base multiplication for secret
// scalars (what Go's crypto/ed25519 does; the default of the library); keyCacheMaxKeys sizes the issuer-key cache (0 = library default).
func NewCUDA(devices []int, signConstantTime bool, keyCacheMaxKeys int) (*CUDA, error) {
	c := &CUDA{}
	for _, id := range devices {
		var ctx *C.afc_ctx
		if rc := C.afc_init(C.int(id), &ctx); rc != 0 {
			c.Close()
			return nil, fmt.Errorf("afc_init(%d): %s", id, C.GoString(C.afc_strerror(rc)))
		}
		ct := C.int(0)
		if signConstantTime {
			ct = 1
		}
		C.afc_sign_configure(ctx, ct)
		if keyCacheMaxKeys > 0 {
			C.afc_keycache_configure(ctx, C.uint32_t(keyCacheMaxKeys))
		}
		c.devs = append(c.devs, &device{ctx: ctx, free: map[int][]unsafe.Pointer{}})
	}
	if len(c.devs) == 0 {
		return nil, fmt.Errorf("crypto: no CUDA devices configured")
	}
	return c, nil
}

func (c *CUDA) Name() string { return "cuda" }
func (c *CUDA) Close() error {
	for _, d := range c.devs {
		for _, l := range d.free {
			for _, p := range l {
				C.afc_free_pinned(p)
			}
		}
		C.afc_destroy(d.ctx)
	}
	c.devs = nil
	return nil
}
func (c *CUDA) pick() *device { return c.devs[int(c.next.Add(1))%len(c.devs)] }

// packed is one batch in the layout of include/afcrypto.h inside ONE pooled pinned buffer:
// [offsets (n+1) x 8][fixed-size records a][fixed-size records b][messages][results]
type packed struct {
	d           *device
	p           unsafe.Pointer
	class       int
	off         *C.uint64_t
	a, b        *C.uint8_t
	msgs        *C.uint8_t
	out         unsafe.Pointer
	outBytes    int
}

func (k *packed) release() { k.d.put(k.p, k.class) }

func align16(n int) int { return (n + 15) &^ 15 }

// pack lays a batch out in pinned memory: cgo cannot pass [][]byte, and from pinned memory the library's H2D copies are
// asynchronous and need no bounce buffer.  aItem / bItem: bytes per item of the two fixed-size inputs (0 = absent).
func (d *device) pack(msgs [][]byte, aItem, bItem, outItem int) *packed {
	n := len(msgs)
	total := 0
	for _, m := range msgs {
		total += len(m)
	}
	oOff, oA := 0, align16((n+1)*8)
	oB := oA + align16(n*aItem)
	oM := oB + align16(n*bItem)
	oOut := oM + align16(total+1)
	size := oOut + align16(n*outItem+1) // +1: &base[oOut] must exist even when the call has no fixed-size output (the HMAC key pack)
	p, class := d.get(size)
	if p == nil {
		return nil
	}
	base := unsafe.Slice((*byte)(p), size)
	off := unsafe.Slice((*uint64)(unsafe.Pointer(&base[oOff])), n+1)
	pos := 0
	for i, m := range msgs {
		off[i] = uint64(pos)
		copy(base[oM+pos:], m)
		pos += len(m)
	}
	off[n] = uint64(pos)
	return &packed{d: d, p: p, class: class, off: (*C.uint64_t)(unsafe.Pointer(&base[oOff])), a: (*C.uint8_t)(unsafe.Pointer(&base[oA])),
		b: (*C.uint8_t)(unsafe.Pointer(&base[oB])), msgs: (*C.uint8_t)(unsafe.Pointer(&base[oM])), out: unsafe.Pointer(&base[oOut]), outBytes: n * outItem}
}

func (c *CUDA) VerifyBatch(pks [][32]byte, msgs [][]byte, sigs [][64]byte) ([]bool, error) {
	n := len(msgs)
	if len(pks) != n || len(sigs) != n {
		return nil, fmt.Errorf("crypto: mismatched batch lengths")
	}
	if n == 0 {
		return nil, nil
	}
	t0 := time.Now()
	d := c.pick()
	k := d.pack(msgs, 32, 64, 1)
	if k == nil {
		observeFallback("verify")
		return c.fallback.VerifyBatch(pks, msgs, sigs)
	}
	defer k.release()
	copy(unsafe.Slice((*[32]byte)(unsafe.Pointer(k.a)), n), pks)
	copy(unsafe.Slice((*[64]byte)(unsafe.Pointer(k.b)), n), sigs)
	rc := C.afc_ed25519_verify_batch(d.ctx, k.a, k.b, k.msgs, k.off, C.uint32_t(n), (*C.uint8_t)(k.out))
	if rc != 0 {
		observeFallback("verify")
		return c.fallback.VerifyBatch(pks, msgs, sigs)
	}
	ok := unsafe.Slice((*byte)(k.out), n)
	out := make([]bool, n)
	for i := range out {
		out[i] = ok[i] == 1
	}
	observeBatch("verify", n, time.Since(t0))
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
	t0 := time.Now()
	d := c.pick()
	k := d.pack(msgs, 32, 0, 64)
	if k == nil {
		observeFallback("sign")
		return c.fallback.SignBatch(seeds, msgs)
	}
	defer k.release()
	sd := unsafe.Slice((*[32]byte)(unsafe.Pointer(k.a)), n)
	copy(sd, seeds)
	rc := C.afc_ed25519_sign_batch(d.ctx, k.a, k.msgs, k.off, C.uint32_t(n), (*C.uint8_t)(k.out))
	for i := range sd { // key material does not stay behind in a pooled buffer
		sd[i] = [32]byte{}
	}
	if rc != 0 {
		observeFallback("sign")
		return c.fallback.SignBatch(seeds, msgs)
	}
	out := make([][64]byte, n)
	copy(out, unsafe.Slice((*[64]byte)(k.out), n))
	observeBatch("sign", n, time.Since(t0))
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
	t0 := time.Now()
	d := c.pick()
	k := d.pack(msgs, 0, 0, 32)
	kk := d.pack(keys, 0, 0, 0)
	if k == nil || kk == nil {
		if k != nil {
			k.release()
		}
		if kk != nil {
			kk.release()
		}
		observeFallback("hmac")
		return c.fallback.HMACSHA256Batch(keys, msgs)
	}
	defer k.release()
	defer kk.release()
	koff64 := unsafe.Slice((*uint64)(unsafe.Pointer(kk.off)), n+1)
	koff := make([]C.uint32_t, n+1)
	for i := range koff {
		koff[i] = C.uint32_t(koff64[i])
	}
	rc := C.afc_hmac_sha256_batch(d.ctx, kk.msgs, &koff[0], k.msgs, k.off, C.uint32_t(n), (*C.uint8_t)(k.out))
	if rc != 0 {
		observeFallback("hmac")
		return c.fallback.HMACSHA256Batch(keys, msgs)
	}
	out := make([][32]byte, n)
	copy(out, unsafe.Slice((*[32]byte)(k.out), n))
	observeBatch("hmac", n, time.Since(t0))
	return out, nil
}

func (c *CUDA) SHA256Batch(msgs [][]byte) ([][32]byte, error) {
	n := len(msgs)
	if n == 0 {
		return nil, nil
	}
	t0 := time.Now()
	d := c.pick()
	k := d.pack(msgs, 0, 0, 32)
	if k == nil {
		observeFallback("sha256")
		return c.fallback.SHA256Batch(msgs)
	}
	defer k.release()
	rc := C.afc_sha256_batch(d.ctx, k.msgs, k.off, C.uint32_t(n), (*C.uint8_t)(k.out))
	if rc != 0 {
		observeFallback("sha256")
		return c.fallback.SHA256Batch(msgs)
	}
	out := make([][32]byte, n)
	copy(out, unsafe.Slice((*[32]byte)(k.out), n))
	observeBatch("sha256", n, time.Since(t0))
	return out, nil
}

func openCUDA(devices []int, signConstantTime bool, keyCacheMaxKeys int) (Backend, error) {
	return NewCUDA(devices, signConstantTime, keyCacheMaxKeys)
}
