// Package audit adds the tamper-evident log the reference lacks (its chain check is a stub that returns true,
// internal/cli/vc_verification_enhanced.go:531-534): an RFC 6962 §2.1 Merkle tree over the stored vc_document bytes
// (internal/services/vc_service.go:201,218).  leaf hash = SHA-256(0x00 || leaf); node = SHA-256(0x01 || l || r).
package audit

import (
	"crypto/sha256"
	"sync"
)

// Auditor appends leaves and reports the current root and size.
type Auditor interface {
	Append(leaves [][]byte) (root [32]byte, size uint64, err error)
	Root() (root [32]byte, size uint64, err error)
}

// Stdlib keeps the compact frontier (one complete-subtree hash per set bit of size) — the same state the cuda backend
// keeps on the device (afc_merkle_save / afc_merkle_load interchange it).
type Stdlib struct {
	mu       sync.Mutex
	size     uint64
	frontier [64][32]byte
}

func leafHash(d []byte) [32]byte {
	h := sha256.New()
	h.Write([]byte{0})
	h.Write(d)
	var o [32]byte
	copy(o[:], h.Sum(nil))
	return o
}

func nodeHash(l, r [32]byte) [32]byte {
	var b [65]byte
	b[0] = 1
	copy(b[1:], l[:])
	copy(b[33:], r[:])
	return sha256.Sum256(b[:])
}

func (a *Stdlib) Append(leaves [][]byte) ([32]byte, uint64, error) {
	a.mu.Lock()
	defer a.mu.Unlock()
	for _, d := range leaves {
		cur, h := leafHash(d), 0
		for (a.size>>uint(h))&1 == 1 {
			cur = nodeHash(a.frontier[h], cur)
			h++
		}
		a.frontier[h] = cur
		a.size++
	}
	r, s := a.rootLocked()
	return r, s, nil
}

func (a *Stdlib) Root() ([32]byte, uint64, error) {
	a.mu.Lock()
	defer a.mu.Unlock()
	r, s := a.rootLocked()
	return r, s, nil
}

func (a *Stdlib) rootLocked() ([32]byte, uint64) {
	if a.size == 0 {
		return sha256.Sum256(nil), 0
	}
	var acc [32]byte
	first := true
	for h := 0; h < 64; h++ {
		if (a.size>>uint(h))&1 == 1 {
			if first {
				acc, first = a.frontier[h], false
			} else {
				acc = nodeHash(a.frontier[h], acc)
			}
		}
	}
	return acc, a.size
}
