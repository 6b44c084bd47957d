// Package identity caches expanded Ed25519 keys per DID so that issuing a VC costs ONE fixed-base scalar
// multiplication instead of the three the reference performs (NewKeyFromSeed in DIDService.derivePrivateKey on every
// ResolveDID — internal/services/did_service.go:515-525, 585-599 — NewKeyFromSeed again in signVC and then Sign —
// internal/services/vc_service.go:460-463).
package identity

import (
	"crypto/ed25519"
	"crypto/sha256"
	"sync"
)

// Expanded is the key material the signer needs: the 32-byte seed (from which the backend derives s and prefix) and
// the public key.
type Expanded struct {
	Seed [32]byte
	Pub  [32]byte
}

// Cache maps a DID string to its expanded key.  Derivation follows the reference exactly:
// seed' = SHA-256(masterSeed || derivationPath); key = ed25519.NewKeyFromSeed(seed').
type Cache struct {
	mu   sync.RWMutex
	keys map[string]Expanded
}

func NewCache() *Cache { return &Cache{keys: make(map[string]Expanded)} }

func Derive(masterSeed []byte, path string) Expanded {
	h := sha256.New()
	h.Write(masterSeed)
	h.Write([]byte(path))
	var e Expanded
	copy(e.Seed[:], h.Sum(nil))
	copy(e.Pub[:], ed25519.NewKeyFromSeed(e.Seed[:]).Public().(ed25519.PublicKey))
	return e
}

func (c *Cache) Get(did string) (Expanded, bool) {
	c.mu.RLock()
	e, ok := c.keys[did]
	c.mu.RUnlock()
	return e, ok
}

func (c *Cache) Put(did string, e Expanded) {
	c.mu.Lock()
	c.keys[did] = e
	c.mu.Unlock()
}
