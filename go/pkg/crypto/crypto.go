// Package crypto defines the batch seams the B200 path plugs into and their stdlib implementation, which is
// byte-for-byte what the reference does today:
//
//	ed25519.NewKeyFromSeed + ed25519.Sign   internal/services/vc_service.go:460-463, 712-715
//	ed25519.Verify                          internal/services/vc_service.go:504, 1624; internal/cli/vc_verification_enhanced.go:453
//	hmac.New(sha256.New, secret)            internal/services/webhook_dispatcher.go:470-474
//	sha256.Sum256                           internal/services/vc_service.go:513; payload_store.go:69
package crypto

import (
	"crypto/ed25519"
	"crypto/hmac"
	"crypto/sha256"
	"fmt"
)

// Signer signs msgs[i] with the Ed25519 key derived from seeds[i] (RFC 8032, deterministic).
type Signer interface {
	SignBatch(seeds [][32]byte, msgs [][]byte) ([][64]byte, error)
}

// Verifier reports ed25519.Verify(pks[i], msgs[i], sigs[i]) for every i.  A false signature is not an error.
type Verifier interface {
	VerifyBatch(pks [][32]byte, msgs [][]byte, sigs [][64]byte) ([]bool, error)
}

// MAC computes HMAC-SHA256(keys[i], msgs[i]).
type MAC interface {
	HMACSHA256Batch(keys, msgs [][]byte) ([][32]byte, error)
}

// Hasher computes SHA-256(msgs[i]).
type Hasher interface {
	SHA256Batch(msgs [][]byte) ([][32]byte, error)
}

// Backend bundles the four seams; Name is reported in logs/metrics.
type Backend interface {
	Signer
	Verifier
	MAC
	Hasher
	Name() string
	Close() error
}

// Stdlib is the reference behaviour (default backend, and the fallback of the cuda backend).
type Stdlib struct{}

func (Stdlib) Name() string { return "stdlib" }
func (Stdlib) Close() error { return nil }

func (Stdlib) SignBatch(seeds [][32]byte, msgs [][]byte) ([][64]byte, error) {
	if len(seeds) != len(msgs) {
		return nil, fmt.Errorf("crypto: %d seeds for %d messages", len(seeds), len(msgs))
	}
	out := make([][64]byte, len(msgs))
	for i := range msgs {
		copy(out[i][:], ed25519.Sign(ed25519.NewKeyFromSeed(seeds[i][:]), msgs[i]))
	}
	return out, nil
}

func (Stdlib) VerifyBatch(pks [][32]byte, msgs [][]byte, sigs [][64]byte) ([]bool, error) {
	if len(pks) != len(msgs) || len(pks) != len(sigs) {
		return nil, fmt.Errorf("crypto: mismatched batch lengths")
	}
	out := make([]bool, len(msgs))
	for i := range msgs {
		out[i] = ed25519.Verify(ed25519.PublicKey(pks[i][:]), msgs[i], sigs[i][:])
	}
	return out, nil
}

func (Stdlib) HMACSHA256Batch(keys, msgs [][]byte) ([][32]byte, error) {
	if len(keys) != len(msgs) {
		return nil, fmt.Errorf("crypto: mismatched batch lengths")
	}
	out := make([][32]byte, len(msgs))
	for i := range msgs {
		m := hmac.New(sha256.New, keys[i])
		m.Write(msgs[i])
		copy(out[i][:], m.Sum(nil))
	}
	return out, nil
}

func (Stdlib) SHA256Batch(msgs [][]byte) ([][32]byte, error) {
	out := make([][32]byte, len(msgs))
	for i := range msgs {
		out[i] = sha256.Sum256(msgs[i])
	}
	return out, nil
}

// Default is what VCService / WebhookDispatcher call; main() replaces it with NewCUDA(...) when
// features.did.crypto_backend == "cuda".
var Default Backend = Stdlib{}
