// Go CPU baseline for the hot path (SOURCE ONLY here: the build image has no Go toolchain; run it wherever Go 1.24
// exists with `go test -bench . -benchtime 5s` and do not invent its numbers).  Same shapes as bench.py: 512-byte
// messages, 1024 keys, 256-byte HMAC bodies.  These are the stdlib calls the reference makes at
// internal/services/vc_service.go:460-463,504 and internal/services/webhook_dispatcher.go:470-474.
package baseline

import (
	"crypto/ed25519"
	"crypto/hmac"
	"crypto/rand"
	"crypto/sha256"
	"testing"
)

func BenchmarkEd25519Verify512(b *testing.B) {
	pub, priv, _ := ed25519.GenerateKey(rand.Reader)
	msg := make([]byte, 512)
	rand.Read(msg)
	sig := ed25519.Sign(priv, msg)
	b.SetBytes(609)
	b.RunParallel(func(pb *testing.PB) {
		for pb.Next() {
			if !ed25519.Verify(pub, msg, sig) {
				b.Fatal("verify failed")
			}
		}
	})
}

func BenchmarkEd25519SignFromSeed512(b *testing.B) {
	seed := make([]byte, 32)
	rand.Read(seed)
	msg := make([]byte, 512)
	rand.Read(msg)
	b.RunParallel(func(pb *testing.PB) {
		for pb.Next() {
			ed25519.Sign(ed25519.NewKeyFromSeed(seed), msg) // what signVC does per VC
		}
	})
}

func BenchmarkHMACSHA256_256(b *testing.B) {
	key := make([]byte, 32)
	body := make([]byte, 256)
	rand.Read(key)
	rand.Read(body)
	b.SetBytes(320)
	b.RunParallel(func(pb *testing.PB) {
		for pb.Next() {
			m := hmac.New(sha256.New, key)
			m.Write(body)
			m.Sum(nil)
		}
	})
}

func BenchmarkSHA256_512(b *testing.B) {
	msg := make([]byte, 512)
	rand.Read(msg)
	b.SetBytes(512)
	b.RunParallel(func(pb *testing.PB) {
		for pb.Next() {
			sha256.Sum256(msg)
		}
	})
}
