package crypto

import (
	"sync"
	"time"
)

// VerifyBatcher turns the reference's one-credential-per-request calls (one goroutine per HTTP request,
// handlers/did_handlers.go:104) into batches: requests queue until BatchMax items are pending or LingerMicros have
// passed since the first one, then ONE VerifyBatch call serves them all.  cgo calls block an OS thread, so there is one
// flusher goroutine per backend rather than one cgo call per credential (SURVEY.md §8b "Threading").
type VerifyBatcher struct {
	V            Verifier
	BatchMax     int
	LingerMicros int

	mu      sync.Mutex
	pending []verifyReq
	timer   *time.Timer
}

type verifyReq struct {
	pk  [32]byte
	msg []byte
	sig [64]byte
	out chan bool
}

// Verify has the exact contract of ed25519.Verify (it panics on a bad public-key length, as Go does).
func (b *VerifyBatcher) Verify(pk, msg, sig []byte) bool {
	if len(pk) != 32 {
		panic("ed25519: bad public key length")
	}
	if len(sig) != 64 {
		return false
	}
	r := verifyReq{msg: msg, out: make(chan bool, 1)}
	copy(r.pk[:], pk)
	copy(r.sig[:], sig)
	b.mu.Lock()
	b.pending = append(b.pending, r)
	if len(b.pending) >= b.BatchMax {
		batch := b.take()
		b.mu.Unlock()
		b.flush(batch)
	} else {
		if b.timer == nil {
			b.timer = time.AfterFunc(time.Duration(b.LingerMicros)*time.Microsecond, func() {
				b.mu.Lock()
				batch := b.take()
				b.mu.Unlock()
				b.flush(batch)
			})
		}
		b.mu.Unlock()
	}
	return <-r.out
}

func (b *VerifyBatcher) take() []verifyReq {
	batch := b.pending
	b.pending = nil
	if b.timer != nil {
		b.timer.Stop()
		b.timer = nil
	}
	return batch
}

func (b *VerifyBatcher) flush(batch []verifyReq) {
	if len(batch) == 0 {
		return
	}
	pks := make([][32]byte, len(batch))
	sigs := make([][64]byte, len(batch))
	msgs := make([][]byte, len(batch))
	for i, r := range batch {
		pks[i], sigs[i], msgs[i] = r.pk, r.sig, r.msg
	}
	ok, err := b.V.VerifyBatch(pks, msgs, sigs)
	for i, r := range batch {
		r.out <- err == nil && ok[i]
	}
}
