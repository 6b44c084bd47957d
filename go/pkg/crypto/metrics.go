package crypto

import (
	"time"

	"github.com/prometheus/client_golang/prometheus"
	"github.com/prometheus/client_golang/prometheus/promauto"
)

// Prometheus series of the crypto backend, registered the way the reference registers its gateway metrics
// (internal/services/execution_metrics.go:13-45, served at /metrics by internal/server/server.go:607).
var (
	cryptoBatchSize = promauto.NewHistogramVec(prometheus.HistogramOpts{
		Name:    "agentfield_crypto_batch_size",
		Help:    "Credentials / messages per batch handed to the crypto backend, by operation.",
		Buckets: prometheus.ExponentialBuckets(1, 4, 12), // 1 ... 4 M
	}, []string{"op"})

	cryptoGPUSeconds = promauto.NewCounterVec(prometheus.CounterOpts{
		Name: "agentfield_crypto_gpu_seconds",
		Help: "Wall time spent inside accelerated crypto batch calls (pack + copy + kernels + unpack), by operation.",
	}, []string{"op"})

	cryptoFallbackTotal = promauto.NewCounterVec(prometheus.CounterOpts{
		Name: "agentfield_crypto_fallback_total",
		Help: "Batches recomputed with the Go standard library because the accelerated backend returned an error, by operation.",
	}, []string{"op"})
)

func observeBatch(op string, n int, d time.Duration) {
	cryptoBatchSize.WithLabelValues(op).Observe(float64(n))
	cryptoGPUSeconds.WithLabelValues(op).Add(d.Seconds())
}

func observeFallback(op string) { cryptoFallbackTotal.WithLabelValues(op).Inc() }
