package crypto

import "fmt"

// Config is the `features.did.crypto` block added to config.DIDConfig (internal/config/config.go:119-148), next to
// vc_requirements and keystore; defaults keep the reference's behaviour (stdlib, one credential per call).
//
//	features:
//	  did:
//	    crypto:
//	      backend: cuda            # cpu | cuda
//	      devices: [0, 1, 2, 3, 4, 5, 6, 7]
//	      batch_max: 8192          # flush a batcher at this many pending items ...
//	      batch_linger_us: 200     # ... or this long after the first one
//	      sign_constant_time: true # constant-time fixed-base multiplication for secret scalars (as crypto/ed25519)
//	      keycache_max_keys: 4096  # issuer-key cache capacity per GPU (384 KB per key)
type Config struct {
	Backend          string `yaml:"backend" default:"cpu"`
	Devices          []int  `yaml:"devices"`
	BatchMax         int    `yaml:"batch_max" default:"8192"`
	BatchLingerUs    int    `yaml:"batch_linger_us" default:"200"`
	SignConstantTime bool   `yaml:"sign_constant_time" default:"true"`
	KeyCacheMaxKeys  int    `yaml:"keycache_max_keys" default:"4096"`
}

// Open returns the backend the configuration names.  "cuda" without the cuda build tag, or without a usable GPU, is an error at
// start-up (the operator asked for it); once running, a failing batch falls back to the stdlib per call (cuda_cgo.go).
func Open(c Config) (Backend, error) {
	switch c.Backend {
	case "", "cpu", "stdlib":
		return Stdlib{}, nil
	case "cuda":
		devs := c.Devices
		if len(devs) == 0 {
			devs = []int{0}
		}
		return openCUDA(devs, c.SignConstantTime, c.KeyCacheMaxKeys)
	default:
		return nil, fmt.Errorf("crypto: unknown backend %q (want cpu or cuda)", c.Backend)
	}
}

// NewVerifyBatcher wires the batching seam in front of a backend with the configured flush rule.
func (c Config) NewVerifyBatcher(v Verifier) *VerifyBatcher {
	bm, lg := c.BatchMax, c.BatchLingerUs
	if bm <= 0 {
		bm = 8192
	}
	if lg <= 0 {
		lg = 200
	}
	return &VerifyBatcher{V: v, BatchMax: bm, LingerMicros: lg}
}
