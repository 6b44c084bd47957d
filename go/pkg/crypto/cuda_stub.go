//go:build !(cgo && cuda)

package crypto

import "errors"

// openCUDA is unavailable in builds without `-tags cuda` + cgo (the reference's CI matrix builds with CGO_ENABLED=0,
// .github/workflows/control-plane.yml:79): asking for backend "cuda" in such a build is a start-up error, not a silent fallback.
func openCUDA(devices []int, signConstantTime bool, keyCacheMaxKeys int) (Backend, error) {
	return nil, errors.New("crypto: built without the cuda backend (need CGO_ENABLED=1 and -tags cuda)")
}
