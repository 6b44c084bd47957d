//go:build !(cgo && cuda)

package crypto

import "errors"

// NewCUDA is unavailable in builds without `-tags cuda` + cgo; callers keep the stdlib backend.
func NewCUDA(devices []int) (Backend, error) {
	return nil, errors.New("crypto: built without the cuda backend (need CGO_ENABLED=1 and -tags cuda)")
}
