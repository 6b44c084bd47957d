"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by agentfield_b200/ (the product).

Big-integer restatement of Go 1.24 `crypto/ed25519` (toolchain go1.24.2, pinned by
/root/reference/control-plane/go.mod:3-5) exactly as the reference calls it:

  * ed25519.NewKeyFromSeed  control-plane/internal/services/vc_service.go:460,712
                            control-plane/internal/services/did_service.go:523
  * ed25519.Sign            control-plane/internal/services/vc_service.go:463,715
  * ed25519.Verify          control-plane/internal/services/vc_service.go:504,1624
                            control-plane/internal/cli/vc_verification_enhanced.go:453

The Go standard library is NOT under /root/reference and there is no Go toolchain here, so
this file restates the published algorithm (RFC 8032 §5.1 + the accept/reject rules of Go's
crypto/internal/fips140/ed25519 + edwards25519 packages):

  verify(pk, msg, sig):
    len(pk) != 32                      -> Go panics            (here: raises ValueError)
    len(sig) != 64 or sig[63] & 0xE0   -> False
    A = decompress(pk): bit 255 is the x sign; y = low 255 bits taken mod p WITHOUT a
        canonical check (y >= p accepted); x = sqrt((y^2-1)/(d y^2+1)) must exist else False;
        x == 0 with sign bit set is accepted (x stays 0); small-order A accepted
    S = sig[32:] must be canonical (< L) else False
    k = SHA-512(sig[:32] || pk || msg) mod L            (the 32 GIVEN pk bytes are hashed)
    R' = [S]B + [k](-A)        (cofactor-less)
    accept iff canonical_encoding(R') == sig[:32] byte-wise  (non-canonical R never matches)

Parity status: the reference holds no golden vector for this path, so the
oracle is pinned against RFC 8032 §7.1 vectors (tests/golden/rfc8032.json) and cross-checked
against two independent RFC 8032 implementations present in this image (OpenSSL 3.0.13 through
`cryptography`, libsodium through PyNaCl) on honest and randomly-corrupted inputs.
"""
import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)


def _sqrt_ratio(u, v):
    """Go field.Element.SqrtRatio: returns (r, was_square) with r the non-negative root."""
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    u = u % P
    if check == u:
        ok = True
    elif check == (-u) % P:
        r = r * SQRT_M1 % P
        ok = True
    elif check == (-u) * SQRT_M1 % P:
        r = r * SQRT_M1 % P
        ok = False
    else:
        ok = False
    if r & 1:           # Absolute(): pick the even ("non-negative") representative
        r = P - r
    return r, ok


def decompress(b32):
    """edwards25519.Point.SetBytes semantics.  Returns (x, y) affine or None."""
    if len(b32) != 32:
        raise ValueError("ed25519: bad public key length: %d" % len(b32))
    yi = int.from_bytes(b32, "little")
    sign = yi >> 255
    y = (yi & ((1 << 255) - 1)) % P          # non-canonical y accepted, reduced
    u = (y * y - 1) % P
    v = (D * y * y + 1) % P
    x, ok = _sqrt_ratio(u, v)
    if not ok:
        return None
    if sign:
        x = (-x) % P                          # x == 0 stays 0: accepted
    return x, y


def _add(p, q):
    x1, y1 = p
    x2, y2 = q
    t = D * x1 * x2 % P * y1 % P * y2 % P
    x3 = (x1 * y2 + x2 * y1) * pow(1 + t, P - 2, P) % P
    y3 = (y1 * y2 + x1 * x2) * pow(1 - t, P - 2, P) % P
    return x3, y3


# Projective (extended) arithmetic for speed: (X, Y, Z, T)
def _ext(p):
    return (p[0], p[1], 1, p[0] * p[1] % P)


def _ext_add(p, q):
    x1, y1, z1, t1 = p
    x2, y2, z2, t2 = q
    a = (y1 - x1) * (y2 - x2) % P
    b = (y1 + x1) * (y2 + x2) % P
    c = 2 * D * t1 % P * t2 % P
    d = 2 * z1 * z2 % P
    e, f, g, h = b - a, d - c, d + c, b + a
    return (e * f % P, g * h % P, f * g % P, e * h % P)


def _ext_mul(k, p):
    q = (0, 1, 1, 0)
    while k:
        if k & 1:
            q = _ext_add(q, p)
        p = _ext_add(p, p)
        k >>= 1
    return q


def _encode(pt):
    x, y, z, _ = pt
    zi = pow(z, P - 2, P)
    x, y = x * zi % P, y * zi % P
    return int(y | ((x & 1) << 255)).to_bytes(32, "little")


_BY = 4 * pow(5, P - 2, P) % P
_BX = decompress(int(_BY).to_bytes(32, "little"))[0]
BASE = (_BX, _BY)
_BASE_EXT = _ext(BASE)


def expand_seed(seed):
    """NewKeyFromSeed: returns (s, prefix, A_bytes).  s is the clamped scalar (not reduced)."""
    if len(seed) != 32:
        raise ValueError("ed25519: bad seed length: %d" % len(seed))
    h = hashlib.sha512(seed).digest()
    a = bytearray(h[:32])
    a[0] &= 248
    a[31] &= 63
    a[31] |= 64
    s = int.from_bytes(a, "little")
    return s, h[32:], _encode(_ext_mul(s % L, _BASE_EXT))


def public_key(seed):
    return expand_seed(seed)[2]


def sign(seed, msg):
    s, prefix, a_bytes = expand_seed(seed)
    r = int.from_bytes(hashlib.sha512(prefix + msg).digest(), "little") % L
    r_bytes = _encode(_ext_mul(r, _BASE_EXT))
    k = int.from_bytes(hashlib.sha512(r_bytes + a_bytes + msg).digest(), "little") % L
    big_s = (k * s + r) % L
    return r_bytes + int(big_s).to_bytes(32, "little")


def verify(pk, msg, sig):
    if len(pk) != 32:
        raise ValueError("ed25519: bad public key length: %d" % len(pk))   # Go panics
    if len(sig) != 64 or (sig[63] & 0xE0):
        return False
    a = decompress(pk)
    if a is None:
        return False
    big_s = int.from_bytes(sig[32:], "little")
    if big_s >= L:
        return False
    k = int.from_bytes(hashlib.sha512(sig[:32] + pk + msg).digest(), "little") % L
    minus_a = _ext(((-a[0]) % P, a[1]))
    rp = _ext_add(_ext_mul(big_s, _BASE_EXT), _ext_mul(k, minus_a))
    return _encode(rp) == sig[:32]


def small_order_points():
    """The 8 torsion points' canonical encodings (edge-case inputs for verify)."""
    out = []
    # order-8 point: y with (y^2-1)/(dy^2+1) square such that 8P = O; enumerate by brute force
    # on the known encodings (RFC 8032 / libsodium blocklist, canonical forms only).
    for hx in (
        "0100000000000000000000000000000000000000000000000000000000000000",  # order 1
        "ecffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",  # order 2
        "0000000000000000000000000000000000000000000000000000000000000000",  # order 4
        "0000000000000000000000000000000000000000000000000000000000000080",  # order 4
        "c7176a703d4dd84fba3c0b760d10670f2a2053fa2c39ccc64ec7fd7792ac037a",  # order 8
        "c7176a703d4dd84fba3c0b760d10670f2a2053fa2c39ccc64ec7fd7792ac03fa",  # order 8
        "26e8958fc2b227b045c3f489f2ef98f0d5dfac05d3c63339b13802886d53fc05",  # order 8
        "26e8958fc2b227b045c3f489f2ef98f0d5dfac05d3c63339b13802886d53fc85",  # order 8
    ):
        out.append(bytes.fromhex(hx))
    return out
