"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/_build/libafc_oracle.so and
libafc_openssl.so (built by oracle/Makefile; see afc_oracle.h for the reference citations)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)


def build(force=False):
    so = os.path.join(_HERE, "_build", "libafc_oracle.so")
    so2 = os.path.join(_HERE, "_build", "libafc_openssl.so")
    srcs = [os.path.join(_HERE, f) for f in ("afc_oracle.c", "afc_oracle.h", "afc_consts.inc", "afc_openssl.c")]
    stale = force or not (os.path.exists(so) and os.path.exists(so2)) or \
        min(os.path.getmtime(so), os.path.getmtime(so2)) < max(os.path.getmtime(s) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return so, so2


_lib = _ossl = None


def _ptr(a, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib, _ossl
    if _lib is None:
        so, so2 = build()
        _lib = C.CDLL(so)
        _ossl = C.CDLL(so2)
    return _lib


def ossl():
    lib()
    return _ossl


def _pack(msgs):
    off = np.zeros(len(msgs) + 1, dtype=np.uint64)
    if len(msgs):
        off[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8).copy()
    return buf, off


def sha256_batch(buf, off, nthreads=1, impl="oracle"):
    n = len(off) - 1
    out = np.zeros((n, 32), dtype=np.uint8)
    f = lib().afo_sha256_batch if impl == "oracle" else ossl().afx_sha256_batch
    f(_ptr(buf), _ptr(off, C.c_uint64), C.c_uint32(n), _ptr(out), C.c_int(nthreads))
    return out


def hmac_sha256_batch(keys, koff, buf, off, nthreads=1, impl="oracle"):
    n = len(off) - 1
    out = np.zeros((n, 32), dtype=np.uint8)
    f = lib().afo_hmac_sha256_batch if impl == "oracle" else ossl().afx_hmac_sha256_batch
    f(_ptr(keys), _ptr(koff, C.c_uint32), _ptr(buf), _ptr(off, C.c_uint64), C.c_uint32(n), _ptr(out), C.c_int(nthreads))
    return out


def ed25519_verify_batch(pks, sigs, buf, off, nthreads=1, impl="oracle"):
    n = len(off) - 1
    out = np.zeros(n, dtype=np.uint8)
    f = lib().afo_ed25519_verify_batch if impl == "oracle" else ossl().afx_ed25519_verify_batch
    f(_ptr(pks), _ptr(sigs), _ptr(buf), _ptr(off, C.c_uint64), C.c_uint32(n), _ptr(out), C.c_int(nthreads))
    return out


def ed25519_sign_batch(seeds, buf, off, nthreads=1, impl="oracle"):
    n = len(off) - 1
    out = np.zeros((n, 64), dtype=np.uint8)
    f = lib().afo_ed25519_sign_batch if impl == "oracle" else ossl().afx_ed25519_sign_batch
    f(_ptr(seeds), _ptr(buf), _ptr(off, C.c_uint64), C.c_uint32(n), _ptr(out), C.c_int(nthreads))
    return out


def ed25519_pubkey_batch(seeds, nthreads=1):
    n = seeds.shape[0]
    out = np.zeros((n, 32), dtype=np.uint8)
    lib().afo_ed25519_pubkey_batch(_ptr(seeds), C.c_uint32(n), _ptr(out), C.c_int(nthreads))
    return out


def merkle_root(buf, off, nthreads=1):
    n = len(off) - 1
    out = np.zeros(32, dtype=np.uint8)
    lib().afo_merkle_root(_ptr(buf), _ptr(off, C.c_uint64), C.c_uint32(n), _ptr(out), C.c_int(nthreads))
    return out.tobytes()


def merkle_root_from_hashes(hashes):
    hashes = np.ascontiguousarray(hashes, dtype=np.uint8)
    n = hashes.shape[0] if hashes.ndim == 2 else hashes.size // 32
    out = np.zeros(32, dtype=np.uint8)
    lib().afo_merkle_root_from_hashes(_ptr(hashes), C.c_uint64(n), _ptr(out))
    return out.tobytes()


# single-item helpers (bytes in / bytes out)
def sha256(m):
    b, o = _pack([m]); return sha256_batch(b, o)[0].tobytes()


def sha512(m):
    out = (C.c_uint8 * 64)()
    lib().afo_sha512(C.c_char_p(m), C.c_size_t(len(m)), out)
    return bytes(out)


def hmac_sha256(k, m):
    out = (C.c_uint8 * 32)()
    lib().afo_hmac_sha256(C.c_char_p(k), C.c_size_t(len(k)), C.c_char_p(m), C.c_size_t(len(m)), out)
    return bytes(out)


def pubkey(seed):
    out = (C.c_uint8 * 32)()
    lib().afo_ed25519_pubkey(C.c_char_p(seed), out)
    return bytes(out)


def sign(seed, m):
    out = (C.c_uint8 * 64)()
    lib().afo_ed25519_sign(C.c_char_p(seed), C.c_char_p(m), C.c_size_t(len(m)), out)
    return bytes(out)


def verify(pk, m, sig):
    if len(pk) != 32:
        raise ValueError("ed25519: bad public key length: %d" % len(pk))
    if len(sig) != 64:
        return False
    return bool(lib().afo_ed25519_verify(C.c_char_p(pk), C.c_char_p(m), C.c_size_t(len(m)), C.c_char_p(sig)))
