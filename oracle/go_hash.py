"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by agentfield_b200/ (the product).

CPU restatement of the reference's SHA-256 / HMAC-SHA256 call sites.  The arithmetic is Go's
crypto/sha256 + crypto/hmac (FIPS 180-4, RFC 2104), restated here through CPython's hashlib/hmac
(OpenSSL) which are pinned by tests/golden/{fips180,rfc4231}.json.

  hash_data            control-plane/internal/services/vc_service.go:508-515  (sha256 -> base64url no pad)
  marshal_data_or_null control-plane/internal/services/vc_service.go:1298-1306
  webhook_signature    control-plane/internal/services/webhook_dispatcher.go:470-474
  derive_seed          control-plane/internal/services/did_service.go:515-525
  did_key              control-plane/internal/services/did_service.go:528-536
  payload sha256 hex   control-plane/internal/services/payload_store.go:69-94
"""
import base64
import hashlib
import hmac as _hmac


def sha256(data: bytes) -> bytes:
    return hashlib.sha256(data).digest()


def sha512(data: bytes) -> bytes:
    return hashlib.sha512(data).digest()


def hmac_sha256(key: bytes, msg: bytes) -> bytes:
    return _hmac.new(key, msg, hashlib.sha256).digest()


def b64url_nopad(b: bytes) -> str:
    return base64.urlsafe_b64encode(b).rstrip(b"=").decode()


def marshal_data_or_null(data):
    """json.Marshal([]byte) -> '"<std base64>"'; nil -> 'null' (vc_service.go:1298-1306)."""
    if data is None:
        return b"null"
    return b'"' + base64.b64encode(data) + b'"'


def hash_data(data: bytes, hash_sensitive_data: bool = True) -> str:
    if not hash_sensitive_data:
        return ""
    return b64url_nopad(sha256(data))


def webhook_signature(secret: str, body: bytes) -> str:
    return "sha256=" + hmac_sha256(secret.encode(), body).hex()


def derive_seed(master_seed: bytes, path: str) -> bytes:
    return sha256(master_seed + path.encode())


def did_key(pk: bytes) -> str:
    # NOTE: the reference uses base64url (not base58btc) after the 'z' prefix (did_service.go:533-535)
    return "did:key:z" + b64url_nopad(b"\xed\x01" + pk)


def payload_sha256_hex(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()
