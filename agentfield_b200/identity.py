"""Host-side mirror of the Go `internal/identity` package this build adds: the key cache.

Two caches, both replacing work the reference redoes on every request:

  * ExpandedKeys — DID -> (clamped scalar s, prefix, pk) so that issuing a VC costs one fixed-base multiplication
    instead of the three the reference performs (NewKeyFromSeed inside every ResolveDID, did_service.go:515-525,585-599;
    NewKeyFromSeed again in signVC, vc_service.go:460; then Sign, :463).
  * KeySet — a set of issuer public keys with per-key radix-256 tables on the GPU, so that verification against a
    known issuer needs no doublings (VerifyVC resolves the issuer from the registry first, vc_service.go:259;
    VerifyWorkflowVCComprehensive loops over one workflow's VCs, :1442-1512).

Derivation follows the reference exactly: seed' = SHA-256(masterSeed || path), did:key = "did:key:z" + base64url(0xED 0x01 || pk)
(did_service.go:515-536 — base64url, not base58btc, despite the 'z').
"""
import base64
import ctypes as C

import numpy as np

from . import _abi
from .crypto import default_context, pack


def did_key(pk: bytes) -> str:
    return "did:key:z" + base64.urlsafe_b64encode(b"\xed\x01" + pk).rstrip(b"=").decode()


class ExpandedKeys:
    """DID -> expanded signing key, filled in bulk on the GPU."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._index = {}
        self._expanded = np.zeros((0, 96), dtype=np.uint8)

    def derive(self, master_seed: bytes, paths):
        """derivePrivateKey for many paths at once: returns the DIDs, in order, and caches their expanded keys."""
        buf, off = pack([master_seed + p.encode() for p in paths])
        seeds = self.ctx.sha256_packed(buf, off)                 # SHA-256(masterSeed || path)
        exp = self.ctx.expand(seeds)
        dids = [did_key(bytes(e[64:])) for e in exp]
        base = self._expanded.shape[0]
        self._expanded = np.concatenate([self._expanded, exp])
        for i, d in enumerate(dids):
            self._index[d] = base + i
        return dids

    def public_key(self, did):
        return bytes(self._expanded[self._index[did], 64:])

    def sign_batch(self, dids, msgs):
        """signVC for a batch: one fixed-base multiplication per credential."""
        ki = np.fromiter((self._index[d] for d in dids), dtype=np.uint32, count=len(dids))
        buf, off = pack(msgs)
        return [bytes(s) for s in self.ctx.sign_expanded_packed(self._expanded, ki, buf, off)]

    def sign_dev(self, dids, d_msgs, d_off, n):
        """The same for messages already on the device (e.g. assembled there by canonical.JsonTemplate.fill_dev)."""
        import torch
        dev = d_off.device
        ki = np.fromiter((self._index[d] for d in dids), dtype=np.uint32, count=len(dids))
        d_exp = torch.from_numpy(self._expanded).to(dev)
        d_ki = torch.from_numpy(ki.view(np.int32)).to(dev)
        d_sigs = torch.empty((n, 64), dtype=torch.uint8, device=dev)
        self.ctx.sign_expanded_dev(d_exp, d_ki, d_msgs, d_off, n, d_sigs)
        torch.cuda.synchronize(dev)
        return [bytes(s) for s in d_sigs.cpu().numpy()]


class KeySet:
    """afc_keyset: issuer public keys with device-resident verification tables (384 KB per key)."""

    def __init__(self, pks, ctx=None):
        self.ctx = ctx or default_context()
        self._lib = _abi.load()
        for p in pks:
            if len(p) != 32:
                raise ValueError("ed25519: bad public key length: %d" % len(p))        # Go panics
        self.pks = [bytes(p) for p in pks]
        arr = np.frombuffer(b"".join(self.pks), dtype=np.uint8).reshape(-1, 32).copy()
        h = C.c_void_p()
        _abi.check(self._lib.afc_keyset_new(self.ctx.handle, _abi.ptr(arr), arr.shape[0], C.byref(h)), self.ctx.handle)
        self.handle = h
        self._index = {p: i for i, p in enumerate(self.pks)}

    def close(self):
        if getattr(self, "handle", None):
            self._lib.afc_keyset_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        n, b = C.c_uint32(), C.c_uint64()
        _abi.check(self._lib.afc_keyset_info(self.handle, C.byref(n), C.byref(b)))
        return {"n_keys": n.value, "table_bytes": b.value}

    def index_of(self, pk):
        return self._index[bytes(pk)]

    def verify_packed(self, key_index, sigs, buf, off):
        n = len(off) - 1
        ki = np.ascontiguousarray(key_index, dtype=np.uint32)
        out = np.empty(n, dtype=np.uint8)
        _abi.check(self._lib.afc_ed25519_verify_keyed_batch(self.ctx.handle, self.handle, _abi.ptr(ki), _abi.ptr(sigs), _abi.ptr(buf), _abi.ptr(off),
                                                            n, _abi.ptr(out)), self.ctx.handle)
        return out

    def verify_dev(self, d_key_index, d_sigs, d_msgs, d_off, n, d_ok, stream=None):
        _abi.check(self._lib.afc_ed25519_verify_keyed_batch_dev(self.ctx.handle, self.handle, _abi.ptr(d_key_index), _abi.ptr(d_sigs), _abi.ptr(d_msgs),
                                                                _abi.ptr(d_off), n, _abi.ptr(d_ok), self.ctx._stream(stream)), self.ctx.handle)

    def verify_batch(self, pks, msgs, sigs):
        """VerifyBatch against registered issuers (same contract as Verifier.verify_batch)."""
        n = len(pks)
        if not (n == len(msgs) == len(sigs)):
            raise ValueError("pks, msgs and sigs must have the same length")
        if n == 0:
            return []
        ki = np.fromiter((self._index[bytes(p)] for p in pks), dtype=np.uint32, count=n)
        bad_len = [len(s) != 64 for s in sigs]
        sg = np.zeros((n, 64), dtype=np.uint8)
        for i, s in enumerate(sigs):
            if not bad_len[i]:
                sg[i] = np.frombuffer(s, dtype=np.uint8)
        buf, off = pack(msgs)
        ok = self.verify_packed(ki, sg, buf, off)
        return [bool(o) and not b for o, b in zip(ok, bad_len)]
