"""Host-side mirror of the batching seam (BASELINE.json configs[4]): the native ingest dispatcher of
libafcrypto (csrc/afc_ingest.cu).  It plays the role of the reference's webhook worker pool (`WebhookDispatcher.Start/
Notify`, control-plane/internal/services/webhook_dispatcher.go:109-210) and of the one-VC-per-request issuing path
(internal/handlers/did_handlers.go:192) — with GPU batches formed by `batch_max` / `linger_us`.
"""
import ctypes as C

import numpy as np

from . import _abi
from .crypto import default_context


class Ingest:
    def __init__(self, expanded96, ctx=None, batch_max=8192, linger_us=200, max_msg=2048, max_key=64, max_body=1024):
        self.ctx = ctx or default_context()
        self._lib = _abi.load()
        exp = np.ascontiguousarray(expanded96, dtype=np.uint8).reshape(-1, 96)
        h = C.c_void_p()
        _abi.check(self._lib.afc_ingest_new(self.ctx.handle, _abi.ptr(exp), exp.shape[0], batch_max, linger_us, max_msg, max_key, max_body,
                                            C.byref(h)), self.ctx.handle)
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self._lib.afc_ingest_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, key_index, msg: bytes, hmac_key: bytes, body: bytes) -> int:
        t = C.c_uint64()
        _abi.check(self._lib.afc_ingest_submit(self.handle, key_index, _abi.ptr(msg) if msg else None, len(msg),
                                               _abi.ptr(hmac_key) if hmac_key else None, len(hmac_key), _abi.ptr(body) if body else None, len(body),
                                               C.byref(t)))
        return t.value

    def wait(self, ticket):
        sig, tag = (C.c_uint8 * 64)(), (C.c_uint8 * 32)()
        _abi.check(self._lib.afc_ingest_wait(self.handle, ticket, C.cast(sig, C.c_void_p), C.cast(tag, C.c_void_p)), self.ctx.handle)
        return bytes(sig), bytes(tag)

    def flush(self):
        _abi.check(self._lib.afc_ingest_flush(self.handle), self.ctx.handle)

    @staticmethod
    def _stats(s):
        return {"submitted": s.submitted, "completed": s.completed, "batches": s.batches, "avg_batch": s.avg_batch, "p50_us": s.p50_us,
                "p99_us": s.p99_us, "max_us": s.max_us, "log_size": s.log_size, "log_root": bytes(s.log_root).hex(), "last_error": s.last_error}

    def stats(self):
        s = _abi.IngestStats()
        _abi.check(self._lib.afc_ingest_stats_get(self.handle, C.byref(s)), self.ctx.handle)
        return self._stats(s)

    def soak(self, rate_per_s, seconds, producers=4, msg_len=512, body_len=256, seed=0xAF05):
        s, rate, late = _abi.IngestStats(), C.c_double(), C.c_uint64()
        _abi.check(self._lib.afc_ingest_soak(self.handle, float(rate_per_s), float(seconds), producers, msg_len, body_len, seed, C.byref(s),
                                             C.byref(rate), C.byref(late)), self.ctx.handle)
        d = self._stats(s)
        d.update({"target_rate": rate_per_s, "achieved_rate": rate.value, "late_submits": late.value, "seconds": seconds, "producers": producers})
        return d
