"""Host-side mirror of the reference's service-level seams for this path, batched:

    VCService.hashData / signVC / verifyVCSignature / GenerateExecutionVC / VerifyVC
        control-plane/internal/services/vc_service.go:138-239, 242-289, 434-515
    generateWebhookSignature
        control-plane/internal/services/webhook_dispatcher.go:470-474

Same names (snake_cased) and behaviour; each takes a BATCH (the natural seam is the loop in
VerifyWorkflowVCComprehensive, vc_service.go:1442-1512, or a linger queue in front of one-per-request calls) and does
every hash / signature on the GPU through the C ABI.  The canonical JSON bytes come from go_json.py on the host or, with
canonical_on_device=True, from the device template (canonical.py): values -> documents -> signatures without a host round trip.
"""
import base64

import numpy as np

from . import go_json
from .crypto import Hasher, MAC, default_context, pack
from .identity import ExpandedKeys, KeySet


def _b64url(b: bytes) -> str:
    return base64.urlsafe_b64encode(b).rstrip(b"=").decode()


def _b64url_decode(s: str) -> bytes:
    return base64.urlsafe_b64decode(s + "=" * (-len(s) % 4))


def marshal_data_or_null(data):
    """vc_service.go:1298-1306 — json.Marshal([]byte): nil -> null, else the std-base64 text in quotes."""
    if data is None:
        return b"null"
    return b'"' + base64.b64encode(data) + b'"'


def generate_webhook_signature_batch(secrets, bodies, ctx=None):
    """generateWebhookSignature for many deliveries: "sha256=" + hex(HMAC-SHA256(secret, body))."""
    tags = MAC(ctx or default_context()).hmac_sha256_batch([s.encode() if isinstance(s, str) else s for s in secrets], bodies)
    return ["sha256=" + t.hex() for t in tags]


class VCService:
    """Batched issue / verify of execution VCs over a key cache (DID -> expanded key) and an issuer key set."""

    def __init__(self, keys: ExpandedKeys, ctx=None, hash_sensitive_data=True, canonical_on_device=False):
        self.ctx = ctx or keys.ctx
        self.keys = keys
        self.hash_sensitive_data = hash_sensitive_data
        self.canonical_on_device = canonical_on_device
        self._templates = None
        self._keyset = None
        self._keyset_dids = None

    # -- H1
    def hash_data_batch(self, datas):
        if not self.hash_sensitive_data:
            return [""] * len(datas)
        return [_b64url(d) for d in Hasher(self.ctx).sha256_batch(datas)]

    # -- E1 / D1
    def generate_execution_vc_batch(self, requests):
        """requests: dicts with execution_id, workflow_id, session_id, caller_did, target_did, agent_node_did, caller_type,
        function_name, input (bytes|None), output (bytes|None), status, error_message (str|None), duration_ms, timestamp,
        vc_id, issuance_date, proof_created.  Returns [{"vc_document": bytes, "signature": str, "input_hash", "output_hash"}]."""
        n = len(requests)
        hashes = self.hash_data_batch([marshal_data_or_null(r.get("input")) for r in requests] +
                                      [marshal_data_or_null(r.get("output")) for r in requests])
        docs, msgs = [], []
        for i, r in enumerate(requests):
            ih, oh = hashes[i], hashes[n + i]
            em = r.get("error_message")
            if em is not None and len(em) > 500:
                em = em[:500] + "...[truncated]"                                         # vc_service.go:153-160
            doc = {
                "@context": ["https://www.w3.org/2018/credentials/v1", "https://agentfield.example.com/contexts/execution/v1"],
                "type": ["VerifiableCredential", "AgentFieldExecutionCredential"],
                "id": "urn:agentfield:vc:%s" % r["vc_id"], "issuer": r["caller_did"], "issuanceDate": r["issuance_date"],
                "credentialSubject": {
                    "executionId": r["execution_id"], "workflowId": r["workflow_id"], "sessionId": r["session_id"],
                    "caller": {"did": r["caller_did"], "type": r.get("caller_type", "agent"), "agentNodeDid": r["agent_node_did"]},
                    "target": {"did": r.get("target_did", ""), "agentNodeDid": r["agent_node_did"], "functionName": r.get("function_name", "")},
                    "execution": {"inputHash": ih, "outputHash": oh, "timestamp": r["timestamp"], "durationMs": r["duration_ms"],
                                  "status": r["status"], "errorMessage": em or ""},
                    "audit": {"inputDataHash": ih, "outputDataHash": oh, "metadata": {"agentfield_version": "1.0.0", "vc_version": "1.0"}},
                },
            }
            docs.append(doc)
            if not self.canonical_on_device:
                msgs.append(go_json.vc_document(doc))                                    # zero-valued proof: what signVC signs
        dids = [r["caller_did"] for r in requests]
        if self.canonical_on_device and n:
            import torch
            from . import canonical
            if self._templates is None:
                self._templates = (canonical.vc_document_template(False, self.ctx), canonical.vc_document_template(True, self.ctx))
            t0, t1 = self._templates
            dev = torch.device("cuda", self.ctx.device)
            fields, off = t0.pack_values([canonical.vc_document_values(d) for d in docs])
            d_msgs, d_off = t0.fill_dev(torch.from_numpy(fields).to(dev), torch.from_numpy(off.view(np.int64)).to(dev), n)
            sigs = self.keys.sign_dev(dids, d_msgs, d_off, n)                            # documents never leave the device
        else:
            sigs = self.keys.sign_batch(dids, msgs)                                      # one GPU batch
        proofs = [{"type": "Ed25519Signature2020", "created": r["proof_created"], "verificationMethod": "%s#key-1" % r["caller_did"],
                   "proofPurpose": "assertionMethod", "proofValue": _b64url(sig)} for r, sig in zip(requests, sigs)]
        if self.canonical_on_device and n:
            stored = self._templates[1].fill([canonical.vc_document_values(d, p) for d, p in zip(docs, proofs)])
        else:
            stored = [go_json.vc_document(d, p) for d, p in zip(docs, proofs)]
        out = []
        for doc, proof, vc_bytes in zip(docs, proofs, stored):
            out.append({"vc_document": vc_bytes, "signature": proof["proofValue"],
                        "input_hash": doc["credentialSubject"]["execution"]["inputHash"],
                        "output_hash": doc["credentialSubject"]["execution"]["outputHash"], "doc": doc, "proof": proof})
        return out

    # -- E2 / D1
    def verify_vc_batch(self, vcs):
        """VerifyVC for many stored documents: re-marshal the parsed document with a zero proof, resolve the issuer key from
        the cache, one keyed GPU verification batch.  vcs: the dicts returned by generate_execution_vc_batch (doc + proof)."""
        dids = [v["doc"]["issuer"] for v in vcs]
        uniq = sorted(set(dids))
        if self._keyset is None or self._keyset_dids != uniq:
            if self._keyset is not None:
                self._keyset.close()
            self._keyset = KeySet([self.keys.public_key(d) for d in uniq], self.ctx)
            self._keyset_dids = uniq
        pos = {d: i for i, d in enumerate(uniq)}
        msgs = [go_json.vc_document(v["doc"]) for v in vcs]
        sigs_raw = [_b64url_decode(v["proof"]["proofValue"]) for v in vcs]
        bad = [len(s) != 64 for s in sigs_raw]
        sg = np.zeros((len(vcs), 64), dtype=np.uint8)
        for i, s in enumerate(sigs_raw):
            if not bad[i]:
                sg[i] = np.frombuffer(s, dtype=np.uint8)
        buf, off = pack(msgs)
        ok = self._keyset.verify_packed(np.array([pos[d] for d in dids], dtype=np.uint32), sg, buf, off)
        return [bool(o) and not b for o, b in zip(ok, bad)]
