"""ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's issue -> verify flow for execution VCs, restated on the CPU
(BASELINE.json configs[0]: "issue+verify 1 000 synthetic 512 B agent-action VCs via the reference Go identity/audit path").

Follows control-plane/internal/services/vc_service.go:138-239 (GenerateExecutionVC), :374-431 (createVCDocument),
:434-505 (signVC / verifyVCSignature), :508-515 (hashData) and pkg/types/did_types.go:135-220 (field order).  Canonical
bytes come from Python's json module over insertion-ordered dicts plus Go's HTML/line-separator escapes — deliberately a
different construction from agentfield_b200/go_json.py so the two check each other.  Signatures: `cryptography` (OpenSSL).
"""
import base64
import json

from . import go_hash as H


def go_marshal(obj) -> bytes:
    s = json.dumps(obj, ensure_ascii=False, separators=(",", ":"))
    return (s.replace("<", "\\u003c").replace(">", "\\u003e").replace("&", "\\u0026").replace("\u2028", "\\u2028")
            .replace("\u2029", "\\u2029")).encode("utf-8")


def vc_document(r, input_hash, output_hash, proof=None):
    ex = {"inputHash": input_hash, "outputHash": output_hash, "timestamp": r["timestamp"], "durationMs": r["duration_ms"], "status": r["status"]}
    em = r.get("error_message")
    if em is not None and len(em) > 500:
        em = em[:500] + "...[truncated]"
    if em:
        ex["errorMessage"] = em
    return {
        "@context": ["https://www.w3.org/2018/credentials/v1", "https://agentfield.example.com/contexts/execution/v1"],
        "type": ["VerifiableCredential", "AgentFieldExecutionCredential"],
        "id": "urn:agentfield:vc:%s" % r["vc_id"], "issuer": r["caller_did"], "issuanceDate": r["issuance_date"],
        "credentialSubject": {
            "executionId": r["execution_id"], "workflowId": r["workflow_id"], "sessionId": r["session_id"],
            "caller": {"did": r["caller_did"], "type": r.get("caller_type", "agent"), "agentNodeDid": r["agent_node_did"]},
            "target": {"did": r.get("target_did", ""), "agentNodeDid": r["agent_node_did"], "functionName": r.get("function_name", "")},
            "execution": ex,
            "audit": {"inputDataHash": input_hash, "outputDataHash": output_hash,
                      "metadata": dict(sorted({"agentfield_version": "1.0.0", "vc_version": "1.0"}.items()))},
        },
        "proof": proof or {"type": "", "created": "", "verificationMethod": "", "proofPurpose": "", "proofValue": ""},
    }


def generate_execution_vc(r, seed: bytes):
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey
    ih = H.hash_data(H.marshal_data_or_null(r.get("input")))
    oh = H.hash_data(H.marshal_data_or_null(r.get("output")))
    canonical = go_marshal(vc_document(r, ih, oh))
    sig = Ed25519PrivateKey.from_private_bytes(seed).sign(canonical)
    proof = {"type": "Ed25519Signature2020", "created": r["proof_created"], "verificationMethod": "%s#key-1" % r["caller_did"],
             "proofPurpose": "assertionMethod", "proofValue": H.b64url_nopad(sig)}
    return {"vc_document": go_marshal(vc_document(r, ih, oh, proof)), "signature": proof["proofValue"], "canonical": canonical}


def verify_vc(vc_document_bytes: bytes, pk: bytes) -> bool:
    from . import go_ed25519 as G
    doc = json.loads(vc_document_bytes)
    sig = base64.urlsafe_b64decode(doc["proof"]["proofValue"] + "=" * (-len(doc["proof"]["proofValue"]) % 4))
    doc["proof"] = {"type": "", "created": "", "verificationMethod": "", "proofPurpose": "", "proofValue": ""}
    return G.verify(pk, go_marshal(doc), sig)
