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


_INVALID_BYTE = "\ue0ff"      # stands for ONE byte of invalid UTF-8 inside a Go string (json.Marshal writes the six characters \\ufffd for it)


def _go_bytes_as_str(raw: bytes) -> str:
    """A Go string that may hold invalid UTF-8 (a message cut inside a rune) as a Python str: valid runes as themselves, every
    byte DecodeRune rejects as the private-use marker above."""
    from . import go_json as OJ
    out, i = [], 0
    while i < len(raw):
        if raw[i] < 0x80:
            out.append(chr(raw[i])); i += 1
            continue
        size, ok = OJ._decode_rune(raw, i)
        if ok:
            out.append(raw[i:i + size].decode("utf-8")); i += size
        else:
            out.append(_INVALID_BYTE); i += 1
    return "".join(out)


def _go_string(x: str) -> str:
    s = json.dumps(x, ensure_ascii=False).replace(_INVALID_BYTE, "\\ufffd")
    s = s.replace("\x7f", "\x7f")                       # DEL is not escaped by Go either
    return (s.replace("<", "\\u003c").replace(">", "\\u003e").replace("&", "\\u0026").replace("\u2028", "\\u2028")
            .replace("\u2029", "\\u2029"))


def go_marshal(obj) -> bytes:
    """json.Marshal of a value tree: dicts in insertion order (struct fields in declaration order; build maps with sorted keys),
    strings through Python's json module plus Go's HTML / line-separator escapes, float64 through oracle.go_json.number_bytes,
    ints as Go ints."""
    from . import go_json as OJ

    def enc(v):
        if v is None:
            return "null"
        if v is True:
            return "true"
        if v is False:
            return "false"
        if isinstance(v, int):
            return str(v)
        if isinstance(v, float):
            return OJ.number_bytes(v).decode()
        if isinstance(v, str):
            return _go_string(v)
        if isinstance(v, (list, tuple)):
            return "[" + ",".join(enc(x) for x in v) + "]"
        if isinstance(v, dict):
            return "{" + ",".join(_go_string(k) + ":" + enc(x) for k, x in v.items()) + "}"
        raise TypeError(type(v))
    return enc(obj).encode("utf-8")


def sorted_map(m):
    """A Go map as json.Marshal writes it: keys sorted bytewise, nested maps too; numbers that went through json.Unmarshal into
    interface{} are float64 (floats=True)."""
    if isinstance(m, dict):
        return {k: sorted_map(m[k]) for k in sorted(m, key=lambda k: k.encode("utf-8", "surrogatepass"))}
    if isinstance(m, (list, tuple)):
        return [sorted_map(x) for x in m]
    return m


def unmarshal_interface(text: str):
    """json.Unmarshal(text, &interface{}): every number a float64."""
    return json.loads(text, parse_int=float)


def vc_document_from_fields(f, proof=None):
    """types.VCDocument (pkg/types/did_types.go:135-144, 157-220) from its field values, as an insertion-ordered dict: what
    json.Marshal sees.  f: context, type (lists | None), id, issuer, issuance_date, execution_id, workflow_id, session_id,
    caller {did, type, agent_node_did}, target {did, agent_node_did, function_name}, input_hash, output_hash, timestamp,
    duration_ms, status, error_message (str, already truncated; "" omitted), input_data_hash, output_data_hash, metadata."""
    ex = {"inputHash": f["input_hash"], "outputHash": f["output_hash"], "timestamp": f["timestamp"], "durationMs": f["duration_ms"],
          "status": f["status"]}
    if f.get("error_message"):
        ex["errorMessage"] = f["error_message"]
    return {"@context": f["context"], "type": f["type"], "id": f["id"], "issuer": f["issuer"], "issuanceDate": f["issuance_date"],
            "credentialSubject": {
                "executionId": f["execution_id"], "workflowId": f["workflow_id"], "sessionId": f["session_id"],
                "caller": {"did": f["caller"]["did"], "type": f["caller"]["type"], "agentNodeDid": f["caller"]["agent_node_did"]},
                "target": {"did": f["target"]["did"], "agentNodeDid": f["target"]["agent_node_did"], "functionName": f["target"]["function_name"]},
                "execution": ex,
                "audit": {"inputDataHash": f["input_data_hash"], "outputDataHash": f["output_data_hash"], "metadata": sorted_map(f["metadata"])}},
            "proof": proof or {"type": "", "created": "", "verificationMethod": "", "proofPurpose": "", "proofValue": ""}}


def truncate_error_message(msg: str) -> str:
    """vc_service.go:153-160 on a Go string: byte length, byte slice; bytes of a rune cut in half become the invalid-byte marker."""
    raw = msg.encode("utf-8")
    return msg if len(raw) <= 500 else _go_bytes_as_str(raw[:500]) + "...[truncated]"


def webhook_payload(p):
    """types.ExecutionWebhookPayload (pkg/types/webhook.go:42-53) as an insertion-ordered dict (omitempty members dropped)."""
    d = {"event": p["event"], "execution_id": p["execution_id"], "workflow_id": p["workflow_id"], "status": p["status"], "target": p["target"],
         "type": p["type"]}
    if p.get("duration_ms") is not None:
        d["duration_ms"] = p["duration_ms"]
    if p.get("result") is not None:
        d["result"] = sorted_map(p["result"])
    if p.get("error_message") is not None:
        d["error_message"] = p["error_message"]
    d["timestamp"] = p["timestamp"]
    return d


def vc_document(r, input_hash, output_hash, proof=None):
    ex = {"inputHash": input_hash, "outputHash": output_hash, "timestamp": r["timestamp"], "durationMs": r["duration_ms"], "status": r["status"]}
    em = r.get("error_message")
    if em is not None and len(em.encode("utf-8")) > 500:                      # Go: len(msg) and msg[:500] count bytes (:153-160)
        em = _go_bytes_as_str(em.encode("utf-8")[:500]) + "...[truncated]"
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


# ---- workflow-level credential (vc_service.go:525-718, 1589-1625; pkg/types/did_types.go:146-180)
def workflow_vc_document(w, proof=None):
    """createWorkflowVCDocument (:635-683) as an insertion-ordered dict in struct order.  w: workflow_id, session_id,
    component_vc_ids (list | None), status, start_time, end_time (str | None), snapshot_time, issuer_did, vc_id, issuance_date."""
    ids = w["component_vc_ids"]
    n = len(ids) if ids is not None else 0
    cs = {"workflowId": w["workflow_id"], "sessionId": w["session_id"], "componentVcIds": ids, "totalSteps": n, "completedSteps": n,
          "status": w["status"], "startTime": w["start_time"]}
    if w.get("end_time") is not None:                                          # *string, omitempty
        cs["endTime"] = w["end_time"]
    cs["snapshotTime"] = w["snapshot_time"]
    cs["orchestrator"] = {"did": w["issuer_did"], "type": "agentfield_server", "agentNodeDid": w["issuer_did"]}
    cs["audit"] = {"inputDataHash": "", "outputDataHash": "",
                   "metadata": dict(sorted({"agentfield_version": "1.0.0", "vc_version": "1.0", "workflow_type": "agent_execution_chain",
                                            "total_executions": n}.items()))}
    return {"@context": ["https://www.w3.org/2018/credentials/v1", "https://agentfield.example.com/contexts/workflow/v1"],
            "type": ["VerifiableCredential", "AgentFieldWorkflowCredential"],
            "id": "urn:agentfield:workflow-vc:%s" % w["vc_id"], "issuer": w["issuer_did"], "issuanceDate": w["issuance_date"],
            "credentialSubject": cs,
            "proof": proof or {"type": "", "created": "", "verificationMethod": "", "proofPurpose": "", "proofValue": ""}}


def determine_workflow_status(statuses):
    """determineWorkflowStatus (:721-772) over already-normalised execution statuses."""
    if not statuses:
        return "pending"
    for s in ("failed", "timeout", "cancelled", "running", "queued", "pending", "unknown"):
        if s in statuses:
            return s
    return "succeeded"


def generate_workflow_vc(w, seed: bytes):
    """generateWorkflowVCDocument (:525-632): sign json.Marshal(doc with zero proof), attach the proof, marshal for storage."""
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey
    canonical = go_marshal(workflow_vc_document(w))
    sig = Ed25519PrivateKey.from_private_bytes(seed).sign(canonical)
    proof = {"type": "Ed25519Signature2020", "created": w["proof_created"], "verificationMethod": "%s#key-1" % w["issuer_did"],
             "proofPurpose": "assertionMethod", "proofValue": H.b64url_nopad(sig)}
    return {"vc_document": go_marshal(workflow_vc_document(w, proof)), "signature": proof["proofValue"], "canonical": canonical}


def verify_workflow_vc(vc_document_bytes: bytes, pk: bytes) -> bool:
    """verifyWorkflowVCSignature (:1589-1625) on the stored bytes: parse (numbers in metadata become float64), zero the proof,
    re-marshal, verify under Go's rules."""
    from . import go_ed25519 as G
    from . import go_json as OJ
    doc = json.loads(vc_document_bytes)
    sig = base64.urlsafe_b64decode(doc["proof"]["proofValue"] + "=" * (-len(doc["proof"]["proofValue"]) % 4))
    cs = doc["credentialSubject"]
    au = cs["audit"]
    # struct fields re-marshal in declaration order; the metadata map goes through float64 and sorted keys
    parts = [b'{"@context":' + OJ.value_bytes(doc["@context"]), b'"type":' + OJ.value_bytes(doc["type"]), b'"id":' + OJ.value_bytes(doc["id"]),
             b'"issuer":' + OJ.value_bytes(doc["issuer"]), b'"issuanceDate":' + OJ.value_bytes(doc["issuanceDate"])]
    sub = [b'"workflowId":' + OJ.value_bytes(cs["workflowId"]), b'"sessionId":' + OJ.value_bytes(cs["sessionId"]),
           b'"componentVcIds":' + OJ.value_bytes(cs["componentVcIds"]), b'"totalSteps":%d' % cs["totalSteps"],
           b'"completedSteps":%d' % cs["completedSteps"], b'"status":' + OJ.value_bytes(cs["status"]), b'"startTime":' + OJ.value_bytes(cs["startTime"])]
    if cs.get("endTime") is not None:
        sub.append(b'"endTime":' + OJ.value_bytes(cs["endTime"]))
    o = cs["orchestrator"]
    sub += [b'"snapshotTime":' + OJ.value_bytes(cs["snapshotTime"]),
            b'"orchestrator":{"did":' + OJ.value_bytes(o["did"]) + b',"type":' + OJ.value_bytes(o["type"]) + b',"agentNodeDid":' + OJ.value_bytes(o["agentNodeDid"]) + b"}",
            b'"audit":{"inputDataHash":' + OJ.value_bytes(au["inputDataHash"]) + b',"outputDataHash":' + OJ.value_bytes(au["outputDataHash"]) +
            b',"metadata":' + OJ.value_bytes(au["metadata"]) + b"}"]
    parts.append(b'"credentialSubject":{' + b",".join(sub) + b"}")
    parts.append(b'"proof":{"type":"","created":"","verificationMethod":"","proofPurpose":"","proofValue":""}}')
    return G.verify(pk, b",".join(parts), sig)
