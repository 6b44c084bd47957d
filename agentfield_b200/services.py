"""Host-side mirror of the reference's service-level seams for this path, batched:

    VCService.hashData / signVC / verifyVCSignature / GenerateExecutionVC / VerifyVC
        control-plane/internal/services/vc_service.go:138-239, 242-289, 434-515
    generateWorkflowVCDocument / createWorkflowVCDocument / signWorkflowVC / determineWorkflowStatus / countCompletedSteps
        control-plane/internal/services/vc_service.go:525-787
    VerifyWorkflowVCComprehensive / performIntegrityChecks / performSecurityAnalysis / performComplianceChecks /
    verifyWorkflowVCSignature / calculateOverallScore
        control-plane/internal/services/vc_service.go:1037-1383, 1386-1644; pkg/types/status.go:28-75
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


_STATUS_CANONICAL = {"unknown", "pending", "queued", "running", "succeeded", "failed", "cancelled", "timeout"}
_STATUS_ALIASES = {"success": "succeeded", "successful": "succeeded", "completed": "succeeded", "complete": "succeeded", "done": "succeeded",
                   "ok": "succeeded", "error": "failed", "failure": "failed", "errored": "failed", "canceled": "cancelled", "cancel": "cancelled",
                   "timed_out": "timeout", "wait": "queued", "waiting": "queued", "in_progress": "running", "processing": "running"}


def normalize_execution_status(status: str) -> str:
    """types.NormalizeExecutionStatus (pkg/types/status.go:51-63)."""
    s = status.strip().lower()
    if not s:
        return "unknown"
    if s in _STATUS_CANONICAL:
        return s
    return _STATUS_ALIASES.get(s, "unknown")


def is_terminal_execution_status(status: str) -> bool:
    return normalize_execution_status(status) in ("succeeded", "failed", "cancelled", "timeout")


def determine_workflow_status(execution_vcs) -> str:
    """determineWorkflowStatus (vc_service.go:721-772): the worst state present wins."""
    if not execution_vcs:
        return "pending"
    seen = {normalize_execution_status(v["status"]) for v in execution_vcs}
    for s in ("failed", "timeout", "cancelled", "running", "queued", "pending", "unknown"):
        if s in seen:
            return s
    return "succeeded"


def count_completed_steps(execution_vcs) -> int:
    return sum(1 for v in execution_vcs if normalize_execution_status(v["status"]) == "succeeded")


def create_workflow_vc_document(workflow_id, session_id, component_vc_ids, status, start_time, end_time, issuer_did, vc_id, issuance_date,
                                snapshot_time):
    """createWorkflowVCDocument (vc_service.go:635-683).  Times are RFC 3339 strings (the reference formats time.Time with
    UTC().Format(time.RFC3339)); vc_id / issuance_date / snapshot_time are inputs here where the reference reads the clock."""
    n = len(component_vc_ids) if component_vc_ids is not None else 0
    cs = {"workflowId": workflow_id, "sessionId": session_id, "componentVcIds": component_vc_ids, "totalSteps": n, "completedSteps": n,
          "status": status, "startTime": start_time, "endTime": end_time, "snapshotTime": snapshot_time,
          "orchestrator": {"did": issuer_did, "type": "agentfield_server", "agentNodeDid": issuer_did},
          "audit": {"inputDataHash": "", "outputDataHash": "",
                    "metadata": {"agentfield_version": "1.0.0", "vc_version": "1.0", "workflow_type": "agent_execution_chain",
                                 "total_executions": n}}}
    return {"@context": ["https://www.w3.org/2018/credentials/v1", "https://agentfield.example.com/contexts/workflow/v1"],
            "type": ["VerifiableCredential", "AgentFieldWorkflowCredential"],
            "id": "urn:agentfield:workflow-vc:%s" % vc_id, "issuer": issuer_did, "issuanceDate": issuance_date, "credentialSubject": cs}


def generate_webhook_signature_batch(secrets, bodies, ctx=None):
    """generateWebhookSignature for many deliveries: "sha256=" + hex(HMAC-SHA256(secret, body))."""
    tags = MAC(ctx or default_context()).hmac_sha256_batch([s.encode() if isinstance(s, str) else s for s in secrets], bodies)
    return ["sha256=" + t.hex() for t in tags]


class VCService:
    """Batched issue / verify of execution VCs over a key cache (DID -> expanded key) and an issuer key set."""

    def __init__(self, keys: ExpandedKeys, ctx=None, hash_sensitive_data=True, canonical_on_device=False, store=None):
        self.ctx = ctx or keys.ctx
        self.keys = keys
        self.store = store                                        # vc_store.ExecutionVCStore: the issued batch is persisted in ONE transaction
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
            em = go_json.truncate_error_message(r.get("error_message"))                 # vc_service.go:153-160 (bytes, not runes)
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
        if self.store is not None and out:                        # vc_service.go:229-233, for the whole batch at once (group commit)
            self.store.store_execution_vcs([
                {"vc_id": r["vc_id"], "execution_id": r["execution_id"], "workflow_id": r["workflow_id"], "session_id": r["session_id"],
                 "issuer_did": r["caller_did"], "target_did": r.get("target_did", ""), "caller_did": r["caller_did"], "vc_document": v["vc_document"],
                 "signature": v["signature"], "input_hash": v["input_hash"], "output_hash": v["output_hash"],
                 "status": normalize_execution_status(r["status"])} for r, v in zip(requests, out)])
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

    # -- workflow-level credentials (vc_service.go:525-718)
    def generate_workflow_vc_batch(self, workflows):
        """generateWorkflowVCDocument for many workflows, all signatures in one GPU batch.  Each workflow: workflow_id,
        execution_vcs (the ExecutionVC records: vc_id, session_id, status, created_at (RFC 3339, UTC), issuer_did), root_did (the
        issuer when the workflow has no execution VCs), vc_id, workflow_vc_id, issuance_date, snapshot_time, proof_created.
        Returns WorkflowVC-shaped dicts (pkg/types/did_types.go:88-104) plus "doc" / "proof"."""
        docs, metas = [], []
        for w in workflows:
            evs = w["execution_vcs"]
            status = determine_workflow_status(evs)
            ids = [v["vc_id"] for v in evs]
            session_id = evs[0]["session_id"] if evs else ""
            if evs:
                times = [v["created_at"] for v in evs]                # RFC 3339 in UTC with equal precision sorts chronologically
                start, latest = min(times), max(times)
                end = latest if is_terminal_execution_status(status) else None
                issuer = evs[0]["issuer_did"]
            else:
                start, end, issuer = w["snapshot_time"], None, w["root_did"]
            doc = create_workflow_vc_document(w["workflow_id"], session_id, ids, status, start, end, issuer, w["vc_id"], w["issuance_date"],
                                              w["snapshot_time"])
            docs.append(doc)
            metas.append((w, evs, status, ids, session_id, start, end, issuer))
        sigs = self.keys.sign_batch([m[7] for m in metas], [go_json.workflow_vc_document(d) for d in docs])   # signWorkflowVC: zero proof
        out = []
        for doc, (w, evs, status, ids, session_id, start, end, issuer), sig in zip(docs, metas, sigs):
            proof = {"type": "Ed25519Signature2020", "created": w["proof_created"], "verificationMethod": "%s#key-1" % issuer,
                     "proofPurpose": "assertionMethod", "proofValue": _b64url(sig)}
            vc_bytes = go_json.workflow_vc_document(doc, proof)
            out.append({"workflow_id": w["workflow_id"], "session_id": session_id, "component_vcs": ids, "workflow_vc_id": w["workflow_vc_id"],
                        "status": status, "start_time": start, "end_time": end, "total_steps": len(evs),
                        "completed_steps": count_completed_steps(evs), "vc_document": vc_bytes, "signature": proof["proofValue"],
                        "issuer_did": issuer, "snapshot_time": w["snapshot_time"], "storage_uri": "", "document_size_bytes": len(vc_bytes),
                        "doc": doc, "proof": proof})
        return out

    def _keyed_verify(self, dids, msgs, proof_values):
        """ed25519.Verify for many (issuer DID, canonical bytes, base64url signature) through one keyed GPU batch.  A DID the
        cache cannot resolve or a signature that does not decode gives None (the reference reports those as errors, not False)."""
        known = [d in self.keys._index for d in dids]
        uniq = sorted({d for d, k in zip(dids, known) if k})
        if self._keyset is None or self._keyset_dids != uniq:
            if self._keyset is not None:
                self._keyset.close()
            self._keyset = KeySet([self.keys.public_key(d) for d in uniq], self.ctx) if uniq else None
            self._keyset_dids = uniq
        pos = {d: i for i, d in enumerate(uniq)}
        sg = np.zeros((len(dids), 64), dtype=np.uint8)
        state = []
        for i, (d, k, pv) in enumerate(zip(dids, known, proof_values)):
            try:
                raw = _b64url_decode_strict(pv)
            except ValueError:
                state.append("bad_encoding"); continue
            if not k:
                state.append("unresolved"); continue
            if len(raw) != 64:
                state.append("bad_length"); continue            # ed25519.Verify: plain false
            sg[i] = np.frombuffer(raw, dtype=np.uint8); state.append("ok")
        if not uniq:
            return [None if st in ("unresolved", "bad_encoding") else False for st in state]
        buf, off = pack(msgs)
        ki = np.array([pos.get(d, 0) for d in dids], dtype=np.uint32)
        ok = self._keyset.verify_packed(ki, sg, buf, off)
        return [bool(o) if st == "ok" else (False if st == "bad_length" else None) for o, st in zip(ok, state)]

    def verify_workflow_vc_batch(self, workflow_vcs):
        """verifyWorkflowVCSignature (vc_service.go:1589-1625) for many stored workflow VCs: parse, zero the proof, re-marshal
        (metadata numbers come back as float64), verify against the issuer's key."""
        import json
        docs = [json.loads(w["vc_document"], parse_int=float) for w in workflow_vcs]
        for d in docs:
            d["credentialSubject"]["audit"]["metadata"] = go_json.unmarshal_numbers(d["credentialSubject"]["audit"].get("metadata"))
        res = self._keyed_verify([d["issuer"] for d in docs], [go_json.workflow_vc_document(d) for d in docs],
                                 [d["proof"]["proofValue"] for d in docs])
        return [bool(r) for r in res]

    def verify_workflow_vc_comprehensive(self, chain):
        """VerifyWorkflowVCComprehensive (vc_service.go:1386-1586) over one workflow's VC chain: chain = {"component_vcs": [ExecutionVC
        records: vc_id, execution_id, workflow_id, session_id, issuer_did, target_did, caller_did, vc_document (bytes), signature,
        input_hash, output_hash, status], "workflow_vc": WorkflowVC record or None}.  Every signature of the chain — the component
        VCs' and the workflow VC's — goes through ONE keyed GPU batch (the reference verifies them one by one, :1442-1546).
        `verification_timestamp` is the caller's (the reference reads the clock)."""
        import json
        integrity = {"metadata_consistency": True, "field_consistency": True, "timestamp_validation": True, "hash_validation": True,
                     "structural_integrity": True, "issues": []}
        security = {"signature_strength": "Ed25519", "key_validation": True, "did_authenticity": True, "replay_protection": True,
                    "tamper_evidence": [], "security_score": 100.0, "issues": []}
        compliance = {"w3c_compliance": True, "agentfield_standard_compliance": True, "audit_trail_integrity": True,
                      "data_integrity_checks": True, "issues": []}
        critical = []

        def issue(kind, severity, component, description, field="", expected="", actual=""):
            return {"type": kind, "severity": severity, "component": component, "field": field, "expected": expected, "actual": actual,
                    "description": description}
        parsed = []
        for ev in chain["component_vcs"]:
            try:
                doc = json.loads(ev["vc_document"], parse_int=float)
                doc["credentialSubject"]["audit"]["metadata"] = go_json.unmarshal_numbers(doc["credentialSubject"]["audit"].get("metadata"))
                doc["credentialSubject"]["execution"].setdefault("errorMessage", "")
                parsed.append((ev, doc))
            except (ValueError, KeyError, TypeError) as ex:
                critical.append(issue("parse_error", "critical", ev["vc_id"], "Failed to parse VC document: %s" % ex))
        wf, wf_doc = chain.get("workflow_vc"), None
        if wf is not None and wf.get("vc_document") is not None:
            try:
                wf_doc = json.loads(wf["vc_document"], parse_int=float)
                wf_doc["credentialSubject"]["audit"]["metadata"] = go_json.unmarshal_numbers(wf_doc["credentialSubject"]["audit"].get("metadata"))
            except (ValueError, KeyError, TypeError) as ex:
                critical.append(issue("workflow_vc_parse_error", "critical", wf["workflow_vc_id"], "Failed to parse workflow VC document: %s" % ex))
        # one GPU batch for every signature of the chain
        dids = [d["issuer"] for _, d in parsed] + ([wf_doc["issuer"]] if wf_doc else [])
        msgs = [go_json.vc_document(d) for _, d in parsed] + ([go_json.workflow_vc_document(wf_doc)] if wf_doc else [])
        pvs = [d["proof"]["proofValue"] for _, d in parsed] + ([wf_doc["proof"]["proofValue"]] if wf_doc else [])
        verdicts = self._keyed_verify(dids, msgs, pvs) if dids else []
        for (ev, doc), verdict, did in zip(parsed, verdicts, dids):
            cs, ex = doc["credentialSubject"], doc["credentialSubject"]["execution"]
            ic_issues, tamper = [], []
            pairs = (("issuer_mismatch", "metadata_consistency", "issuer_did", ev["issuer_did"], doc["issuer"], "Issuer DID"),
                     ("execution_id_mismatch", "field_consistency", "execution_id", ev["execution_id"], cs["executionId"], "Execution ID"),
                     ("workflow_id_mismatch", "field_consistency", "workflow_id", ev["workflow_id"], cs["workflowId"], "Workflow ID"),
                     ("session_id_mismatch", "field_consistency", "session_id", ev["session_id"], cs["sessionId"], "Session ID"),
                     ("caller_did_mismatch", "field_consistency", "caller_did", ev["caller_did"], cs["caller"]["did"], "Caller DID"),
                     ("target_did_mismatch", "field_consistency", "target_did", ev["target_did"], cs["target"]["did"], "Target DID"))
            for kind, flag, field, want, got, label in pairs:
                if want != got:
                    integrity[flag] = False
                    ic_issues.append(issue(kind, "critical", ev["vc_id"], "%s mismatch between metadata and VC document" % label, field, want, got))
            if normalize_execution_status(ev["status"]) != normalize_execution_status(ex["status"]):
                integrity["field_consistency"] = False
                ic_issues.append(issue("status_mismatch", "critical", ev["vc_id"], "Status mismatch between metadata and VC document", "status", ev["status"], ex["status"]))
            for kind, field, want, got, label in (("input_hash_mismatch", "input_hash", ev["input_hash"], ex["inputHash"], "Input hash"),
                                                  ("output_hash_mismatch", "output_hash", ev["output_hash"], ex["outputHash"], "Output hash")):
                if want != got:
                    integrity["hash_validation"] = False
                    ic_issues.append(issue(kind, "critical", ev["vc_id"], "%s mismatch between metadata and VC document" % label, field, want, got))
            if ev["signature"] != doc["proof"]["proofValue"]:
                integrity["structural_integrity"] = False
                ic_issues.append(issue("signature_mismatch", "critical", ev["vc_id"], "Signature mismatch between metadata and VC document", "signature",
                                       ev["signature"], doc["proof"]["proofValue"]))
            if not _is_rfc3339(doc["issuanceDate"]):
                integrity["timestamp_validation"] = False
                ic_issues.append(issue("invalid_timestamp", "critical", ev["vc_id"], "Invalid timestamp", "issuance_date"))
            missing = next((name for name, v in (("@context", doc.get("@context")), ("type", doc.get("type")), ("id", doc.get("id")),
                                                 ("issuer", doc.get("issuer")), ("issuanceDate", doc.get("issuanceDate"))) if not v), None)
            if missing:
                integrity["structural_integrity"] = False
                ic_issues.append(issue("invalid_structure", "critical", ev["vc_id"], "Invalid VC structure: missing %s" % missing))
            integrity["issues"] += ic_issues
            score = 100.0
            if verdict is None and did not in self.keys._index:
                security["did_authenticity"] = False; score -= 50.0
                security["issues"].append(issue("did_resolution_failed", "critical", ev["vc_id"], "Failed to resolve issuer DID"))
            elif not verdict:
                security["key_validation"] = False; score -= 40.0
                security["issues"].append(issue("signature_verification_failed", "critical", ev["vc_id"], "Signature verification failed"))
            if ev["issuer_did"] != doc["issuer"]:
                tamper.append("issuer_did_inconsistency")
            if ev["execution_id"] != cs["executionId"]:
                tamper.append("execution_id_inconsistency")
            if ev["signature"] != doc["proof"]["proofValue"]:
                tamper.append("signature_inconsistency")
            if tamper:
                score -= 20.0
                security["tamper_evidence"] += tamper
                security["issues"].append(issue("tamper_evidence", "warning", ev["vc_id"], "Tamper evidence detected: %s" % tamper))
            security["security_score"] = min(security["security_score"], score)
            if "https://www.w3.org/2018/credentials/v1" not in (doc.get("@context") or []):
                compliance["w3c_compliance"] = False
                compliance["issues"].append(issue("w3c_compliance_failure", "warning", doc.get("id", ""), "VC does not meet W3C standards"))
            if not {"VerifiableCredential", "AgentFieldExecutionCredential"} <= set(doc.get("type") or []):
                compliance["agentfield_standard_compliance"] = False
                compliance["issues"].append(issue("agentfield_compliance_failure", "warning", doc.get("id", ""),
                                                  "VC does not meet AgentField standard requirements"))
        if wf_doc is not None:
            verdict = verdicts[-1]
            if verdict is None and wf_doc["issuer"] not in self.keys._index:
                security["did_authenticity"] = False
                security["issues"].append(issue("workflow_did_resolution_failed", "critical", wf["workflow_vc_id"], "Failed to resolve workflow VC issuer DID"))
            elif not verdict:
                security["key_validation"] = False
                security["issues"].append(issue("workflow_signature_verification_failed", "critical", wf["workflow_vc_id"],
                                                "Workflow VC signature verification failed"))
            if not {"VerifiableCredential", "AgentFieldWorkflowCredential"} <= set(wf_doc.get("type") or []):
                compliance["agentfield_standard_compliance"] = False
                compliance["issues"].append(issue("workflow_compliance_failure", "warning", wf["workflow_vc_id"],
                                                  "Workflow VC does not meet AgentField standard requirements"))
        all_issues = integrity["issues"] + security["issues"] + compliance["issues"]
        critical += [i for i in all_issues if i["severity"] == "critical"]
        warnings = [i for i in all_issues if i["severity"] == "warning"]
        score = max(0.0, ((100.0 - 25.0 * len(critical) - 5.0 * len(warnings)) + security["security_score"]) / 2.0)   # calculateOverallScore
        return {"valid": not critical, "overall_score": score, "critical_issues": critical, "warnings": warnings, "integrity_checks": integrity,
                "security_analysis": security, "compliance_checks": compliance,
                "verification_timestamp": chain.get("verification_timestamp", "")}


def _b64url_decode_strict(s: str) -> bytes:
    """base64.RawURLEncoding.DecodeString: URL alphabet, no padding, no stray characters (Go returns an error otherwise)."""
    import re
    if not re.fullmatch(r"[A-Za-z0-9_-]*", s) or len(s) % 4 == 1:
        raise ValueError("illegal base64 data")
    raw = base64.urlsafe_b64decode(s + "=" * (-len(s) % 4))
    if _b64url(raw) != s:                      # non-zero trailing bits: Go's strict decoder accepts them, Python ignores them too
        pass
    return raw


def _is_rfc3339(ts: str) -> bool:
    """time.Parse(time.RFC3339, ts) == nil (vc_service.go:1248-1251), for the forms the control plane emits and the obvious
    malformed ones: date, 'T', time, optional fraction, 'Z' or a numeric offset."""
    import re
    m = re.fullmatch(r"(\d{4})-(\d{2})-(\d{2})T(\d{2}):(\d{2}):(\d{2})(\.\d+)?(Z|[+-](\d{2}):(\d{2}))", ts)
    if not m:
        return False
    mo, d, h, mi, sec = (int(m.group(i)) for i in (2, 3, 4, 5, 6))
    if not (1 <= mo <= 12 and 1 <= d <= 31 and h <= 23 and mi <= 59 and sec <= 59):
        return False
    if m.group(9) is not None and not (int(m.group(9)) <= 23 and int(m.group(10)) <= 59):
        return False
    import calendar
    return d <= calendar.monthrange(int(m.group(1)), mo)[1]
