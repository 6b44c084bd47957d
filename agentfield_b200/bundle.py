"""Bulk offline audit: the export bundle an auditor receives and `verify_bundle`, the GPU twin of `af vc verify`.

Reference being mirrored (control-plane/):
    internal/cli/vc.go:159-331            verifyVC: read file, parse EnhancedVCChain, collect unique DIDs, resolve them from the
                                          bundled did_resolution_bundle (:362-387), verify every component, summarise
    internal/cli/vc.go:78-97              EnhancedVCChain / VerificationMetadata (the JSON shape kept here field for field)
    internal/cli/vc_verification_enhanced.go:94-137, 418-454   per-VC consistency checks + ed25519.Verify over the re-marshalled document
    internal/services/vc_service.go:810-870   collectDIDResolutionBundle (method, public_key_jwk, resolved_from, resolved_at)
    internal/cli/vc_verification_enhanced.go:531-534   checkChainIntegrity — a stub that returns true: the part this build makes real

What is new (the reference has no tamper-evident log): an `audit_log` section —

    "audit_log": {
      "format": "rfc6962-sha256/1",
      "leaf": "json.Marshal(VCDocument with proof)",          # leaf i of the log = the stored vc_document bytes of one credential
      "signed_tree_head": {"tree_size": N, "root_hash": b64url, "timestamp": RFC 3339, "issuer": did, "signature": b64url},
      "entries": [{"vc_id": ..., "leaf_index": i, "inclusion_proof": [b64url, ...]}, ...],          # RFC 6962 §2.1.1 audit paths
      "checkpoints": [{"tree_size": m, "root_hash": b64url, "consistency_proof": [b64url, ...]}, ...]   # RFC 6962 §2.1.2, m <= N
    }

The signed tree head is signed over its canonical bytes `{"tree_size":N,"root_hash":"...","timestamp":"..."}` (Go json.Marshal
order of a three-field struct) by the server's root DID key, with the same Ed25519 call as every other signature on this path.

`verify_bundle` does on the GPU, in three calls, what the reference's CLI does credential by credential on one core:
  1. ONE keyed Ed25519 batch: every execution VC + the workflow VC + the signed tree head     (afc_ed25519_verify_keyed_batch)
  2. leaf hashes of the stored documents + ONE inclusion batch against the tree head          (afc_sha256_batch, afc_merkle_verify_inclusion_batch)
  3. ONE consistency batch for the earlier checkpoints the auditor already trusts             (afc_merkle_verify_consistency_batch)
and reports in the shape of the reference's VCVerificationResult (vc.go:99-145) plus an `audit_log` block.
"""
import base64
import json

import numpy as np

from . import go_json
from .audit import MerkleTree, verify_consistency_batch, verify_inclusion_batch
from .crypto import Hasher, default_context, pack
from .identity import KeySet
from .services import _b64url_decode_strict, normalize_execution_status

FORMAT = "rfc6962-sha256/1"


def _b64(b: bytes) -> str:
    return base64.urlsafe_b64encode(b).rstrip(b"=").decode()


def public_key_jwk(pk: bytes) -> dict:
    """ed25519PublicKeyToJWK (did_service.go:560-575); json.Marshal of the map sorts the keys: alg, crv, kty, use, x."""
    return {"alg": "EdDSA", "crv": "Ed25519", "kty": "OKP", "use": "sig", "x": _b64(pk)}


def tree_head_bytes(tree_size: int, root: bytes, timestamp: str) -> bytes:
    return ('{"tree_size":%d,"root_hash":%s,"timestamp":%s}' % (tree_size, go_json.string(_b64(root)), go_json.string(timestamp))).encode()


def stored_leaf(execution_vc) -> bytes:
    """The log leaf of one credential: its stored vc_document bytes (json.Marshal(VCDocument) with the proof)."""
    d = execution_vc["vc_document"]
    return d if isinstance(d, (bytes, bytearray)) else _remarshal_stored(d)


def _remarshal_stored(doc: dict) -> bytes:
    doc = json.loads(json.dumps(doc), parse_int=float)
    doc["credentialSubject"]["audit"]["metadata"] = go_json.unmarshal_numbers(doc["credentialSubject"]["audit"].get("metadata"))
    doc["credentialSubject"]["execution"].setdefault("errorMessage", "")
    return go_json.vc_document(doc, doc.get("proof") or go_json.EMPTY_PROOF)


def export_bundle(workflow_id, execution_vcs, workflow_vc, public_keys, log_leaves, leaf_index_of, sign_tree_head, head_issuer, generated_at,
                  checkpoints=(), ctx=None):
    """Builds the export an auditor verifies offline.
      execution_vcs   ExecutionVC records (vc_id, execution_id, workflow_id, session_id, issuer_did, target_did, caller_did,
                      vc_document (bytes), signature, input_hash, output_hash, status, created_at)
      workflow_vc     WorkflowVC record (services.generate_workflow_vc_batch) or None
      public_keys     DID -> 32-byte public key for every issuer (the did_resolution_bundle)
      log_leaves      every leaf of the audit log so far (bytes each), in log order; leaf_index_of: vc_id -> index into it
      sign_tree_head  callable(bytes) -> 64-byte signature by `head_issuer` (e.g. ExpandedKeys.sign_batch)
      checkpoints     earlier tree sizes the auditor may already hold a root for: a consistency proof is attached for each
    The tree levels are materialised on the GPU (afc_merkle_tree_build) and all audit paths read out in one call."""
    ctx = ctx or default_context()
    leaf_hashes = Hasher(ctx).sha256_batch([b"\x00" + bytes(l) for l in log_leaves])
    tree = MerkleTree(np.frombuffer(b"".join(leaf_hashes), dtype=np.uint8).reshape(-1, 32), ctx)
    root = tree.root
    n = len(log_leaves)
    idx = [leaf_index_of[v["vc_id"]] for v in execution_vcs]
    paths = tree.inclusion_proofs(idx) if idx else []
    cps = []
    for m in checkpoints:
        sub = MerkleTree(np.frombuffer(b"".join(leaf_hashes[:m]), dtype=np.uint8).reshape(-1, 32), ctx)
        cps.append({"tree_size": int(m), "root_hash": _b64(sub.root), "consistency_proof": [_b64(x) for x in tree.consistency_proof(m)]})
        sub.close()
    tree.close()
    ts = generated_at
    sig = sign_tree_head(tree_head_bytes(n, root, ts))

    def record(v):
        r = {k: v[k] for k in ("vc_id", "execution_id", "workflow_id", "session_id", "issuer_did", "target_did", "caller_did", "signature",
                               "input_hash", "output_hash", "status")}
        r["vc_document"] = json.loads(v["vc_document"])                 # json.RawMessage: embedded as the JSON value it is
        r["storage_uri"], r["document_size_bytes"], r["created_at"] = "", len(v["vc_document"]), v.get("created_at", ts)
        return r
    wf = None
    if workflow_vc is not None:
        wf = {k: workflow_vc[k] for k in ("workflow_id", "session_id", "component_vcs", "workflow_vc_id", "status", "start_time", "end_time",
                                          "total_steps", "completed_steps", "signature", "issuer_did", "snapshot_time", "storage_uri",
                                          "document_size_bytes")}
        wf["vc_document"] = json.loads(workflow_vc["vc_document"])
    dids = {v["issuer_did"] for v in execution_vcs} | ({workflow_vc["issuer_did"]} if workflow_vc else set()) | {head_issuer}
    bundle = {
        "workflow_id": workflow_id, "generated_at": generated_at, "total_executions": len(execution_vcs),
        "completed_executions": sum(1 for v in execution_vcs if normalize_execution_status(v["status"]) == "succeeded"),
        "workflow_status": workflow_vc["status"] if workflow_vc else "", "execution_vcs": [record(v) for v in execution_vcs],
        "workflow_vc": wf,
        "did_resolution_bundle": {d: {"did": d, "method": "key", "public_key_jwk": public_key_jwk(public_keys[d]), "resolved_from": "bundled",
                                      "resolved_at": generated_at} for d in sorted(dids)},
        "verification_metadata": {"export_version": "1.1+audit-log", "total_signatures": len(execution_vcs) + (1 if workflow_vc else 0) + 1,
                                  "bundled_dids": len(dids), "export_timestamp": generated_at},
        "audit_log": {"format": FORMAT, "leaf": "json.Marshal(VCDocument with proof)",
                      "signed_tree_head": {"tree_size": n, "root_hash": _b64(root), "timestamp": ts, "issuer": head_issuer, "signature": _b64(sig)},
                      "entries": [{"vc_id": v["vc_id"], "leaf_index": int(i), "inclusion_proof": [_b64(x) for x in p]}
                                  for v, i, p in zip(execution_vcs, idx, paths)],
                      "checkpoints": cps},
    }
    return bundle


def verify_bundle(bundle, ctx=None, trusted_checkpoints=None, verified_at=""):
    """`af vc verify <file>` for an export bundle, batched on the GPU (module docstring).  `bundle`: dict or JSON bytes/str.
    trusted_checkpoints: {tree_size: root bytes} the auditor already holds — a checkpoint of the bundle with a matching size must
    carry that very root (otherwise the log was rewritten).  Returns the reference's VCVerificationResult shape + "audit_log"."""
    ctx = ctx or default_context()
    res = {"valid": False, "type": "", "signature_valid": False, "format_valid": False, "message": "", "verified_at": verified_at,
           "component_results": [], "did_resolutions": [], "verification_steps": [],
           "summary": {"total_components": 0, "valid_components": 0, "total_dids": 0, "resolved_dids": 0, "total_signatures": 0, "valid_signatures": 0}}

    def step(n, desc, ok, details="", error=""):
        res["verification_steps"].append({"step": n, "description": desc, "success": bool(ok), "details": details, "error": error})
    if isinstance(bundle, (bytes, bytearray, str)):
        try:
            bundle = json.loads(bundle, parse_int=float)
        except ValueError as ex:
            step(2, "Parsing VC structure", False, error="Invalid VC format: %s" % ex)
            res["error"] = "Invalid VC format: not a recognized AgentField VC structure"
            return res
    elif isinstance(bundle, dict):
        bundle = json.loads(json.dumps(bundle), parse_int=float)      # work on a copy: documents are normalised in place below
    if not isinstance(bundle, dict) or not bundle.get("workflow_id"):
        step(2, "Parsing VC structure", False, error="Invalid VC format: not a recognized AgentField VC structure")
        res["error"] = "Invalid VC format: not a recognized AgentField VC structure"
        return res
    evs = bundle.get("execution_vcs") or []
    wf = bundle.get("workflow_vc")
    res.update(type="workflow", workflow_id=bundle["workflow_id"], format_valid=True)
    step(2, "Parsing VC structure", True, "Parsed enhanced VC chain with %d execution VCs" % len(evs))

    # ---- steps 3-4: unique issuer DIDs, resolved from the bundle (vc.go:331-387)
    docs = []
    for v in evs:
        d = v.get("vc_document")
        if isinstance(d, (bytes, bytearray, str)):
            try:
                d = json.loads(d, parse_int=float)
            except ValueError:
                d = None
        docs.append(d if isinstance(d, dict) else None)
    wf_doc = wf.get("vc_document") if isinstance(wf, dict) else None
    if isinstance(wf_doc, (bytes, bytearray, str)):
        wf_doc = json.loads(wf_doc, parse_int=float)
    log = bundle.get("audit_log") or {}
    sth = log.get("signed_tree_head") or {}
    order = []
    for d in docs + [wf_doc]:
        if d and d.get("issuer") not in order:
            order.append(d["issuer"])
    if sth.get("issuer") and sth["issuer"] not in order:
        order.append(sth["issuer"])
    keys = {}
    for did in order:
        entry = (bundle.get("did_resolution_bundle") or {}).get(did)
        r = {"did": did, "method": did.split(":")[1] if did.count(":") >= 2 else "unknown", "resolved_from": "", "success": False}
        try:
            if entry is None:
                raise ValueError("DID resolution failed: no resolution method available for %s" % did)
            pk = _b64url_decode_strict(entry["public_key_jwk"]["x"])
            if len(pk) != 32:
                raise ValueError("ed25519: bad public key length: %d" % len(pk))      # Go would panic inside ed25519.Verify
            keys[did] = pk
            r.update(success=True, resolved_from="bundled")
        except (KeyError, TypeError, ValueError) as ex:
            r["error"] = str(ex)
        res["did_resolutions"].append(r)
    step(3, "Collecting unique DIDs", True, "Found %d unique DIDs" % len(order))
    step(4, "Resolving DIDs", bool(keys), "Resolved %d/%d DIDs" % (len(keys), len(order)), "" if keys else "Failed to resolve any DIDs")

    # ---- step 5a: the per-credential consistency checks of verifyExecutionVCComprehensive (cheap, host) and ONE signature batch
    comps, msgs, sigs, dids, owner = [], [], [], [], []
    for v, d in zip(evs, docs):
        c = {"vc_id": v.get("vc_id", ""), "execution_id": v.get("execution_id", ""), "issuer_did": v.get("issuer_did", ""), "valid": True,
             "signature_valid": False, "format_valid": True, "status": v.get("status", ""), "error": ""}
        comps.append(c)
        if d is None:
            c.update(valid=False, format_valid=False, error="Failed to parse VC document")
            continue
        try:
            cs = d["credentialSubject"]
            for label, a, b in (("Issuer DID", v.get("issuer_did"), d["issuer"]), ("Execution ID", v.get("execution_id"), cs["executionId"]),
                                ("Workflow ID", v.get("workflow_id"), cs["workflowId"]), ("Session ID", v.get("session_id"), cs["sessionId"]),
                                ("Caller DID", v.get("caller_did"), cs["caller"]["did"]), ("Target DID", v.get("target_did"), cs["target"]["did"]),
                                ("Input hash", v.get("input_hash"), cs["execution"]["inputHash"]),
                                ("Output hash", v.get("output_hash"), cs["execution"]["outputHash"]),
                                ("Signature", v.get("signature"), d["proof"]["proofValue"])):
                if a != b:
                    raise ValueError("%s mismatch: metadata=%s, vc_document=%s" % (label, a, b))
            if normalize_execution_status(v.get("status", "")) != normalize_execution_status(cs["execution"]["status"]):
                raise ValueError("Status mismatch: metadata=%s, vc_document=%s" % (v.get("status"), cs["execution"]["status"]))
            if d["issuer"] not in keys:
                raise ValueError("DID resolution failed for issuer %s" % d["issuer"])
            sig = _b64url_decode_strict(d["proof"]["proofValue"])
            cs["audit"]["metadata"] = go_json.unmarshal_numbers(cs["audit"].get("metadata"))
            cs["execution"].setdefault("errorMessage", "")
            msgs.append(go_json.vc_document(d)); sigs.append(sig); dids.append(d["issuer"]); owner.append(c)
        except (KeyError, TypeError, ValueError) as ex:
            c.update(valid=False, error=str(ex))
    wf_res = None
    if wf_doc:
        wf_res = {"valid": False, "signature_valid": False, "error": ""}
        try:
            if wf_doc["issuer"] not in keys:
                raise ValueError("DID resolution failed for issuer %s" % wf_doc["issuer"])
            wf_doc["credentialSubject"]["audit"]["metadata"] = go_json.unmarshal_numbers(wf_doc["credentialSubject"]["audit"].get("metadata"))
            msgs.append(go_json.workflow_vc_document(wf_doc)); sigs.append(_b64url_decode_strict(wf_doc["proof"]["proofValue"]))
            dids.append(wf_doc["issuer"]); owner.append(wf_res)
        except (KeyError, TypeError, ValueError) as ex:
            wf_res["error"] = str(ex)
    sth_res = None
    if sth:
        sth_res = {"valid": False, "signature_valid": False, "error": ""}
        try:
            root = _b64url_decode_strict(sth["root_hash"])
            if len(root) != 32 or sth["issuer"] not in keys:
                raise ValueError("signed tree head: bad root or unresolved issuer")
            msgs.append(tree_head_bytes(int(sth["tree_size"]), root, sth["timestamp"])); sigs.append(_b64url_decode_strict(sth["signature"]))
            dids.append(sth["issuer"]); owner.append(sth_res)
        except (KeyError, TypeError, ValueError) as ex:
            sth_res["error"] = str(ex)
    if msgs:
        uniq = sorted(set(dids))
        ks = KeySet([keys[d] for d in uniq], ctx)
        pos = {d: i for i, d in enumerate(uniq)}
        sg = np.zeros((len(msgs), 64), dtype=np.uint8)
        good_len = [len(s) == 64 for s in sigs]                        # ed25519.Verify: a signature of another length is simply false
        for i, s in enumerate(sigs):
            if good_len[i]:
                sg[i] = np.frombuffer(s, dtype=np.uint8)
        buf, off = pack(msgs)
        ok = ks.verify_packed(np.array([pos[d] for d in dids], dtype=np.uint32), sg, buf, off)
        ks.close()
        for o, g_, own in zip(ok, good_len, owner):
            own["signature_valid"] = bool(o) and g_
            if not own["signature_valid"]:
                own["valid"] = False
                own["error"] = own["error"] or "signature verification failed"
            elif own is wf_res or own is sth_res:
                own["valid"] = True

    # ---- step 5b: the audit log — inclusion of every credential under the signed tree head, consistency with earlier checkpoints
    audit = {"present": bool(log), "format_ok": log.get("format") == FORMAT if log else False, "tree_head": sth_res, "included": 0,
             "not_included": [], "checkpoints_ok": True, "checkpoints": []}
    if log and sth_res is not None and not sth_res.get("error"):
        root, n = _b64url_decode_strict(sth["root_hash"]), int(sth["tree_size"])
        entries = {e["vc_id"]: e for e in log.get("entries") or []}
        leaves, idx, proofs, who = [], [], [], []
        for v, d, c in zip(evs, docs, comps):
            e = entries.get(v.get("vc_id"))
            if e is None or d is None:
                audit["not_included"].append(v.get("vc_id", ""))
                continue
            try:
                leaves.append(b"\x00" + _remarshal_stored(d)); idx.append(int(e["leaf_index"]))
                proofs.append([_b64url_decode_strict(x) for x in e["inclusion_proof"]]); who.append((v.get("vc_id", ""), c))
                if any(len(x) != 32 for x in proofs[-1]):
                    raise ValueError("proof node of the wrong size")
            except (KeyError, TypeError, ValueError):
                if len(leaves) > len(who):
                    leaves.pop(); idx.pop()
                if len(proofs) > len(who):
                    proofs.pop()
                audit["not_included"].append(v.get("vc_id", ""))
        if leaves:
            lh = Hasher(ctx).sha256_batch(leaves)
            inc = verify_inclusion_batch(lh, idx, n, proofs, root, ctx)
            for o, (vid, c) in zip(inc, who):
                if o:
                    audit["included"] += 1
                else:
                    audit["not_included"].append(vid)
        cps = log.get("checkpoints") or []
        if cps:
            try:
                sizes = [int(c_["tree_size"]) for c_ in cps]
                roots = [_b64url_decode_strict(c_["root_hash"]) for c_ in cps]
                prf = [[_b64url_decode_strict(x) for x in c_["consistency_proof"]] for c_ in cps]
                okc = verify_consistency_batch(sizes, roots, n, root, prf, ctx)
            except (KeyError, TypeError, ValueError):
                sizes, roots, okc = [], [], []
                audit["checkpoints_ok"] = False
            for m, r, o in zip(sizes, roots, okc):
                trusted = (trusted_checkpoints or {}).get(m)
                good = bool(o) and (trusted is None or trusted == r)
                audit["checkpoints"].append({"tree_size": m, "consistent": bool(o), "matches_trusted_root": None if trusted is None else trusted == r})
                audit["checkpoints_ok"] = audit["checkpoints_ok"] and good
        for vid in audit["not_included"]:
            for c in comps:
                if c["vc_id"] == vid:
                    c["valid"] = False
                    c["error"] = c["error"] or "not included in the audit log under the signed tree head"
    res["audit_log"] = audit

    # ---- summary (vc.go:300-331)
    res["component_results"] = comps
    total_sigs = len(evs) + (1 if wf_doc else 0) + (1 if sth else 0)
    valid_sigs = sum(c["signature_valid"] for c in comps) + int(bool(wf_res and wf_res["signature_valid"])) + int(bool(sth_res and sth_res["signature_valid"]))
    res["workflow_verification"] = wf_res
    res["summary"] = {"total_components": len(evs), "valid_components": sum(c["valid"] for c in comps), "total_dids": len(order),
                      "resolved_dids": len(keys), "total_signatures": total_sigs, "valid_signatures": valid_sigs}
    audit_ok = (not log) or (bool(sth_res and sth_res["valid"]) and not audit["not_included"] and audit["checkpoints_ok"] and audit["format_ok"])
    res["signature_valid"] = valid_sigs == total_sigs
    res["valid"] = all(c["valid"] for c in comps) and (wf_res is None or wf_res["valid"]) and audit_ok and bool(keys or not order)
    bad = sum(1 for c in comps if not c["valid"]) + int(bool(wf_res and not wf_res["valid"])) + int(not audit_ok)
    step(5, "Performing comprehensive verification", res["valid"], "%d/%d signatures valid, %d/%d credentials under the signed tree head"
         % (valid_sigs, total_sigs, audit["included"], len(evs)), "" if res["valid"] else "Found %d critical issues" % bad)
    res["message"] = "Workflow VC chain verified successfully" if res["valid"] else "Workflow VC chain verification failed"
    if not res["valid"]:
        res["error"] = "%d critical issues detected" % bad
    return res
