"""Builds the host mirror's inputs (agentfield_b200.go_json / services) from tests/golden/go_cases.json entries — the same field
values the Go generator (baseline/go/gen_golden_test.go) feeds to the reference's structs."""
import json

from agentfield_b200 import go_json as GJ


def execution_doc(c):
    """types.VCDocument as the host mirror's dict, the way VerifyVC sees it: metadata through json.Unmarshal (float64 numbers),
    the error message truncated the way GenerateExecutionVC does (bytes, vc_service.go:153-160)."""
    md = GJ.unmarshal_interface(c["metadata_json"])
    return {"@context": c["context"], "type": c["type"], "id": c["id"], "issuer": c["issuer"], "issuanceDate": c["issuance_date"],
            "credentialSubject": {
                "executionId": c["execution_id"], "workflowId": c["workflow_id"], "sessionId": c["session_id"],
                "caller": {"did": c["caller"]["did"], "type": c["caller"]["type"], "agentNodeDid": c["caller"]["agent_node_did"]},
                "target": {"did": c["target"]["did"], "agentNodeDid": c["target"]["agent_node_did"], "functionName": c["target"]["function_name"]},
                "execution": {"inputHash": c["input_hash"], "outputHash": c["output_hash"], "timestamp": c["timestamp"], "durationMs": c["duration_ms"],
                              "status": c["status"], "errorMessage": GJ.truncate_error_message(c["error_message_input"]) or ""},
                "audit": {"inputDataHash": c["input_data_hash"], "outputDataHash": c["output_data_hash"], "metadata": md}}}


def proof(did, sig: bytes, created):
    import base64
    return {"type": "Ed25519Signature2020", "created": created, "verificationMethod": "%s#key-1" % did, "proofPurpose": "assertionMethod",
            "proofValue": base64.urlsafe_b64encode(sig).rstrip(b"=").decode()}


def workflow_doc(c):
    from agentfield_b200.services import create_workflow_vc_document
    return create_workflow_vc_document(c["workflow_id"], c["session_id"], c["component_vc_ids"], c["status"], c["start_time"], c["end_time"],
                                       c["issuer_did"], c["vc_id"], c["issuance_date"], c["snapshot_time"])


def webhook(c):
    p = {k: c[k] for k in ("event", "execution_id", "workflow_id", "status", "target", "type", "duration_ms", "error_message", "timestamp")}
    p["result"] = None if c["result_json"] is None else GJ.unmarshal_interface(c["result_json"])
    return p
