"""Byte-exact restatement of Go `encoding/json.Marshal` for the structs that ARE the canonical signing form on the hot path
(control-plane/pkg/types/did_types.go:135-220, pkg/types/webhook.go:42-53).  Host-side string work: in the real
deployment this stays in Go; the Python mirror needs it so that the issue -> verify flow of
BASELINE.json configs[0] can be driven end to end through the GPU path with the same bytes Go would sign.

Rules restated (Go 1.24): struct fields in declaration order under their json tags; `omitempty` drops "" / nil;
strings escape `"` `\\` and control characters (\\b \\f \\n \\r \\t short forms, others \\u00XX), and — because Marshal
uses escapeHTML — `<` `>` `&` as \\u003c \\u003e \\u0026, plus U+2028 / U+2029; nil slices are `null`; map keys sorted;
float64 uses the shortest representation, 'e' form when exp < -6 or >= 21.
"""
import math

_ESC = {'"': '\\"', "\\": "\\\\", "\b": "\\b", "\f": "\\f", "\n": "\\n", "\r": "\\r", "\t": "\\t",
        "<": "\\u003c", ">": "\\u003e", "&": "\\u0026", "\u2028": "\\u2028", "\u2029": "\\u2029"}


def string(s: str) -> str:
    out = ['"']
    for ch in s:
        e = _ESC.get(ch)
        if e is not None:
            out.append(e)
        elif ch < " ":
            out.append("\\u%04x" % ord(ch))
        elif 0xD800 <= ord(ch) <= 0xDFFF:
            out.append("\\ufffd")            # invalid UTF-8 in Go becomes U+FFFD
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def number(x) -> str:
    if isinstance(x, bool):
        return "true" if x else "false"
    if isinstance(x, int):
        return str(x)
    if math.isnan(x) or math.isinf(x):
        raise ValueError("json: unsupported value: %r" % x)    # Go returns UnsupportedValueError
    if x == 0:
        return "-0" if math.copysign(1, x) < 0 else "0"
    a = abs(x)
    if a < 1e-6 or a >= 1e21:
        m, e = ("%r" % x).split("e") if "e" in ("%r" % x) else (None, None)
        if m is None:
            r = "%.17e" % x
            m, e = r.split("e")
            m = repr(float(m)).rstrip("0").rstrip(".")
        m = m[:-2] if m.endswith(".0") else m
        sign = "-" if e.startswith("-") else "+"
        return "%se%s%d" % (m, sign, abs(int(e)))
    r = repr(float(x))
    if "e" in r or "E" in r:
        r = ("%.17f" % x).rstrip("0")
    return r[:-2] if r.endswith(".0") else r


def value(v) -> str:
    if v is None:
        return "null"
    if isinstance(v, str):
        return string(v)
    if isinstance(v, (bool, int, float)):
        return number(v)
    if isinstance(v, dict):
        return "{" + ",".join(string(k) + ":" + value(v[k]) for k in sorted(v)) + "}"
    if isinstance(v, (list, tuple)):
        return "[" + ",".join(value(x) for x in v) + "]"
    raise TypeError("unsupported JSON value %r" % type(v))


def string_list(xs) -> str:
    return "null" if xs is None else "[" + ",".join(string(x) for x in xs) + "]"


def vc_proof(p) -> str:
    return ('{"type":%s,"created":%s,"verificationMethod":%s,"proofPurpose":%s,"proofValue":%s}'
            % (string(p.get("type", "")), string(p.get("created", "")), string(p.get("verificationMethod", "")),
               string(p.get("proofPurpose", "")), string(p.get("proofValue", ""))))


EMPTY_PROOF = {"type": "", "created": "", "verificationMethod": "", "proofPurpose": "", "proofValue": ""}


def vc_document(doc, proof=None) -> bytes:
    """json.Marshal(types.VCDocument).  proof=None -> the zero-valued Proof the reference signs over (vc_service.go:436-439)."""
    cs = doc["credentialSubject"]
    ex = cs["execution"]
    execution = ('{"inputHash":%s,"outputHash":%s,"timestamp":%s,"durationMs":%d,"status":%s%s}'
                 % (string(ex["inputHash"]), string(ex["outputHash"]), string(ex["timestamp"]), int(ex["durationMs"]), string(ex["status"]),
                    (',"errorMessage":' + string(ex["errorMessage"])) if ex.get("errorMessage") else ""))
    au = cs["audit"]
    audit = '{"inputDataHash":%s,"outputDataHash":%s,"metadata":%s}' % (string(au["inputDataHash"]), string(au["outputDataHash"]), value(au.get("metadata")))
    ca, ta = cs["caller"], cs["target"]
    subject = ('{"executionId":%s,"workflowId":%s,"sessionId":%s,"caller":{"did":%s,"type":%s,"agentNodeDid":%s},'
               '"target":{"did":%s,"agentNodeDid":%s,"functionName":%s},"execution":%s,"audit":%s}'
               % (string(cs["executionId"]), string(cs["workflowId"]), string(cs["sessionId"]), string(ca["did"]), string(ca["type"]),
                  string(ca["agentNodeDid"]), string(ta["did"]), string(ta["agentNodeDid"]), string(ta["functionName"]), execution, audit))
    return ('{"@context":%s,"type":%s,"id":%s,"issuer":%s,"issuanceDate":%s,"credentialSubject":%s,"proof":%s}'
            % (string_list(doc["@context"]), string_list(doc["type"]), string(doc["id"]), string(doc["issuer"]), string(doc["issuanceDate"]),
               subject, vc_proof(proof if proof is not None else EMPTY_PROOF))).encode("utf-8")


def webhook_payload(p) -> bytes:
    """json.Marshal(types.ExecutionWebhookPayload) (pkg/types/webhook.go:42-53)."""
    parts = ['"event":' + string(p["event"]), '"execution_id":' + string(p["execution_id"]), '"workflow_id":' + string(p["workflow_id"]),
             '"status":' + string(p["status"]), '"target":' + string(p["target"]), '"type":' + string(p["type"])]
    if p.get("duration_ms") is not None:
        parts.append('"duration_ms":%d' % int(p["duration_ms"]))
    if p.get("result") is not None:
        parts.append('"result":' + value(p["result"]))
    if p.get("error_message") is not None:
        parts.append('"error_message":' + string(p["error_message"]))
    parts.append('"timestamp":' + string(p["timestamp"]))
    return ("{" + ",".join(parts) + "}").encode("utf-8")
