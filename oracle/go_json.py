"""TEST INFRASTRUCTURE (oracle): byte-level restatement of Go 1.24 encoding/json string escaping and of the canonical
VCDocument bytes the reference signs.  Independent of agentfield_b200/go_json.py (which works on Python str): this one follows
Go's own loop over BYTES, so it also defines what happens to invalid UTF-8.

Follows: Go stdlib encoding/json/encode.go appendString (escapeHTML = true, as json.Marshal uses), unicode/utf8 DecodeRune;
call sites in the reference: json.Marshal(vcDoc) internal/services/vc_service.go:436-439, :201, json.Marshal(payload)
internal/services/webhook_dispatcher.go:286-300.  The reference holds no golden vectors for these bytes:
parity is pinned on the language specification of encoding/json, cross-checked against Python's json for the common subset."""

_HEX = "0123456789abcdef"


def _decode_rune(s: bytes, i: int):
    """unicode/utf8.DecodeRune(s[i:]) -> (size, valid).  Invalid or truncated -> (1, False)."""
    n = len(s) - i
    b0 = s[i]
    lo, hi = 0x80, 0xBF
    if 0xC2 <= b0 <= 0xDF:
        need = 2
    elif 0xE0 <= b0 <= 0xEF:
        need = 3
        if b0 == 0xE0:
            lo = 0xA0
        elif b0 == 0xED:
            hi = 0x9F
    elif 0xF0 <= b0 <= 0xF4:
        need = 4
        if b0 == 0xF0:
            lo = 0x90
        elif b0 == 0xF4:
            hi = 0x8F
    else:
        return 1, False
    if n < need or not (lo <= s[i + 1] <= hi):
        return 1, False
    for k in range(2, need):
        if not (0x80 <= s[i + k] <= 0xBF):
            return 1, False
    return need, True


def escape_bytes(s: bytes) -> bytes:
    """The bytes Go writes between the quotes for a string whose UTF-8 bytes are s."""
    out = bytearray()
    i = 0
    while i < len(s):
        b = s[i]
        if b < 0x80:
            if b >= 0x20 and b not in b'"\\<>&':
                out.append(b)
            elif b in b'"\\':
                out += b"\\" + bytes([b])
            elif b == 0x08:
                out += b"\\b"
            elif b == 0x0C:
                out += b"\\f"
            elif b == 0x0A:
                out += b"\\n"
            elif b == 0x0D:
                out += b"\\r"
            elif b == 0x09:
                out += b"\\t"
            else:
                out += b"\\u00" + _HEX[b >> 4].encode() + _HEX[b & 15].encode()
            i += 1
            continue
        size, ok = _decode_rune(s, i)
        if not ok:
            out += b"\\ufffd"
            i += 1
            continue
        if size == 3 and s[i:i + 3] in (b"\xe2\x80\xa8", b"\xe2\x80\xa9"):
            out += b"\\u202" + _HEX[s[i + 2] & 15].encode()
            i += 3
            continue
        out += s[i:i + size]
        i += size
    return bytes(out)


STRING, RAW = 0, 1


def fill_template(segments, kinds, values) -> bytes:
    """seg[0] v0 seg[1] ... v(F-1) seg[F]; values are bytes; STRING values are escaped, RAW copied."""
    assert len(segments) == len(kinds) + 1 == len(values) + 1
    out = bytearray(segments[0])
    for f, v in enumerate(values):
        out += escape_bytes(v) if kinds[f] == STRING else v
        out += segments[f + 1]
    return bytes(out)


def number_bytes(x: float) -> bytes:
    """json.Marshal(float64): Go stdlib encoding/json/encode.go floatEncoder.encode — strconv.AppendFloat(b, f, fmt, -1, 64) with
    fmt = 'e' iff |f| < 1e-6 or |f| >= 1e21 (else 'f'), then "e-09" -> "e-9".  Reference values that take this path:
    VCAudit.Metadata map[string]interface{} after json.Unmarshal (internal/services/vc_service.go:250-251 -> :471-474) and
    ExecutionWebhookPayload.Result (pkg/types/webhook.go:49).
    Independent of agentfield_b200/go_json.number (which reads repr() through decimal): here the shortest digit string is found
    by trying %.{p}e for p = 1..17 until it round-trips (correctly rounded printf), then laid out by hand."""
    if x != x or x in (float("inf"), float("-inf")):
        raise ValueError("json: unsupported value")
    if x == 0:
        return b"-0" if str(x)[0] == "-" else b"0"
    sign = b"-" if x < 0 else b""
    a = abs(x)
    for p in range(1, 18):
        t = "%.*e" % (p - 1, a)
        if float(t) == a:
            break
    mant, exp = t.split("e")
    digits = mant.replace(".", "").rstrip("0") or "0"
    exp = int(exp)
    if a < 1e-6 or a >= 1e21:
        body = digits[0] + ("." + digits[1:] if len(digits) > 1 else "") + "e" + ("-" if exp < 0 else "+")
        body += ("%02d" % abs(exp))
        if len(body) >= 4 and body[-4] == "e" and body[-2] == "0":       # Go's clean-up of a two-digit exponent with a leading zero
            body = body[:-2] + body[-1]
        return sign + body.encode()
    if exp >= len(digits) - 1:
        return sign + (digits + "0" * (exp - len(digits) + 1)).encode()
    if exp >= 0:
        return sign + (digits[:exp + 1] + "." + digits[exp + 1:]).encode()
    return sign + ("0." + "0" * (-exp - 1) + digits).encode()


def value_bytes(v) -> bytes:
    """json.Marshal of an interface{} tree as the reference holds it after json.Unmarshal: nil, bool, float64, string,
    []interface{}, map[string]interface{} (keys sorted bytewise, encode.go mapEncoder)."""
    if v is None:
        return b"null"
    if v is True:
        return b"true"
    if v is False:
        return b"false"
    if isinstance(v, (int, float)):
        return number_bytes(float(v))
    if isinstance(v, str):
        return b'"' + escape_bytes(v.encode("utf-8", "surrogatepass")) + b'"'
    if isinstance(v, (list, tuple)):
        return b"[" + b",".join(value_bytes(x) for x in v) + b"]"
    if isinstance(v, dict):
        ks = sorted(v, key=lambda k: k.encode("utf-8", "surrogatepass"))
        return b"{" + b",".join(b'"' + escape_bytes(k.encode("utf-8", "surrogatepass")) + b'":' + value_bytes(v[k]) for k in ks) + b"}"
    raise TypeError(type(v))
