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
