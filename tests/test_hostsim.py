"""The kernel logic (agentfield_b200/csrc/*.cuh), compiled for the CPU by tests/hostsim (TEST-ONLY; portable code
paths), against the oracle.  Catches arithmetic / padding / layout mistakes before GPU time is spent; the
inline-PTX paths are checked on the device by afc_selftest (tests/test_gpu_parity.py)."""
import ctypes as C
import hashlib
import hmac

import numpy as np

from conftest import golden
from oracle import c_oracle as CO, go_ed25519 as G, merkle as M


def _b(n):
    return (C.c_uint8 * n)()


def _placed(raw, al):
    """Copy raw into a 16B-aligned scratch at byte offset al; returns (keepalive, pointer)."""
    base = np.zeros(len(raw) + 64, dtype=np.uint8)
    a0 = (-base.ctypes.data) % 16
    base[a0 + al:a0 + al + len(raw)] = np.frombuffer(raw, dtype=np.uint8) if raw else []
    return base, C.cast(base.ctypes.data + a0 + al, C.POINTER(C.c_uint8))


def test_hash_padding_and_alignment(hostsim):
    rng = np.random.default_rng(5)
    for n in list(range(0, 140)) + [255, 256, 257, 511, 512, 513, 1300]:
        for al in (0, 1, 2, 3):
            m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            keep, p = _placed(m, al)
            o = _b(32); hostsim.hs_sha256(p, C.c_uint64(n), o)
            assert bytes(o) == hashlib.sha256(m).digest(), (n, al)
            o = _b(32); hostsim.hs_merkle_leaf(p, C.c_uint64(n), o)
            assert bytes(o) == M.leaf_hash(m), (n, al)
            for kl in (0, 5, 32, 64, 65, 131):
                k = rng.integers(0, 256, kl, dtype=np.uint8).tobytes()
                o = _b(32); hostsim.hs_hmac_sha256(k, C.c_uint32(kl), p, C.c_uint64(n), o)
                assert bytes(o) == hmac.new(k, m, hashlib.sha256).digest(), (n, al, kl)
            pre = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
            o = _b(64); hostsim.hs_sha512_pre64(pre, p, C.c_uint64(n), o)
            assert bytes(o) == hashlib.sha512(pre + m).digest(), (n, al)
            o = _b(64); hostsim.hs_sha512_pre32(pre[:32], p, C.c_uint64(n), o)
            assert bytes(o) == hashlib.sha512(pre[:32] + m).digest(), (n, al)
    l, r = bytes(range(32)), bytes(range(32, 64))
    o = _b(32); hostsim.hs_merkle_node(l, r, o)
    assert bytes(o) == M.node_hash(l, r)


def test_scalar_arithmetic(hostsim):
    rng = np.random.default_rng(6)
    L = G.L
    specials = [b"\xff" * 64, bytes(64)] + [int(L * k + d).to_bytes(64, "little") for k in (1, 2, 2**250) for d in (-1, 0, 1)]
    for i in range(1500):
        x = specials[i] if i < len(specials) else rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        o = _b(32); hostsim.hs_sc_reduce512(x, o)
        assert int.from_bytes(bytes(o), "little") == int.from_bytes(x, "little") % L
        a, b, c = (rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(3))
        o = _b(32); hostsim.hs_sc_muladd(a, b, c, o)
        assert int.from_bytes(bytes(o), "little") == (int.from_bytes(a, "little") * int.from_bytes(b, "little") + int.from_bytes(c, "little")) % L


def test_field_arithmetic(hostsim):
    P = G.P
    rng = np.random.default_rng(8)
    edge = [0, 1, 2, 19, 37, 38, P - 1, P, P + 1, 2**255 - 1, 2**255, 2**256 - 1, 2**256 - 38, 2**256 - 39, 2**256 - 37, 2**32 - 1, 2**64 - 1]
    vals = edge + [int.from_bytes(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), "little") for _ in range(200)]

    def op(code, x, y=0):
        o = _b(32); hostsim.hs_fe_op(code, int(x).to_bytes(32, "little"), int(y).to_bytes(32, "little"), o)
        return int.from_bytes(bytes(o), "little")

    for x in vals[:30]:
        for y in vals[:30]:
            assert op(0, x, y) == x * y % P and op(2, x, y) == (x + y) % P and op(3, x, y) == (x - y) % P
        assert op(1, x) == x * x % P and op(5, x) == x % P
        if x % P:
            assert op(4, x) == pow(x, P - 2, P)
    for x, y in zip(vals[17:], reversed(vals[17:])):
        assert op(0, x, y) == x * y % P and op(1, x) == x * x % P and op(2, x, y) == (x + y) % P and op(3, x, y) == (x - y) % P


def test_ed25519_kernel_logic_vs_golden_and_oracle(hostsim):
    for e in golden("rfc8032.json"):
        seed, pk, msg, sig = (bytes.fromhex(e[k]) for k in ("seed", "pk", "msg", "sig"))
        o = _b(32); hostsim.hs_pubkey(seed, o); assert bytes(o) == pk, e["name"]
        o = _b(64); hostsim.hs_sign(seed, msg, C.c_uint64(len(msg)), o); assert bytes(o) == sig, e["name"]
        o = _b(64); hostsim.hs_sign_grouped(seed, msg, C.c_uint64(len(msg)), o); assert bytes(o) == sig, e["name"]   # k_ed_sign's split
        assert hostsim.hs_verify(pk, msg, C.c_uint64(len(msg)), sig) == 1
    for e in golden("ed25519_edge.json"):
        pk, msg, sig = (bytes.fromhex(e[k]) for k in ("pk", "msg", "sig"))
        assert hostsim.hs_verify(pk, msg, C.c_uint64(len(msg)), sig) == int(e["valid"]), e["name"]
    rng = np.random.default_rng(10)
    for i in range(150):
        sd = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
        m = rng.integers(0, 256, (i * 7) % 700, dtype=np.uint8).tobytes()
        s, pk = CO.sign(sd, m), CO.pubkey(sd)
        o = _b(64); hostsim.hs_sign(sd, m, C.c_uint64(len(m)), o); assert bytes(o) == s
        assert hostsim.hs_verify(pk, m, C.c_uint64(len(m)), s) == 1
        bs = bytearray(s); bs[i % 64] ^= 1 << (i % 8)
        assert hostsim.hs_verify(pk, m, C.c_uint64(len(m)), bytes(bs)) == 0


def test_keyed_table_path_logic_vs_golden(hostsim):
    """Per-key radix-256 table build (chunked Montgomery inversion) + table-driven verification, on the CPU build."""
    es = golden("ed25519_edge.json") + [dict(e, valid=True) for e in golden("rfc8032.json")]
    for e in es:
        pk, msg, sig = (bytes.fromhex(e[k]) for k in ("pk", "msg", "sig"))
        assert hostsim.hs_verify_keyed(pk, msg, C.c_uint64(len(msg)), sig) == int(e["valid"]), e["name"]


def test_base_point_table_chunks(hostsim):
    """What each device thread builds at afc_init (its own start point by double-and-add, then a run of consecutive
    multiples) equals the row-by-row construction the other hostsim checks (sign / verify vs RFC 8032) run on."""
    w = hostsim.hs_base_window()
    assert w in (8, 16)
    rows, cols = 256 // w, 1 << (w - 1)
    for i, j0 in [(0, 0), (0, cols - 64), (1, 64), (rows // 2, (cols // 2) & ~63), (rows - 1, cols - 64)]:
        assert hostsim.hs_base_chunk_mismatches(i, j0) == 0, (i, j0)


def test_group_encoding_shares_one_inversion(hostsim):
    for G in range(1, 9):
        assert hostsim.hs_encode_group_mismatches(G, 1234 + G) == 0, G


def test_key_table_two_step_build_equals_single_thread_rows(hostsim):
    """Row base points from one doubling chain (T dropped between doublings) + rows from their bases == each row from A."""
    pk = bytes.fromhex(golden("rfc8032.json")[0]["pk"])
    for i in (0, 1, 7, 31):
        assert hostsim.hs_key_row_mismatches(pk, i) == 0, i


def test_key_table_staged_chain_and_helper_starts_equal_single_thread_rows(hostsim):
    """Round-2 table build: staged doubling chain leaving P, 32 P, 64 P per row, slices started from the helpers,
    forward / inversion / backward run — equals each row built from A by one thread, for every staging."""
    for e in golden("rfc8032.json")[:2]:
        pk = bytes.fromhex(e["pk"])
        for stages in (1, 2, 4):
            for row in (0, 7, 8, 19, 31):
                assert hostsim.hs_key_table_staged_mismatches(pk, stages, row) == 0, (stages, row)
    # a key that does not decode: flagged, nothing else to compare
    bad = next(bytes.fromhex(e["pk"]) for e in golden("ed25519_edge.json") if "not on curve" in e["name"])
    assert hostsim.hs_key_table_staged_mismatches(bad, 4, 3) == 0


def test_go_json_escaping_and_template_fill(hostsim):
    """Device routine for Go encoding/json string escaping (afc_json.cuh) vs the byte-level oracle: every single byte, every
    byte pair around the UTF-8 boundaries, random strings at every alignment, and documents assembled from a template."""
    from oracle import go_json as OJ
    hostsim.hs_json_escape.restype = C.c_uint64
    hostsim.hs_json_fill_one.restype = C.c_uint64

    def dev_escape(raw, al=0):
        keep, ptr = _placed(raw, al)
        n = hostsim.hs_json_escape(ptr, C.c_uint64(len(raw)), None)
        out = np.zeros(n + 8, dtype=np.uint8)
        a0 = (-out.ctypes.data) % 4
        assert hostsim.hs_json_escape(ptr, C.c_uint64(len(raw)), C.cast(out.ctypes.data + a0 + (al % 4), C.POINTER(C.c_uint8))) == n
        return out[a0 + (al % 4):a0 + (al % 4) + n].tobytes()

    for b in range(256):
        assert dev_escape(bytes([b])) == OJ.escape_bytes(bytes([b])), b
    edge = [0x00, 0x1f, 0x20, 0x22, 0x5c, 0x7f, 0x80, 0x9f, 0xa0, 0xbf, 0xc0, 0xc1, 0xc2, 0xdf, 0xe0, 0xe2, 0xec, 0xed, 0xee, 0xef, 0xf0, 0xf4, 0xf5, 0xff,
            0x8f, 0x90, 0xa8, 0xa9]
    for a in edge:
        for b in edge:
            for tail in (b"", b"\x80", b"\x80\x80", b"\xa8", b"\xbf\xbf\x41"):
                raw = bytes([a, b]) + tail
                assert dev_escape(raw) == OJ.escape_bytes(raw), raw
    rng = np.random.default_rng(0xAF34)
    pool = [bytes([i]) for i in range(128)] + ["é".encode(), "中".encode(), "😀".encode(), b"\xe2\x80\xa8", b"\xe2\x80\xa9", b"\xff", b"\xc3", b"\xed\xa0\x80"]
    for t in range(300):
        raw = b"".join(pool[int(i)] for i in rng.integers(0, len(pool), int(rng.integers(0, 60))))
        assert dev_escape(raw, t % 16) == OJ.escape_bytes(raw), raw
    # long plain runs (the four-bytes-at-a-time path) broken by single special / multi-byte characters, every input alignment and
    # every phase of the output word
    plain = bytes(b for b in range(0x20, 0x80) if b not in b'"\\<>&')
    for t in range(400):
        parts = []
        for _ in range(int(rng.integers(1, 6))):
            parts.append(bytes(plain[int(i)] for i in rng.integers(0, len(plain), int(rng.integers(0, 41)))))
            parts.append(pool[int(rng.integers(0, len(pool)))] if rng.integers(0, 4) else b"")
        raw = b"".join(parts)
        assert dev_escape(raw, t % 16) == OJ.escape_bytes(raw), raw
    for b in range(256):                                   # one odd byte at each position of an otherwise plain word
        for pos in range(4):
            raw = b"abcdefgh"[:pos] + bytes([b]) + b"ijklmnop"
            assert dev_escape(raw, pos) == OJ.escape_bytes(raw), raw
    # template: {"a":"<v0>","n":<v1>,"b":"<v2>"}
    segs = [b'{"a":"', b'","n":', b',"b":"', b'"}']
    kinds = [OJ.STRING, OJ.RAW, OJ.STRING]
    sb = np.frombuffer(b"".join(segs), dtype=np.uint8).copy()
    so = np.zeros(len(segs) + 1, dtype=np.uint32); so[1:] = np.cumsum([len(x) for x in segs])
    kk = np.array(kinds, dtype=np.uint8)
    for t in range(100):
        vals = [b"".join(pool[int(i)] for i in rng.integers(0, len(pool), int(rng.integers(0, 30)))), str(int(rng.integers(0, 10**9))).encode(),
                b"".join(pool[int(i)] for i in rng.integers(0, len(pool), int(rng.integers(0, 30))))]
        fb = np.frombuffer(b"".join(vals) + b"\0", dtype=np.uint8).copy()
        fo = np.zeros(len(vals) + 1, dtype=np.uint64); fo[1:] = np.cumsum([len(x) for x in vals])
        want = OJ.fill_template(segs, kinds, vals)
        args = (sb.ctypes.data_as(C.c_void_p), so.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p), C.c_uint32(len(vals)),
                fb.ctypes.data_as(C.c_void_p), fo.ctypes.data_as(C.c_void_p))
        n = hostsim.hs_json_fill_one(*args, None)
        assert n == len(want)
        out = np.zeros(n + 4, dtype=np.uint8)
        hostsim.hs_json_fill_one(*args, out.ctypes.data_as(C.c_void_p))
        assert out[:n].tobytes() == want


def _py_sha256_compress(h, block):
    """FIPS 180-4 §6.2.2 in plain Python: the independent check of the chaining values inside a streaming state."""
    K = [0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
         0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
         0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
         0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
         0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
         0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
    M = 0xffffffff
    r = lambda x, n: ((x >> n) | (x << (32 - n))) & M
    w = [int.from_bytes(block[4 * i:4 * i + 4], "big") for i in range(16)]
    for i in range(16, 64):
        s0 = r(w[i - 15], 7) ^ r(w[i - 15], 18) ^ (w[i - 15] >> 3); s1 = r(w[i - 2], 17) ^ r(w[i - 2], 19) ^ (w[i - 2] >> 10)
        w.append((w[i - 16] + s0 + w[i - 7] + s1) & M)
    a, b, c, d, e, f, g, hh = h
    for i in range(64):
        t1 = (hh + (r(e, 6) ^ r(e, 11) ^ r(e, 25)) + ((e & f) ^ (~e & g & M)) + K[i] + w[i]) & M
        t2 = ((r(a, 2) ^ r(a, 13) ^ r(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & M
        hh, g, f, e, d, c, b, a = g, f, e, (d + t1) & M, c, b, a, (t1 + t2) & M
    return [(x + y) & M for x, y in zip(h, (a, b, c, d, e, f, g, hh))]


def test_streaming_sha256_state_is_go_marshal_binary_layout(hostsim):
    """H2 (payload_store.go:69-94): chunked absorption == hashlib over the whole stream for random chunkings and alignments, and the
    108-byte state between chunks is "sha\\x03" || h big-endian || 64 zero bytes || length big-endian — crypto/sha256's MarshalBinary
    form for a digest sitting on a block boundary — with h checked by an independent FIPS 180-4 compression."""
    import ctypes as C
    import hashlib
    rng = np.random.default_rng(0xAF68)
    iv = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    for trial in range(60):
        total = int(rng.integers(0, 5000))
        data = rng.integers(0, 256, total + 16, dtype=np.uint8).tobytes()
        shift = int(rng.integers(0, 16))                              # chunk pointers at every alignment
        data = data[shift:shift + total]
        state = (C.c_uint8 * 108)(*(b"sha\x03" + b"".join(x.to_bytes(4, "big") for x in iv) + bytes(64) + bytes(8)))
        out = (C.c_uint8 * 32)()
        pos, h = 0, list(iv)
        while total - pos >= 64 and rng.random() < 0.8:
            k = 64 * int(rng.integers(1, max(2, (total - pos) // 64 + 1)))
            chunk = data[pos:pos + k]
            assert hostsim.hs_sha256_stream_update(state, chunk, len(chunk), 0, out) == 1
            for b in range(0, k, 64):
                h = _py_sha256_compress(h, chunk[b:b + 64])
            pos += k
            assert bytes(state) == b"sha\x03" + b"".join(x.to_bytes(4, "big") for x in h) + bytes(64) + pos.to_bytes(8, "big")
        assert hostsim.hs_sha256_stream_update(state, data[pos:], total - pos, 1, out) == 1
        assert bytes(out) == hashlib.sha256(data).digest(), (trial, total, pos)
    bad = (C.c_uint8 * 108)(*bytes(108))
    assert hostsim.hs_sha256_stream_update(bad, b"", 0, 1, out) == 0                     # not a state
    ok = (C.c_uint8 * 108)(*(b"sha\x03" + b"".join(x.to_bytes(4, "big") for x in iv) + bytes(72)))
    assert hostsim.hs_sha256_stream_update(ok, b"x" * 65, 65, 0, out) == 0              # ragged chunk in the middle of a stream


def test_constant_time_base_multiplication_equals_variable_time(hostsim):
    """ge_scalarmult_base_ct (signed radix 16, masked scans, neutral element for zero digits, conditional negation) computes the
    same point as the radix-65536 gather path — for random scalars and for the digit patterns that exercise every branch-free
    case: zero, all digits 0x8 (-8 everywhere), all 0x7, all 0xf, a single top digit, L - 1."""
    rng = np.random.default_rng(0xAF70)
    cases = [bytes(32), bytes([0x88] * 32), bytes([0x77] * 32), bytes([0xff] * 32), bytes(31) + b"\x0f", b"\x01" + bytes(31),
             (2**252 + 27742317777372353535851937790883648493 - 1).to_bytes(32, "little")]
    cases += [rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(40)]
    for c in cases:
        assert hostsim.hs_scalarmult_ct_mismatches(c) == 0, c.hex()


def test_quad_lane_mixed_addition_plan(hostsim):
    """The data the four-lane kernel runs on (ge_quad_plan: who shuffles what from whom, which sign) replayed on the CPU against the
    one-thread mixed addition, and the one-instruction-stream add/subtract against Python integers."""
    rng = np.random.default_rng(44)
    q = 2**255 - 19
    edge = [0, 1, 37, 38, 39, 2**256 - 1, 2**256 - 38, 2**256 - 39, q, q - 1, q + 1, 2**255, 19]
    vals = edge + [int.from_bytes(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), "little") for _ in range(60)]
    for a in vals:
        for b in vals:
            for op, want in ((5, a + b), (6, a - b), (7, a)):
                o = _b(32); hostsim.hs_fe_op(op, a.to_bytes(32, "little"), b.to_bytes(32, "little"), o)
                assert int.from_bytes(bytes(o), "little") == want % q, (op, a, b)
    for trial in range(6):
        steps = 48
        d = rng.integers(-128, 128, steps, dtype=np.int8)
        d[rng.integers(0, steps, 5)] = 0
        d[0] = [0, 1, -1, 127, -128, 5][trial]
        s = rng.integers(0, 256, 32, dtype=np.uint8); s[31] &= 0x0f
        assert hostsim.hs_quad_madd_mismatches(s.tobytes(), d.ctypes.data_as(C.POINTER(C.c_int8)), steps) == 0, trial
