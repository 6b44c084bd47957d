// Canonical form on the device: Go encoding/json string escaping and template assembly.
//
// The message the reference signs is json.Marshal of a fixed struct (VCDocument, pkg/types/did_types.go:135-220, built by
// createVCDocument internal/services/vc_service.go:374-431 and marshalled at :436-439): constant text interleaved with a
// fixed number of values.  Strings are written the way Go's encoding/json does (encode.go appendString with escapeHTML):
//   "  \  -> \" \\          \b \f \n \r \t -> short forms          other bytes < 0x20, and < > & -> \u00XX
//   U+2028 / U+2029 -> backslash-u 2028 / 2029      invalid UTF-8 -> backslash-u fffd, one input byte at a time (utf8.DecodeRune rules:
//   no overlongs, no surrogates, nothing above U+10FFFF, truncated sequences invalid)          everything else copied.
// Shared by the device kernels (k_hash.cu) and the CPU build of the same logic in tests/hostsim.
#pragma once
#include "afc_common.cuh"

namespace afc {

// Sequential byte reader with 4 bytes of look-ahead over an arbitrarily aligned string (aligned 32-bit loads only).
struct ByteWindow {
    MsgReader rd;
    uint64_t win;        // pending bytes, next byte in the low 8 bits
    uint32_t have;       // valid bytes in win (0..8)
    uint64_t rem;        // bytes of the string not yet consumed
    uint64_t unread;     // bytes of the string not yet pulled into win
    AFC_HDM void init(const uint8_t* p, uint64_t len) { rd.init(p, len); win = 0; have = 0; rem = len; unread = len; }
    AFC_HDM void fill() {
        while (have <= 4 && unread) {
            uint32_t w = rd.next();
            uint32_t k = unread < 4 ? (uint32_t)unread : 4u;
            if (k < 4) w &= (1u << (8 * k)) - 1u;
            win |= (uint64_t)w << (8 * have);
            have += k; unread -= k;
        }
    }
    AFC_HDM uint32_t peek(uint32_t j) const { return (uint32_t)(win >> (8 * j)) & 0xffu; }     // j < have
    AFC_HDM void skip(uint32_t k) { win = k >= 8 ? 0 : win >> (8 * k); have -= k; rem -= k; }
};

// Byte sink that writes aligned 32-byte blocks — one full DRAM sector per store (STG.256, sm_100) — with single bytes only for an
// unaligned head and words + bytes for the tail.  With one document per thread a warp's store touches 32 different lines whatever
// its width, so wider stores are fewer line accesses; and a sector written whole needs no read-for-fill (ncu, 16-byte stores:
// 0.93 GB read + 0.76 GB written for 0.42 GB of input and 0.63 GB of output).
struct ByteWriter {
    uint8_t* p;                 // next store address (once head == 0)
    uint32_t acc, n, head;      // word being assembled: n bytes in acc
    uint32_t s0, s1, s2, s3, s4, s5, s6, wi;    // words of the current block already complete (the eighth goes straight out)
    AFC_HDM void init(uint8_t* dst) {
        p = dst; acc = 0; n = 0; wi = 0; s0 = s1 = s2 = s3 = s4 = s5 = s6 = 0;
        head = (uint32_t)((32 - ((uintptr_t)dst & 31)) & 31);
    }
    AFC_HDM void word(uint32_t x) {
        if (wi == 7) {
#if AFC_DEVICE_CODE
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(s4), "r"(s5), "r"(s6), "r"(x) : "memory");
#else
            uint32_t v[8] = {s0, s1, s2, s3, s4, s5, s6, x}; memcpy(p, v, 32);
#endif
            p += 32; wi = 0; return;
        }
        s0 = wi == 0 ? x : s0; s1 = wi == 1 ? x : s1; s2 = wi == 2 ? x : s2; s3 = wi == 3 ? x : s3;
        s4 = wi == 4 ? x : s4; s5 = wi == 5 ? x : s5; s6 = wi == 6 ? x : s6;
        ++wi;
    }
    AFC_HDM void put(uint32_t b) {
        if (head) { *p++ = (uint8_t)b; --head; return; }
        acc |= b << (8 * n);
        if (++n == 4) { word(acc); acc = 0; n = 0; }
    }
    // four bytes at once (little-endian word), whatever the phase of the pending bytes
    AFC_HDM void put4(uint32_t x) {
        if (head) { put(x & 0xff); put((x >> 8) & 0xff); put((x >> 16) & 0xff); put(x >> 24); return; }
        word(acc | (n ? x << (8 * n) : x));
        acc = n ? x >> (32 - 8 * n) : 0;
    }
    AFC_HDM void finish() {
        uint32_t* q = (uint32_t*)p;
        const uint32_t st[7] = {s0, s1, s2, s3, s4, s5, s6};
        for (uint32_t k = 0; k < wi; k++) q[k] = st[k];
        uint8_t* t = p + 4 * wi;
        for (uint32_t k = 0; k < n; k++) t[k] = (uint8_t)(acc >> (8 * k));
        n = 0; acc = 0; wi = 0;
    }
};
struct ByteCounter {
    uint64_t n;
    AFC_HDM void init() { n = 0; }
    AFC_HDM void put(uint32_t) { ++n; }
    AFC_HDM void put4(uint32_t) { n += 4; }
};

// 1 iff each of the four bytes of x is written unchanged by Go's encoder: 0x20..0x7f and none of  " \\ < > &
// (SWAR: "some byte is zero" = (v - 0x01010101) & ~v & 0x80808080;  < and > differ in bit 1,  " and & in bit 2)
AFC_HD uint32_t json_word_is_plain(uint32_t x) {
    const uint32_t H = 0x80808080u, L = 0x01010101u;
    uint32_t bad = x & H;                                            // >= 0x80
    bad |= (x - 0x20202020u) & ~x & H;                               // < 0x20
    uint32_t v = (x | 0x02020202u) ^ 0x3e3e3e3eu; bad |= (v - L) & ~v & H;      // < >
    v = (x | 0x04040404u) ^ 0x26262626u; bad |= (v - L) & ~v & H;               // " &
    v = x ^ 0x5c5c5c5cu; bad |= (v - L) & ~v & H;                               // backslash
    return bad == 0;
}

AFC_HD uint32_t json_hex_digit(uint32_t v) { return v < 10 ? '0' + v : 'a' + (v - 10); }

// utf8.DecodeRune on the window: returns the sequence length (1..4) if s[0..] starts a valid encoding, 0 if not.
// avail = bytes of the string available from the current position.
AFC_HD uint32_t utf8_valid_len(const ByteWindow& w, uint64_t avail) {
    const uint32_t b0 = w.peek(0);
    uint32_t need, lo = 0x80, hi = 0xBF;
    if (b0 >= 0xC2 && b0 <= 0xDF) need = 2;
    else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 3; if (b0 == 0xE0) lo = 0xA0; else if (b0 == 0xED) hi = 0x9F; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 4; if (b0 == 0xF0) lo = 0x90; else if (b0 == 0xF4) hi = 0x8F; }
    else return 0;
    if (avail < need) return 0;
    const uint32_t b1 = w.peek(1);
    if (b1 < lo || b1 > hi) return 0;
    for (uint32_t k = 2; k < need; k++) { uint32_t b = w.peek(k); if (b < 0x80 || b > 0xBF) return 0; }
    return need;
}

// Writes the Go-escaped form of s[0..len) (no surrounding quotes) to `out`.
template <class Sink>
AFC_HD void go_json_escape(Sink& out, const uint8_t* s, uint64_t len) {
    ByteWindow w; w.init(s, len);
    while (w.rem) {
        w.fill();
        if (w.rem >= 4 && w.have >= 4) {                        // four plain bytes at a time (DIDs, hashes, timestamps: the common case)
            const uint32_t x = (uint32_t)w.win;
            if (json_word_is_plain(x)) { out.put4(x); w.skip(4); continue; }
        }
        const uint32_t b = w.peek(0);
        if (b < 0x80) {
            if (b >= 0x20 && b != '"' && b != '\\' && b != '<' && b != '>' && b != '&') { out.put(b); w.skip(1); continue; }
            out.put('\\');
            uint32_t c = 0;
            switch (b) {
            case '"': case '\\': c = b; break;
            case '\b': c = 'b'; break;
            case '\f': c = 'f'; break;
            case '\n': c = 'n'; break;
            case '\r': c = 'r'; break;
            case '\t': c = 't'; break;
            default: break;
            }
            if (c) out.put(c);
            else { out.put('u'); out.put('0'); out.put('0'); out.put(json_hex_digit(b >> 4)); out.put(json_hex_digit(b & 15)); }
            w.skip(1);
            continue;
        }
        const uint32_t k = utf8_valid_len(w, w.rem);
        if (k == 0) {                                          // RuneError, width 1
            out.put('\\'); out.put('u'); out.put('f'); out.put('f'); out.put('f'); out.put('d');
            w.skip(1);
            continue;
        }
        if (k == 3 && b == 0xE2 && w.peek(1) == 0x80 && (w.peek(2) == 0xA8 || w.peek(2) == 0xA9)) {     // U+2028 / U+2029
            out.put('\\'); out.put('u'); out.put('2'); out.put('0'); out.put('2'); out.put(json_hex_digit(w.peek(2) & 15));
            w.skip(3);
            continue;
        }
        for (uint32_t j = 0; j < k; j++) out.put(w.peek(j));
        w.skip(k);
    }
}

template <class Sink>
AFC_HD void json_copy_raw(Sink& out, const uint8_t* s, uint64_t len) {
    ByteWindow w; w.init(s, len);
    while (w.rem) {
        w.fill();
        if (w.rem >= 4 && w.have >= 4) { out.put4((uint32_t)w.win); w.skip(4); continue; }
        uint32_t k = w.have < w.rem ? w.have : (uint32_t)w.rem;
        for (uint32_t j = 0; j < k; j++) out.put(w.peek(j));
        w.skip(k);
    }
}

enum { JSON_KIND_STRING = 0, JSON_KIND_RAW = 1 };

// One document: seg[0] v0 seg[1] v1 ... v(F-1) seg[F].  segs / seg_off: F+1 constant segments (seg_off has F+2 entries);
// fields / foff: this item's F values, foff[0..F] absolute offsets into `fields`.
template <class Sink>
AFC_HD void json_fill_one(Sink& out, const uint8_t* segs, const uint32_t* seg_off, const uint8_t* kinds, uint32_t F,
                          const uint8_t* fields, const uint64_t* foff) {
    for (uint32_t f = 0; f <= F; f++) {
        json_copy_raw(out, segs + seg_off[f], seg_off[f + 1] - seg_off[f]);
        if (f == F) break;
        const uint8_t* v = fields + foff[f];
        const uint64_t len = foff[f + 1] - foff[f];
        if (kinds[f] == JSON_KIND_RAW) json_copy_raw(out, v, len);
        else go_json_escape(out, v, len);
    }
}

}  // namespace afc
