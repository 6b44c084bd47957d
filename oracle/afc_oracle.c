/* ORACLE — TEST INFRASTRUCTURE ONLY (see afc_oracle.h).  Plain C, gcc, CPU.
 *
 * Restates what the reference computes through Go's stdlib on the hot path:
 *   ed25519.Verify  -> afo_ed25519_verify   (vc_service.go:504,1624; cli/vc_verification_enhanced.go:453)
 *   ed25519.Sign / NewKeyFromSeed -> afo_ed25519_sign / afo_ed25519_pubkey (vc_service.go:460-463, did_service.go:523)
 *   sha256.Sum256   -> afo_sha256           (vc_service.go:513, payload_store.go:69, did_service.go:517)
 *   hmac.New(sha256.New, k) -> afo_hmac_sha256 (webhook_dispatcher.go:470-474)
 * Structure follows Go's crypto/internal/fips140/edwards25519 (51-bit limbs; NAF-5 / NAF-8 vartime
 * double-scalar multiplication; radix-16 fixed-base tables), written from the published algorithms.
 */
#include "afc_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "afc_consts.inc"

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ SHA-256 */
static inline uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

static void sha256_block(uint32_t st[8], const uint8_t *p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ror32(w[i - 15], 7) ^ ror32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ror32(w[i - 2], 17) ^ ror32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25)) + ((e & f) ^ (~e & g)) + AFC_K256[i] + w[i];
        uint32_t t2 = (ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

typedef struct { uint32_t st[8]; uint8_t buf[64]; uint64_t len; } sha256_ctx;
static void sha256_init(sha256_ctx *c) { memcpy(c->st, AFC_H256, 32); c->len = 0; }
static void sha256_update(sha256_ctx *c, const uint8_t *p, size_t n) {
    size_t fill = c->len & 63;
    c->len += n;
    if (fill) {
        size_t take = 64 - fill < n ? 64 - fill : n;
        memcpy(c->buf + fill, p, take);
        p += take; n -= take;
        if (fill + take < 64) return;
        sha256_block(c->st, c->buf);
    }
    for (; n >= 64; p += 64, n -= 64) sha256_block(c->st, p);
    if (n) memcpy(c->buf, p, n);
}
static void sha256_final(sha256_ctx *c, uint8_t out[32]) {
    uint64_t bits = c->len * 8;
    size_t fill = c->len & 63;
    c->buf[fill++] = 0x80;
    if (fill > 56) { memset(c->buf + fill, 0, 64 - fill); sha256_block(c->st, c->buf); fill = 0; }
    memset(c->buf + fill, 0, 56 - fill);
    for (int i = 0; i < 8; i++) c->buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
    sha256_block(c->st, c->buf);
    for (int i = 0; i < 8; i++) { out[4*i] = c->st[i] >> 24; out[4*i+1] = c->st[i] >> 16; out[4*i+2] = c->st[i] >> 8; out[4*i+3] = c->st[i]; }
}
void afo_sha256(const uint8_t *msg, size_t len, uint8_t out[32]) {
    sha256_ctx c; sha256_init(&c); sha256_update(&c, msg, len); sha256_final(&c, out);
}

/* ------------------------------------------------------------------ SHA-512 */
static void sha512_block(uint64_t st[8], const uint8_t *p) {
    uint64_t w[80];
    for (int i = 0; i < 16; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | p[8 * i + j];
        w[i] = v;
    }
    for (int i = 16; i < 80; i++) {
        uint64_t s0 = ror64(w[i - 15], 1) ^ ror64(w[i - 15], 8) ^ (w[i - 15] >> 7);
        uint64_t s1 = ror64(w[i - 2], 19) ^ ror64(w[i - 2], 61) ^ (w[i - 2] >> 6);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 80; i++) {
        uint64_t t1 = h + (ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41)) + ((e & f) ^ (~e & g)) + AFC_K512[i] + w[i];
        uint64_t t2 = (ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
typedef struct { uint64_t st[8]; uint8_t buf[128]; uint64_t len; } sha512_ctx;
static void sha512_init(sha512_ctx *c) { memcpy(c->st, AFC_H512, 64); c->len = 0; }
static void sha512_update(sha512_ctx *c, const uint8_t *p, size_t n) {
    size_t fill = c->len & 127;
    c->len += n;
    if (fill) {
        size_t take = 128 - fill < n ? 128 - fill : n;
        memcpy(c->buf + fill, p, take);
        p += take; n -= take;
        if (fill + take < 128) return;
        sha512_block(c->st, c->buf);
    }
    for (; n >= 128; p += 128, n -= 128) sha512_block(c->st, p);
    if (n) memcpy(c->buf, p, n);
}
static void sha512_final(sha512_ctx *c, uint8_t out[64]) {
    uint64_t bits = c->len * 8;
    size_t fill = c->len & 127;
    c->buf[fill++] = 0x80;
    if (fill > 112) { memset(c->buf + fill, 0, 128 - fill); sha512_block(c->st, c->buf); fill = 0; }
    memset(c->buf + fill, 0, 120 - fill);
    for (int i = 0; i < 8; i++) c->buf[120 + i] = (uint8_t)(bits >> (56 - 8 * i));
    sha512_block(c->st, c->buf);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c->st[i] >> (56 - 8 * j));
}
void afo_sha512(const uint8_t *msg, size_t len, uint8_t out[64]) {
    sha512_ctx c; sha512_init(&c); sha512_update(&c, msg, len); sha512_final(&c, out);
}

/* ------------------------------------------------------------------ HMAC-SHA256 (RFC 2104) */
void afo_hmac_sha256(const uint8_t *key, size_t klen, const uint8_t *msg, size_t mlen, uint8_t out[32]) {
    uint8_t k0[64] = {0}, pad[64], inner[32];
    if (klen > 64) afo_sha256(key, klen, k0); else memcpy(k0, key, klen);
    sha256_ctx c;
    for (int i = 0; i < 64; i++) pad[i] = k0[i] ^ 0x36;
    sha256_init(&c); sha256_update(&c, pad, 64); sha256_update(&c, msg, mlen); sha256_final(&c, inner);
    for (int i = 0; i < 64; i++) pad[i] = k0[i] ^ 0x5c;
    sha256_init(&c); sha256_update(&c, pad, 64); sha256_update(&c, inner, 32); sha256_final(&c, out);
}

/* ------------------------------------------------------------------ field GF(2^255-19), 5x51 */
typedef struct { uint64_t v[5]; } fe;
#define M51 ((1ULL << 51) - 1)

static void fe_0(fe *h) { memset(h, 0, sizeof *h); }
static void fe_1(fe *h) { fe_0(h); h->v[0] = 1; }
static void fe_carry(fe *h) {
    uint64_t c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= M51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= M51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= M51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= M51; h->v[0] += c * 19;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static void fe_add(fe *h, const fe *f, const fe *g) {
    for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
    fe_carry(h);
}
static void fe_sub(fe *h, const fe *f, const fe *g) {
    /* f + 4p - g, limbs of 4p */
    h->v[0] = f->v[0] + 0x1FFFFFFFFFFFB4ULL - g->v[0];
    for (int i = 1; i < 5; i++) h->v[i] = f->v[i] + 0x1FFFFFFFFFFFFCULL - g->v[i];
    fe_carry(h);
}
static void fe_neg(fe *h, const fe *f) { fe z; fe_0(&z); fe_sub(h, &z, f); }
static void fe_mul(fe *h, const fe *f, const fe *g) {
    const uint64_t *a = f->v, *b = g->v;
    uint64_t b1 = b[1] * 19, b2 = b[2] * 19, b3 = b[3] * 19, b4 = b[4] * 19;
    u128 r0 = (u128)a[0] * b[0] + (u128)a[1] * b4 + (u128)a[2] * b3 + (u128)a[3] * b2 + (u128)a[4] * b1;
    u128 r1 = (u128)a[0] * b[1] + (u128)a[1] * b[0] + (u128)a[2] * b4 + (u128)a[3] * b3 + (u128)a[4] * b2;
    u128 r2 = (u128)a[0] * b[2] + (u128)a[1] * b[1] + (u128)a[2] * b[0] + (u128)a[3] * b4 + (u128)a[4] * b3;
    u128 r3 = (u128)a[0] * b[3] + (u128)a[1] * b[2] + (u128)a[2] * b[1] + (u128)a[3] * b[0] + (u128)a[4] * b4;
    u128 r4 = (u128)a[0] * b[4] + (u128)a[1] * b[3] + (u128)a[2] * b[2] + (u128)a[3] * b[1] + (u128)a[4] * b[0];
    uint64_t c;
    r1 += (uint64_t)(r0 >> 51); h->v[0] = (uint64_t)r0 & M51;
    r2 += (uint64_t)(r1 >> 51); h->v[1] = (uint64_t)r1 & M51;
    r3 += (uint64_t)(r2 >> 51); h->v[2] = (uint64_t)r2 & M51;
    r4 += (uint64_t)(r3 >> 51); h->v[3] = (uint64_t)r3 & M51;
    c = (uint64_t)(r4 >> 51);   h->v[4] = (uint64_t)r4 & M51;
    h->v[0] += c * 19;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static void fe_sq(fe *h, const fe *f) { fe_mul(h, f, f); }
static void fe_sqn(fe *h, const fe *f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }
static void fe_frombytes(fe *h, const uint8_t s[32]) {
    uint64_t w[4];
    for (int i = 0; i < 4; i++) { w[i] = 0; for (int j = 7; j >= 0; j--) w[i] = (w[i] << 8) | s[8 * i + j]; }
    h->v[0] = w[0] & M51;
    h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    h->v[4] = (w[3] >> 12) & M51;               /* bit 255 ignored; y >= p accepted (Go SetBytes) */
}
static void fe_tobytes(uint8_t s[32], const fe *f) {
    fe h = *f;
    fe_carry(&h); fe_carry(&h);
    uint64_t q = (h.v[0] + 19) >> 51;
    q = (h.v[1] + q) >> 51; q = (h.v[2] + q) >> 51; q = (h.v[3] + q) >> 51; q = (h.v[4] + q) >> 51;
    h.v[0] += 19 * q;
    uint64_t c;
    c = h.v[0] >> 51; h.v[0] &= M51; h.v[1] += c;
    c = h.v[1] >> 51; h.v[1] &= M51; h.v[2] += c;
    c = h.v[2] >> 51; h.v[2] &= M51; h.v[3] += c;
    c = h.v[3] >> 51; h.v[3] &= M51; h.v[4] += c;
    h.v[4] &= M51;
    uint64_t w[4];
    w[0] = h.v[0] | (h.v[1] << 51);
    w[1] = (h.v[1] >> 13) | (h.v[2] << 38);
    w[2] = (h.v[2] >> 26) | (h.v[3] << 25);
    w[3] = (h.v[3] >> 39) | (h.v[4] << 12);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}
static int fe_isneg(const fe *f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
static int fe_eq(const fe *a, const fe *b) { uint8_t x[32], y[32]; fe_tobytes(x, a); fe_tobytes(y, b); return memcmp(x, y, 32) == 0; }
/* z^(2^250-1) chain shared by invert and pow22523; also returns z^11 */
static void fe_pow_2_250_m1(fe *out, fe *z11, const fe *z) {
    fe t0, t1, t2, t3;
    fe_sq(&t0, z);                         /* 2 */
    fe_sqn(&t1, &t0, 2);                   /* 8 */
    fe_mul(&t1, z, &t1);                   /* 9 */
    fe_mul(&t0, &t0, &t1);                 /* 11 */
    *z11 = t0;
    fe_sq(&t2, &t0);                       /* 22 */
    fe_mul(&t1, &t1, &t2);                 /* 31 = 2^5-1 */
    fe_sqn(&t2, &t1, 5);  fe_mul(&t1, &t2, &t1);    /* 2^10-1 */
    fe_sqn(&t2, &t1, 10); fe_mul(&t2, &t2, &t1);    /* 2^20-1 */
    fe_sqn(&t3, &t2, 20); fe_mul(&t2, &t3, &t2);    /* 2^40-1 */
    fe_sqn(&t2, &t2, 10); fe_mul(&t1, &t2, &t1);    /* 2^50-1 */
    fe_sqn(&t2, &t1, 50); fe_mul(&t2, &t2, &t1);    /* 2^100-1 */
    fe_sqn(&t3, &t2, 100); fe_mul(&t2, &t3, &t2);   /* 2^200-1 */
    fe_sqn(&t2, &t2, 50); fe_mul(out, &t2, &t1);    /* 2^250-1 */
}
static void fe_invert(fe *out, const fe *z) {
    fe t, z11; fe_pow_2_250_m1(&t, &z11, z);
    fe_sqn(&t, &t, 5); fe_mul(out, &t, &z11);       /* 2^255-21 */
}
static void fe_pow22523(fe *out, const fe *z) {
    fe t, z11; fe_pow_2_250_m1(&t, &z11, z);
    fe_sqn(&t, &t, 2); fe_mul(out, &t, z);          /* 2^252-3 */
}
static fe FE_D, FE_D2, FE_SQRTM1;

/* ------------------------------------------------------------------ group */
typedef struct { fe X, Y, Z, T; } ge_p3;
typedef struct { fe X, Y, Z; } ge_p2;
typedef struct { fe X, Y, Z, T; } ge_p1p1;
typedef struct { fe YpX, YmX, Z, T2d; } ge_cached;
typedef struct { fe ypx, ymx, xy2d; } ge_precomp;

static void ge_p3_0(ge_p3 *h) { fe_0(&h->X); fe_1(&h->Y); fe_1(&h->Z); fe_0(&h->T); }
static void ge_p2_0(ge_p2 *h) { fe_0(&h->X); fe_1(&h->Y); fe_1(&h->Z); }
static void ge_p1p1_to_p2(ge_p2 *r, const ge_p1p1 *p) { fe_mul(&r->X, &p->X, &p->T); fe_mul(&r->Y, &p->Y, &p->Z); fe_mul(&r->Z, &p->Z, &p->T); }
static void ge_p1p1_to_p3(ge_p3 *r, const ge_p1p1 *p) { fe_mul(&r->X, &p->X, &p->T); fe_mul(&r->Y, &p->Y, &p->Z); fe_mul(&r->Z, &p->Z, &p->T); fe_mul(&r->T, &p->X, &p->Y); }
static void ge_p3_to_p2(ge_p2 *r, const ge_p3 *p) { r->X = p->X; r->Y = p->Y; r->Z = p->Z; }
static void ge_p3_to_cached(ge_cached *r, const ge_p3 *p) { fe_add(&r->YpX, &p->Y, &p->X); fe_sub(&r->YmX, &p->Y, &p->X); r->Z = p->Z; fe_mul(&r->T2d, &p->T, &FE_D2); }
static void ge_p2_dbl(ge_p1p1 *r, const ge_p2 *p) {
    fe xx, yy, b, a;
    fe_sq(&xx, &p->X); fe_sq(&yy, &p->Y); fe_sq(&b, &p->Z); fe_add(&b, &b, &b);
    fe_add(&a, &p->X, &p->Y); fe_sq(&a, &a);
    fe_add(&r->Y, &yy, &xx); fe_sub(&r->Z, &yy, &xx); fe_sub(&r->X, &a, &r->Y); fe_sub(&r->T, &b, &r->Z);
}
static void ge_p3_dbl(ge_p1p1 *r, const ge_p3 *p) { ge_p2 q; ge_p3_to_p2(&q, p); ge_p2_dbl(r, &q); }
static void ge_add(ge_p1p1 *r, const ge_p3 *p, const ge_cached *q) {
    fe a, b, c, d;
    fe_add(&a, &p->Y, &p->X); fe_sub(&b, &p->Y, &p->X);
    fe_mul(&a, &a, &q->YpX); fe_mul(&b, &b, &q->YmX); fe_mul(&c, &q->T2d, &p->T); fe_mul(&d, &p->Z, &q->Z); fe_add(&d, &d, &d);
    fe_sub(&r->X, &a, &b); fe_add(&r->Y, &a, &b); fe_add(&r->Z, &d, &c); fe_sub(&r->T, &d, &c);
}
static void ge_sub(ge_p1p1 *r, const ge_p3 *p, const ge_cached *q) {
    fe a, b, c, d;
    fe_add(&a, &p->Y, &p->X); fe_sub(&b, &p->Y, &p->X);
    fe_mul(&a, &a, &q->YmX); fe_mul(&b, &b, &q->YpX); fe_mul(&c, &q->T2d, &p->T); fe_mul(&d, &p->Z, &q->Z); fe_add(&d, &d, &d);
    fe_sub(&r->X, &a, &b); fe_add(&r->Y, &a, &b); fe_sub(&r->Z, &d, &c); fe_add(&r->T, &d, &c);
}
static void ge_madd(ge_p1p1 *r, const ge_p3 *p, const ge_precomp *q) {
    fe a, b, c, d;
    fe_add(&a, &p->Y, &p->X); fe_sub(&b, &p->Y, &p->X);
    fe_mul(&a, &a, &q->ypx); fe_mul(&b, &b, &q->ymx); fe_mul(&c, &q->xy2d, &p->T); fe_add(&d, &p->Z, &p->Z);
    fe_sub(&r->X, &a, &b); fe_add(&r->Y, &a, &b); fe_add(&r->Z, &d, &c); fe_sub(&r->T, &d, &c);
}
static void ge_msub(ge_p1p1 *r, const ge_p3 *p, const ge_precomp *q) {
    fe a, b, c, d;
    fe_add(&a, &p->Y, &p->X); fe_sub(&b, &p->Y, &p->X);
    fe_mul(&a, &a, &q->ymx); fe_mul(&b, &b, &q->ypx); fe_mul(&c, &q->xy2d, &p->T); fe_add(&d, &p->Z, &p->Z);
    fe_sub(&r->X, &a, &b); fe_add(&r->Y, &a, &b); fe_sub(&r->Z, &d, &c); fe_add(&r->T, &d, &c);
}
static void ge_p3_to_precomp(ge_precomp *r, const ge_p3 *p) {
    fe zi, x, y, xy;
    fe_invert(&zi, &p->Z); fe_mul(&x, &p->X, &zi); fe_mul(&y, &p->Y, &zi);
    fe_add(&r->ypx, &y, &x); fe_sub(&r->ymx, &y, &x); fe_mul(&xy, &x, &y); fe_mul(&r->xy2d, &xy, &FE_D2);
}
static void ge_encode(uint8_t s[32], const fe *X, const fe *Y, const fe *Z) {
    fe zi, x, y;
    fe_invert(&zi, Z); fe_mul(&x, X, &zi); fe_mul(&y, Y, &zi);
    fe_tobytes(s, &y);
    s[31] ^= (uint8_t)(fe_isneg(&x) << 7);
}
/* Go Point.SetBytes: returns 0 on success, -1 if not on curve */
static int ge_frombytes(ge_p3 *h, const uint8_t s[32]) {
    fe u, v, v3, vxx, check, one;
    fe_1(&one);
    fe_frombytes(&h->Y, s);
    fe_1(&h->Z);
    fe_sq(&u, &h->Y); fe_mul(&v, &u, &FE_D);
    fe_sub(&u, &u, &one);                 /* u = y^2-1 */
    fe_add(&v, &v, &one);                 /* v = dy^2+1 */
    fe_sq(&v3, &v); fe_mul(&v3, &v3, &v); /* v^3 */
    fe_sq(&h->X, &v3); fe_mul(&h->X, &h->X, &v); fe_mul(&h->X, &h->X, &u); /* u v^7 */
    fe_pow22523(&h->X, &h->X);
    fe_mul(&h->X, &h->X, &v3); fe_mul(&h->X, &h->X, &u);                  /* u v^3 (u v^7)^((p-5)/8) */
    fe_sq(&vxx, &h->X); fe_mul(&vxx, &vxx, &v);
    fe_sub(&check, &vxx, &u);
    uint8_t z[32]; static const uint8_t zero[32] = {0};
    fe_tobytes(z, &check);
    if (memcmp(z, zero, 32) != 0) {
        fe_add(&check, &vxx, &u);
        fe_tobytes(z, &check);
        if (memcmp(z, zero, 32) != 0) return -1;
        fe_mul(&h->X, &h->X, &FE_SQRTM1);
    }
    if (fe_isneg(&h->X) != (s[31] >> 7)) fe_neg(&h->X, &h->X);   /* x==0 with sign bit: stays 0, accepted */
    fe_mul(&h->T, &h->X, &h->Y);
    return 0;
}

/* ------------------------------------------------------------------ scalars mod L (Barrett, 32-bit limbs) */
static int sc_geq(const uint32_t *a, const uint32_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 1;
}
static void sc_reduce512(uint8_t out[32], const uint32_t x[16]) {
    uint32_t q2[18] = {0}, r2[9] = {0}, r[9], l9[9];
    const uint32_t *q1 = x + 7;                       /* 9 limbs */
    for (int i = 0; i < 9; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 9; j++) { c += (uint64_t)q1[i] * AFC_MU_32[j] + q2[i + j]; q2[i + j] = (uint32_t)c; c >>= 32; }
        q2[i + 9] = (uint32_t)c;
    }
    const uint32_t *q3 = q2 + 9;                      /* 9 limbs */
    for (int i = 0; i < 9; i++) {
        uint64_t c = 0;
        int j = 0;
        for (; j < 8 && i + j < 9; j++) { c += (uint64_t)q3[i] * AFC_L_32[j] + r2[i + j]; r2[i + j] = (uint32_t)c; c >>= 32; }
        if (i + j < 9) r2[i + j] += (uint32_t)c;
    }
    int64_t bw = 0;
    for (int i = 0; i < 9; i++) { int64_t d = (int64_t)x[i] - r2[i] + bw; r[i] = (uint32_t)d; bw = d >> 32; }
    memcpy(l9, AFC_L_32, 32); l9[8] = 0;
    for (int k = 0; k < 3 && sc_geq(r, l9, 9); k++) {
        bw = 0;
        for (int i = 0; i < 9; i++) { int64_t d = (int64_t)r[i] - l9[i] + bw; r[i] = (uint32_t)d; bw = d >> 32; }
    }
    for (int i = 0; i < 8; i++) { out[4*i] = r[i]; out[4*i+1] = r[i] >> 8; out[4*i+2] = r[i] >> 16; out[4*i+3] = r[i] >> 24; }
}
static void load32le(uint32_t *w, const uint8_t *b, int n) {
    for (int i = 0; i < n; i++) w[i] = (uint32_t)b[4*i] | ((uint32_t)b[4*i+1] << 8) | ((uint32_t)b[4*i+2] << 16) | ((uint32_t)b[4*i+3] << 24);
}
static void sc_reduce_bytes64(uint8_t out[32], const uint8_t in[64]) { uint32_t x[16]; load32le(x, in, 16); sc_reduce512(out, x); }
/* out = (a*b + c) mod L; a, b, c 32-byte little-endian */
static void sc_muladd(uint8_t out[32], const uint8_t a[32], const uint8_t b[32], const uint8_t c[32]) {
    uint32_t A[8], B[8], C[8], x[16] = {0};
    load32le(A, a, 8); load32le(B, b, 8); load32le(C, c, 8);
    for (int i = 0; i < 8; i++) {
        uint64_t cy = 0;
        for (int j = 0; j < 8; j++) { cy += (uint64_t)A[i] * B[j] + x[i + j]; x[i + j] = (uint32_t)cy; cy >>= 32; }
        x[i + 8] = (uint32_t)cy;
    }
    uint64_t cy = 0;
    for (int i = 0; i < 16; i++) { cy += (uint64_t)x[i] + (i < 8 ? C[i] : 0); x[i] = (uint32_t)cy; cy >>= 32; }
    sc_reduce512(out, x);
}
static int sc_is_canonical(const uint8_t s[32]) { uint32_t w[8]; load32le(w, s, 8); return !sc_geq(w, AFC_L_32, 8); }

/* width-w non-adjacent form of a 256-bit little-endian scalar (< 2^253) */
static void sc_naf(int8_t naf[257], const uint8_t s[32], int w) {
    uint64_t k[5] = {0};
    for (int i = 0; i < 4; i++) for (int j = 7; j >= 0; j--) k[i] = (k[i] << 8) | s[8 * i + j];
    memset(naf, 0, 257);
    int width = 1 << w, half = 1 << (w - 1);
    for (int pos = 0; pos < 257; pos++) {
        if (k[0] & 1) {
            int d = (int)(k[0] & (uint64_t)(width - 1));
            if (d >= half) {                 /* negative digit: add (width-d) to k */
                d -= width;
                u128 c = (u128)k[0] + (uint64_t)(-d);
                k[0] = (uint64_t)c; c >>= 64;
                for (int i = 1; i < 5 && c; i++) { c += k[i]; k[i] = (uint64_t)c; c >>= 64; }
            } else {
                k[0] -= (uint64_t)d;
            }
            naf[pos] = (int8_t)d;
        }
        k[0] = (k[0] >> 1) | (k[1] << 63); k[1] = (k[1] >> 1) | (k[2] << 63);
        k[2] = (k[2] >> 1) | (k[3] << 63); k[3] = (k[3] >> 1) | (k[4] << 63); k[4] >>= 1;
    }
}

/* ------------------------------------------------------------------ tables (built once) */
static ge_precomp B_NAF8[64];        /* (2i+1)B affine, for verify */
static ge_precomp B_RADIX16[32][8];  /* (j+1) * 256^i * B affine, for fixed-base mult */
static ge_p3 GE_B;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static void fe_from51(fe *h, const uint64_t l[5]) { memcpy(h->v, l, 40); }
static void init_tables(void) {
    fe_from51(&FE_D, AFC_D_51); fe_from51(&FE_D2, AFC_D2_51); fe_from51(&FE_SQRTM1, AFC_SQRTM1_51);
    fe_from51(&GE_B.X, AFC_BX_51); fe_from51(&GE_B.Y, AFC_BY_51); fe_1(&GE_B.Z); fe_mul(&GE_B.T, &GE_B.X, &GE_B.Y);
    /* odd multiples */
    ge_p1p1 t; ge_p3 b2, cur = GE_B; ge_cached c2;
    ge_p3_dbl(&t, &GE_B); ge_p1p1_to_p3(&b2, &t); ge_p3_to_cached(&c2, &b2);
    for (int i = 0; i < 64; i++) {
        ge_p3_to_precomp(&B_NAF8[i], &cur);
        ge_add(&t, &cur, &c2); ge_p1p1_to_p3(&cur, &t);
    }
    /* radix-16 tables at 256^i */
    ge_p3 base = GE_B;
    for (int i = 0; i < 32; i++) {
        ge_cached cb; ge_p3_to_cached(&cb, &base);
        ge_p3 m = base;
        for (int j = 0; j < 8; j++) {
            ge_p3_to_precomp(&B_RADIX16[i][j], &m);
            ge_add(&t, &m, &cb); ge_p1p1_to_p3(&m, &t);
        }
        for (int k = 0; k < 8; k++) { ge_p3_dbl(&t, &base); ge_p1p1_to_p3(&base, &t); }
    }
}

/* h = a*B, a 32-byte LE scalar < 2^253 (already reduced or clamped-then-reduced) */
static void ge_scalarmult_base(ge_p3 *h, const uint8_t a[32]) {
    int8_t e[64];
    for (int i = 0; i < 32; i++) { e[2 * i] = a[i] & 15; e[2 * i + 1] = (a[i] >> 4) & 15; }
    int carry = 0;
    for (int i = 0; i < 63; i++) { e[i] += carry; carry = (e[i] + 8) >> 4; e[i] -= carry << 4; }
    e[63] += carry;
    ge_p1p1 r; ge_p2 s;
    ge_p3_0(h);
    for (int pass = 1; pass >= 0; pass--) {
        for (int i = pass; i < 64; i += 2) {
            int d = e[i];
            if (d > 0) { ge_madd(&r, h, &B_RADIX16[i / 2][d - 1]); ge_p1p1_to_p3(h, &r); }
            else if (d < 0) { ge_msub(&r, h, &B_RADIX16[i / 2][-d - 1]); ge_p1p1_to_p3(h, &r); }
        }
        if (pass == 1) {
            ge_p3_dbl(&r, h); ge_p1p1_to_p2(&s, &r);
            ge_p2_dbl(&r, &s); ge_p1p1_to_p2(&s, &r);
            ge_p2_dbl(&r, &s); ge_p1p1_to_p2(&s, &r);
            ge_p2_dbl(&r, &s); ge_p1p1_to_p3(h, &r);
        }
    }
}

static void expand_seed(const uint8_t seed[32], uint8_t s_clamped[32], uint8_t prefix[32]) {
    uint8_t h[64];
    afo_sha512(seed, 32, h);
    h[0] &= 248; h[31] &= 63; h[31] |= 64;
    memcpy(s_clamped, h, 32); memcpy(prefix, h + 32, 32);
}
static void scalar_base_encode(uint8_t out[32], const uint8_t scalar_any[32]) {
    /* reduce mod L first (Go: SetBytesWithClamping reduces) so the radix-16 recoding sees < 2^253 */
    uint8_t wide[64] = {0}, red[32];
    memcpy(wide, scalar_any, 32);
    sc_reduce_bytes64(red, wide);
    ge_p3 p; ge_scalarmult_base(&p, red);
    ge_encode(out, &p.X, &p.Y, &p.Z);
}
void afo_ed25519_pubkey(const uint8_t seed[32], uint8_t pk[32]) {
    pthread_once(&g_once, init_tables);
    uint8_t s[32], prefix[32];
    expand_seed(seed, s, prefix);
    scalar_base_encode(pk, s);
}
void afo_ed25519_sign(const uint8_t seed[32], const uint8_t *msg, size_t len, uint8_t sig[64]) {
    pthread_once(&g_once, init_tables);
    uint8_t s[32], prefix[32], pk[32], d[64], r[32], k[32];
    expand_seed(seed, s, prefix);
    scalar_base_encode(pk, s);
    sha512_ctx c;
    sha512_init(&c); sha512_update(&c, prefix, 32); sha512_update(&c, msg, len); sha512_final(&c, d);
    sc_reduce_bytes64(r, d);
    scalar_base_encode(sig, r);
    sha512_init(&c); sha512_update(&c, sig, 32); sha512_update(&c, pk, 32); sha512_update(&c, msg, len); sha512_final(&c, d);
    sc_reduce_bytes64(k, d);
    sc_muladd(sig + 32, k, s, r);
}
int afo_ed25519_verify(const uint8_t pk[32], const uint8_t *msg, size_t len, const uint8_t sig[64]) {
    pthread_once(&g_once, init_tables);
    if (sig[63] & 0xE0) return 0;
    ge_p3 A;
    if (ge_frombytes(&A, pk) != 0) return 0;
    if (!sc_is_canonical(sig + 32)) return 0;
    uint8_t d[64], k[32];
    sha512_ctx c;
    sha512_init(&c); sha512_update(&c, sig, 32); sha512_update(&c, pk, 32); sha512_update(&c, msg, len); sha512_final(&c, d);
    sc_reduce_bytes64(k, d);
    fe_neg(&A.X, &A.X); fe_neg(&A.T, &A.T);           /* -A */
    /* odd multiples 1,3,..,15 of -A */
    ge_cached atab[8]; ge_p1p1 t; ge_p3 a2, u;
    ge_p3_to_cached(&atab[0], &A);
    ge_p3_dbl(&t, &A); ge_p1p1_to_p3(&a2, &t);
    for (int i = 0; i < 7; i++) { ge_add(&t, &a2, &atab[i]); ge_p1p1_to_p3(&u, &t); ge_p3_to_cached(&atab[i + 1], &u); }
    int8_t an[257], bn[257];
    sc_naf(an, k, 5); sc_naf(bn, sig + 32, 8);
    int i = 256;
    while (i >= 0 && !an[i] && !bn[i]) i--;
    ge_p2 r; ge_p2_0(&r);
    for (; i >= 0; i--) {
        ge_p2_dbl(&t, &r);
        if (an[i] > 0) { ge_p1p1_to_p3(&u, &t); ge_add(&t, &u, &atab[an[i] / 2]); }
        else if (an[i] < 0) { ge_p1p1_to_p3(&u, &t); ge_sub(&t, &u, &atab[(-an[i]) / 2]); }
        if (bn[i] > 0) { ge_p1p1_to_p3(&u, &t); ge_madd(&t, &u, &B_NAF8[bn[i] / 2]); }
        else if (bn[i] < 0) { ge_p1p1_to_p3(&u, &t); ge_msub(&t, &u, &B_NAF8[(-bn[i]) / 2]); }
        ge_p1p1_to_p2(&r, &t);
    }
    uint8_t enc[32];
    ge_encode(enc, &r.X, &r.Y, &r.Z);
    return memcmp(enc, sig, 32) == 0;
}

/* ------------------------------------------------------------------ RFC 6962 */
void afo_merkle_leaf_hash(const uint8_t *leaf, size_t len, uint8_t out[32]) {
    sha256_ctx c; uint8_t z = 0;
    sha256_init(&c); sha256_update(&c, &z, 1); sha256_update(&c, leaf, len); sha256_final(&c, out);
}
static void node_hash(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]) {
    uint8_t b[65]; b[0] = 1; memcpy(b + 1, l, 32); memcpy(b + 33, r, 32); afo_sha256(b, 65, out);
}
void afo_merkle_root_from_hashes(const uint8_t *hashes, uint64_t n, uint8_t root[32]) {
    if (n == 0) { afo_sha256((const uint8_t *)"", 0, root); return; }
    uint8_t stack[64][32]; int height[64]; int sp = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t cur[32]; int h = 0;
        memcpy(cur, hashes + 32 * i, 32);
        while (sp && height[sp - 1] == h) { node_hash(stack[sp - 1], cur, cur); sp--; h++; }
        memcpy(stack[sp], cur, 32); height[sp++] = h;
    }
    uint8_t acc[32]; memcpy(acc, stack[--sp], 32);
    while (sp) { sp--; node_hash(stack[sp], acc, acc); }
    memcpy(root, acc, 32);
}

/* ------------------------------------------------------------------ batch + threads */
typedef struct {
    int kind; uint32_t lo, hi;
    const uint8_t *a, *b, *msgs; const uint64_t *off; const uint32_t *koff; uint8_t *out;
} job_t;
static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (uint32_t i = j->lo; i < j->hi; i++) {
        const uint8_t *m = j->msgs ? j->msgs + j->off[i] : NULL;
        size_t len = j->msgs ? (size_t)(j->off[i + 1] - j->off[i]) : 0;
        switch (j->kind) {
        case 0: afo_sha256(m, len, j->out + 32 * (size_t)i); break;
        case 1: afo_hmac_sha256(j->a + j->koff[i], j->koff[i + 1] - j->koff[i], m, len, j->out + 32 * (size_t)i); break;
        case 2: j->out[i] = (uint8_t)afo_ed25519_verify(j->a + 32 * (size_t)i, m, len, j->b + 64 * (size_t)i); break;
        case 3: afo_ed25519_sign(j->a + 32 * (size_t)i, m, len, j->out + 64 * (size_t)i); break;
        case 4: afo_ed25519_pubkey(j->a + 32 * (size_t)i, j->out + 32 * (size_t)i); break;
        case 5: afo_merkle_leaf_hash(m, len, j->out + 32 * (size_t)i); break;
        }
    }
    return NULL;
}
static void run(job_t base, uint32_t n, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = base;
        jobs[t].lo = (uint32_t)((uint64_t)n * t / nthreads);
        jobs[t].hi = (uint32_t)((uint64_t)n * (t + 1) / nthreads);
        if (t) pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
}
void afo_sha256_batch(const uint8_t *msgs, const uint64_t *off, uint32_t n, uint8_t *out32, int nthreads) {
    job_t j = {0}; j.kind = 0; j.msgs = msgs; j.off = off; j.out = out32; run(j, n, nthreads);
}
void afo_hmac_sha256_batch(const uint8_t *keys, const uint32_t *koff, const uint8_t *msgs, const uint64_t *off,
                           uint32_t n, uint8_t *out32, int nthreads) {
    job_t j = {0}; j.kind = 1; j.a = keys; j.koff = koff; j.msgs = msgs; j.off = off; j.out = out32; run(j, n, nthreads);
}
void afo_ed25519_verify_batch(const uint8_t *pks, const uint8_t *sigs, const uint8_t *msgs, const uint64_t *off,
                              uint32_t n, uint8_t *ok, int nthreads) {
    pthread_once(&g_once, init_tables);
    job_t j = {0}; j.kind = 2; j.a = pks; j.b = sigs; j.msgs = msgs; j.off = off; j.out = ok; run(j, n, nthreads);
}
void afo_ed25519_sign_batch(const uint8_t *seeds, const uint8_t *msgs, const uint64_t *off, uint32_t n,
                            uint8_t *sigs, int nthreads) {
    pthread_once(&g_once, init_tables);
    job_t j = {0}; j.kind = 3; j.a = seeds; j.msgs = msgs; j.off = off; j.out = sigs; run(j, n, nthreads);
}
void afo_ed25519_pubkey_batch(const uint8_t *seeds, uint32_t n, uint8_t *pks, int nthreads) {
    pthread_once(&g_once, init_tables);
    job_t j = {0}; j.kind = 4; j.a = seeds; j.out = pks; run(j, n, nthreads);
}
void afo_merkle_root(const uint8_t *leaves, const uint64_t *off, uint32_t n, uint8_t root[32], int nthreads) {
    uint8_t *h = (uint8_t *)malloc(32 * (size_t)(n ? n : 1));
    job_t j = {0}; j.kind = 5; j.msgs = leaves; j.off = off; j.out = h; run(j, n, nthreads);
    afo_merkle_root_from_hashes(h, n, root);
    free(h);
}
