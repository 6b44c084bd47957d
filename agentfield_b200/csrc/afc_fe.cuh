// GF(2^255-19) arithmetic for the Ed25519 kernels — 8 x 32-bit saturated limbs, one element per thread.
//
// Replaces the field layer under Go's crypto/ed25519 (crypto/internal/fips140/edwards25519/field, 5x51-bit
// limbs on amd64) that the reference reaches from vc_service.go:463,504.  B200 has no 64x64 multiplier: a
// 51-bit limb product costs four 32-bit IMADs, so the native radix here is 2^32 — 64 IMAD.WIDE.U32 per
// multiplication issued as carry chains (mad.lo.cc / madc.hi.cc) over separate even- and odd-column
// accumulators, so no carry ever crosses between two chains and each chain is a straight run of
// dependent-by-carry-flag-only instructions.  (north_star suggests 51-bit limbs spread over lanes with
// warp shuffles; thread-per-credential with 32-bit limbs keeps all 64 partial products in one thread's
// registers and needs no shuffles — see DESIGN.md "deviations".)
//
// Representation: any value in [0, 2^256) stands for its residue mod p ("weakly reduced").  2^256 = 38
// (mod p).  fe_tobytes produces the unique canonical encoding.
#pragma once
#include "afc_common.cuh"

#ifndef AFC_FE_PTX
#define AFC_FE_PTX 1
#endif

namespace afc {

struct alignas(16) fe { uint32_t v[8]; };

AFC_HD void fe_0(fe& h) {
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = 0;
}
AFC_HD void fe_1(fe& h) { fe_0(h); h.v[0] = 1; }
AFC_HD void fe_copy(fe& h, const fe& f) {
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = f.v[i];
}
AFC_HD void fe_from_words(fe& h, const uint32_t* w) {
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = w[i];
}


// ---------------------------------------------------------------------------------- portable variants
// Plain-C versions of every primitive: the reference the PTX paths are checked against on the device
// (afc_selftest) and the code tests/hostsim runs on the CPU.
AFC_HD void fe_add_c(fe& h, const fe& f, const fe& g) {
    uint64_t c = 0;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)f.v[i] + g.v[i]; r[i] = (uint32_t)c; c >>= 32; }
    c *= 38;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
    r[0] += (uint32_t)c * 38u;
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}
AFC_HD void fe_sub_c(fe& h, const fe& f, const fe& g) {
    int64_t c = 0;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (int64_t)f.v[i] - g.v[i]; r[i] = (uint32_t)c; c >>= 32; }
    int64_t k = (c < 0) ? 38 : 0;
    c = -k;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
    if (c < 0) r[0] -= 38u;
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}
AFC_HD void fe_fold16_c(fe& h, const uint32_t* t) {
    uint64_t c = 0;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)t[i] + (uint64_t)t[8 + i] * 38u; r[i] = (uint32_t)c; c >>= 32; }
    c *= 38;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += r[i]; r[i] = (uint32_t)c; c >>= 32; }
    r[0] += (uint32_t)c * 38u;
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}
AFC_HD void fe_mul_wide_c(uint32_t* t, const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { c += (uint64_t)a[j] * b[i] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
        t[i + 8] = (uint32_t)c;
    }
}
// t = a^2: off-diagonal products once (28), doubled, plus the 8 diagonal squares
AFC_HD void fe_sq_wide_c(uint32_t* t, const uint32_t* a) {
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = i + 1; j < 8; j++) { c += (uint64_t)a[i] * a[j] + s[i + j]; s[i + j] = (uint32_t)c; c >>= 32; }
        s[i + 8] = (uint32_t)c;
    }
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] * a[i];
        uint32_t lo2 = (s[2 * i] << 1) | (i ? (s[2 * i - 1] >> 31) : 0u);
        uint32_t hi2 = (s[2 * i + 1] << 1) | (s[2 * i] >> 31);
        c += (uint64_t)lo2 + (uint32_t)d;
        t[2 * i] = (uint32_t)c; c >>= 32;
        c += (uint64_t)hi2 + (uint32_t)(d >> 32);
        t[2 * i + 1] = (uint32_t)c; c >>= 32;
    }
}
AFC_HD void fe_mul_c(fe& h, const fe& f, const fe& g) { uint32_t t[16]; fe_mul_wide_c(t, f.v, g.v); fe_fold16_c(h, t); }
AFC_HD void fe_sq_c(fe& h, const fe& f) { uint32_t t[16]; fe_sq_wide_c(t, f.v); fe_fold16_c(h, t); }

// ---------------------------------------------------------------------------------- add / sub
// h = f + g  (mod 2^256-38, weakly reduced)
AFC_HD void fe_add(fe& h, const fe& f, const fe& g) {
#if AFC_DEVICE_CODE && AFC_FE_PTX
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7, c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7), "=r"(c)
        : "r"(f.v[0]), "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[6]), "r"(f.v[7]),
          "r"(g.v[0]), "r"(g.v[1]), "r"(g.v[2]), "r"(g.v[3]), "r"(g.v[4]), "r"(g.v[5]), "r"(g.v[6]), "r"(g.v[7]));
    uint32_t k = c * 38u, c2;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(c2)
        : "r"(k));
    r0 += c2 * 38u;     // second wrap leaves a value < 38: cannot carry
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
    fe_add_c(h, f, g);
#endif
}

// h = f - g  (mod 2^256-38, weakly reduced)
AFC_HD void fe_sub(fe& h, const fe& f, const fe& g) {
#if AFC_DEVICE_CODE && AFC_FE_PTX
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7, b;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"            // 0 or 0xffffffff
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7), "=r"(b)
        : "r"(f.v[0]), "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[6]), "r"(f.v[7]),
          "r"(g.v[0]), "r"(g.v[1]), "r"(g.v[2]), "r"(g.v[3]), "r"(g.v[4]), "r"(g.v[5]), "r"(g.v[6]), "r"(g.v[7]));
    uint32_t k = b & 38u, b2;
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(b2)
        : "r"(k));
    r0 -= b2 & 38u;     // second wrap leaves a value >= 2^256-38: cannot borrow
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
    fe_sub_c(h, f, g);
#endif
}

AFC_HD void fe_neg(fe& h, const fe& f) { fe z; fe_0(z); fe_sub(h, z, f); }
AFC_HD void fe_dbl(fe& h, const fe& f) { fe_add(h, f, f); }

// h = f + (g & keep)  when sgn == 0,  h = f - (g & keep)  when sgn == 0xffffffff  (keep: 0 or 0xffffffff), weakly reduced.
// ONE instruction stream for both signs: the four lanes that share a point in the quad kernels (k_ed25519.cu) each need a
// different one of B-A, B+A, D-C, D+C at the same moment.  f - g is computed as f + ~g + 1: a missing carry-out is a borrow, and
// the correction (2^256 = 38) changes sign with the operation.
AFC_HD void fe_addsub_m_c(fe& h, const fe& f, const fe& g, uint32_t keep, uint32_t sgn) {
    const uint32_t cin = sgn & 1u;
    uint64_t c = cin;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)f.v[i] + ((g.v[i] & keep) ^ sgn); r[i] = (uint32_t)c; c >>= 32; }
    const uint32_t w = (uint32_t)c ^ cin;                   // add: carried out; subtract: borrowed
    const uint32_t k = (0u - w) & 38u;
    c = cin;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)r[i] + (i ? sgn : (k ^ sgn)); r[i] = (uint32_t)c; c >>= 32; }
    const uint32_t w2 = (uint32_t)c ^ cin;
    r[0] += (0u - w2) & ((38u ^ sgn) + cin);                // +38 or -38; cannot carry or borrow again
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}
AFC_HD void fe_addsub_m(fe& h, const fe& f, const fe& g, uint32_t keep, uint32_t sgn) {
#if AFC_DEVICE_CODE && AFC_FE_PTX
    const uint32_t cin = sgn & 1u;
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7;             // (g & keep) ^ sgn as ONE LOP3 each (left to ptxas `keep` becomes a predicate and a SEL)
#define AFC_KEEP_XOR(t, w) asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(t) : "r"(g.v[w]), "r"(keep), "r"(sgn))
    AFC_KEEP_XOR(t0, 0); AFC_KEEP_XOR(t1, 1); AFC_KEEP_XOR(t2, 2); AFC_KEEP_XOR(t3, 3);
    AFC_KEEP_XOR(t4, 4); AFC_KEEP_XOR(t5, 5); AFC_KEEP_XOR(t6, 6); AFC_KEEP_XOR(t7, 7);
#undef AFC_KEEP_XOR
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7, c, dump;
    asm("add.cc.u32 %9, %26, 0xffffffff;\n\t"           // CC.CF = cin
        "addc.cc.u32 %0, %10, %18;\n\t"
        "addc.cc.u32 %1, %11, %19;\n\t"
        "addc.cc.u32 %2, %12, %20;\n\t"
        "addc.cc.u32 %3, %13, %21;\n\t"
        "addc.cc.u32 %4, %14, %22;\n\t"
        "addc.cc.u32 %5, %15, %23;\n\t"
        "addc.cc.u32 %6, %16, %24;\n\t"
        "addc.cc.u32 %7, %17, %25;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7), "=r"(c), "=r"(dump)
        : "r"(f.v[0]), "r"(f.v[1]), "r"(f.v[2]), "r"(f.v[3]), "r"(f.v[4]), "r"(f.v[5]), "r"(f.v[6]), "r"(f.v[7]),
          "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(t4), "r"(t5), "r"(t6), "r"(t7), "r"(cin));
    const uint32_t w = c ^ cin;
    const uint32_t kk = ((0u - w) & 38u) ^ sgn;
    uint32_t c2;
    asm("add.cc.u32 %9, %12, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %0, %10;\n\t"
        "addc.cc.u32 %1, %1, %11;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %11;\n\t"
        "addc.cc.u32 %5, %5, %11;\n\t"
        "addc.cc.u32 %6, %6, %11;\n\t"
        "addc.cc.u32 %7, %7, %11;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(c2), "=r"(dump)
        : "r"(kk), "r"(sgn), "r"(cin));
    const uint32_t w2 = c2 ^ cin;
    r0 += (0u - w2) & ((38u ^ sgn) + cin);
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
    fe_addsub_m_c(h, f, g, keep, sgn);
#endif
}

// ---------------------------------------------------------------------------------- 512 -> 256 fold
// h = (t[0..7] + 38 * t[8..15]) mod 2^256-38, weakly reduced
#ifndef AFC_FOLD_SHIFT
#define AFC_FOLD_SHIFT 0        // 1: 38 * hi as (hi << 5) + (hi << 2) + (hi << 1) on the ALU pipe instead of 8 IMAD.WIDE (experiment)
#endif
AFC_HD void fe_fold16(fe& h, const uint32_t* t) {
#if AFC_DEVICE_CODE && AFC_FE_PTX && AFC_FOLD_SHIFT
    // 38 * hi = (hi << 5) + (hi << 2) + (hi << 1): nine limbs each (funnel shifts, no carries), two add chains, then + lo.
    uint32_t a[9], b[9], c[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const uint32_t lo_ = i ? t[8 + i - 1] : 0u, hi_ = i < 8 ? t[8 + i] : 0u;
        a[i] = __funnelshift_l(lo_, hi_, 5); b[i] = __funnelshift_l(lo_, hi_, 2); c[i] = __funnelshift_l(lo_, hi_, 1);
    }
    asm("add.cc.u32 %0, %0, %9;\n\taddc.cc.u32 %1, %1, %10;\n\taddc.cc.u32 %2, %2, %11;\n\taddc.cc.u32 %3, %3, %12;\n\taddc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\taddc.cc.u32 %6, %6, %15;\n\taddc.cc.u32 %7, %7, %16;\n\taddc.u32 %8, %8, %17;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8])
        : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]));
    asm("add.cc.u32 %0, %0, %9;\n\taddc.cc.u32 %1, %1, %10;\n\taddc.cc.u32 %2, %2, %11;\n\taddc.cc.u32 %3, %3, %12;\n\taddc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\taddc.cc.u32 %6, %6, %15;\n\taddc.cc.u32 %7, %7, %16;\n\taddc.u32 %8, %8, %17;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8])
        : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(c[8]));
    uint32_t r0 = t[0], r1 = t[1], r2 = t[2], r3 = t[3], r4 = t[4], r5 = t[5], r6 = t[6], r7 = t[7], top = a[8];
    asm("add.cc.u32 %0, %0, %9;\n\taddc.cc.u32 %1, %1, %10;\n\taddc.cc.u32 %2, %2, %11;\n\taddc.cc.u32 %3, %3, %12;\n\taddc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\taddc.cc.u32 %6, %6, %15;\n\taddc.cc.u32 %7, %7, %16;\n\taddc.u32 %8, %8, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(top)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
    uint32_t k = top * 38u, c2;          // top <= 38
    asm("add.cc.u32 %0, %0, %9;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\taddc.cc.u32 %6, %6, 0;\n\taddc.cc.u32 %7, %7, 0;\n\taddc.u32 %8, 0, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(c2)
        : "r"(k));
    r0 += c2 * 38u;
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#elif AFC_DEVICE_CODE && AFC_FE_PTX
    uint32_t r0 = t[0], r1 = t[1], r2 = t[2], r3 = t[3], r4 = t[4], r5 = t[5], r6 = t[6], r7 = t[7], ce;
    const uint32_t k38 = 38u;
    // even columns of 38*hi accumulate in place (aligned pairs), one chain
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(ce)
        : "r"(t[8]), "r"(t[10]), "r"(t[12]), "r"(t[14]), "r"(k38));
    // odd columns: fresh products at limbs (1,2) (3,4) (5,6) (7,8)
    uint32_t o0, o1, o2, o3, o4, o5, o6, o7;
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;\n\t"
        : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3), "=r"(o4), "=r"(o5), "=r"(o6), "=r"(o7)
        : "r"(t[9]), "r"(t[11]), "r"(t[13]), "r"(t[15]), "r"(k38));
    uint32_t top;
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %15, %16;\n\t"
        : "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(top)
        : "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6), "r"(o7), "r"(ce));
    uint32_t k = top * 38u, c2;          // top <= 39
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "=r"(c2)
        : "r"(k));
    r0 += c2 * 38u;
    h.v[0] = r0; h.v[1] = r1; h.v[2] = r2; h.v[3] = r3; h.v[4] = r4; h.v[5] = r5; h.v[6] = r6; h.v[7] = r7;
#else
    fe_fold16_c(h, t);
#endif
}

// ---------------------------------------------------------------------------------- multiply
#if AFC_DEVICE_CODE && AFC_FE_PTX
// acc[0..7] += {x0,x1,x2,x3} * b at pairs (0,1)(2,3)(4,5)(6,7); carry-out -> top (fresh limb)
#define AFC_CHAIN_ACC_TOP(a0, a1, a2, a3, a4, a5, a6, a7, top, x0, x1, x2, x3, b)                     \
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                          \
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                         \
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                        \
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                        \
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                        \
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                        \
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                        \
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                        \
        "addc.u32 %8, 0, 0;\n\t"                                                                      \
        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7), "=r"(top)   \
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b))
// acc[0..6] += ... ; the last product's high half lands in a fresh limb a7
#define AFC_CHAIN_ACC_NEW(a0, a1, a2, a3, a4, a5, a6, a7, x0, x1, x2, x3, b)                          \
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"                                                          \
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"                                                         \
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"                                                         \
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"                                                         \
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"                                                        \
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"                                                        \
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"                                                        \
        "madc.hi.u32 %7, %11, %12, 0;\n\t"                                                            \
        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "=r"(a7)              \
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b))
// fresh products, no carries
#define AFC_ROW_FRESH(a0, a1, a2, a3, a4, a5, a6, a7, x0, x1, x2, x3, b)                              \
    asm("mul.lo.u32 %0, %8, %12;\n\t"                                                                 \
        "mul.hi.u32 %1, %8, %12;\n\t"                                                                 \
        "mul.lo.u32 %2, %9, %12;\n\t"                                                                 \
        "mul.hi.u32 %3, %9, %12;\n\t"                                                                 \
        "mul.lo.u32 %4, %10, %12;\n\t"                                                                \
        "mul.hi.u32 %5, %10, %12;\n\t"                                                                \
        "mul.lo.u32 %6, %11, %12;\n\t"                                                                \
        "mul.hi.u32 %7, %11, %12;\n\t"                                                                \
        : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7)              \
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b))
#endif

// t[0..15] = a * b (full 512-bit product)
AFC_HD void fe_mul_wide(uint32_t* t, const uint32_t* a, const uint32_t* b) {
#if AFC_DEVICE_CODE && AFC_FE_PTX
    // E[k] holds limb k (even-column products: i+j even); O[k] holds limb k+1 (odd-column products)
    uint32_t E[16], O[15];
    AFC_ROW_FRESH(E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], a[0], a[2], a[4], a[6], b[0]);
    AFC_ROW_FRESH(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], a[1], a[3], a[5], a[7], b[0]);
    E[8] = 0;
    // row 1 (odd): E takes odd j at E[2..9] (E[9] fresh); O takes even j at O[0..7], top O[8]
    AFC_CHAIN_ACC_NEW(E[2], E[3], E[4], E[5], E[6], E[7], E[8], E[9], a[1], a[3], a[5], a[7], b[1]);
    AFC_CHAIN_ACC_TOP(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], O[8], a[0], a[2], a[4], a[6], b[1]);
    // row 2 (even): E takes even j at E[2..9], top E[10]; O takes odd j at O[2..9] (O[9] fresh)
    AFC_CHAIN_ACC_TOP(E[2], E[3], E[4], E[5], E[6], E[7], E[8], E[9], E[10], a[0], a[2], a[4], a[6], b[2]);
    AFC_CHAIN_ACC_NEW(O[2], O[3], O[4], O[5], O[6], O[7], O[8], O[9], a[1], a[3], a[5], a[7], b[2]);
    // row 3
    AFC_CHAIN_ACC_NEW(E[4], E[5], E[6], E[7], E[8], E[9], E[10], E[11], a[1], a[3], a[5], a[7], b[3]);
    AFC_CHAIN_ACC_TOP(O[2], O[3], O[4], O[5], O[6], O[7], O[8], O[9], O[10], a[0], a[2], a[4], a[6], b[3]);
    // row 4
    AFC_CHAIN_ACC_TOP(E[4], E[5], E[6], E[7], E[8], E[9], E[10], E[11], E[12], a[0], a[2], a[4], a[6], b[4]);
    AFC_CHAIN_ACC_NEW(O[4], O[5], O[6], O[7], O[8], O[9], O[10], O[11], a[1], a[3], a[5], a[7], b[4]);
    // row 5
    AFC_CHAIN_ACC_NEW(E[6], E[7], E[8], E[9], E[10], E[11], E[12], E[13], a[1], a[3], a[5], a[7], b[5]);
    AFC_CHAIN_ACC_TOP(O[4], O[5], O[6], O[7], O[8], O[9], O[10], O[11], O[12], a[0], a[2], a[4], a[6], b[5]);
    // row 6
    AFC_CHAIN_ACC_TOP(E[6], E[7], E[8], E[9], E[10], E[11], E[12], E[13], E[14], a[0], a[2], a[4], a[6], b[6]);
    AFC_CHAIN_ACC_NEW(O[6], O[7], O[8], O[9], O[10], O[11], O[12], O[13], a[1], a[3], a[5], a[7], b[6]);
    // row 7
    AFC_CHAIN_ACC_NEW(E[8], E[9], E[10], E[11], E[12], E[13], E[14], E[15], a[1], a[3], a[5], a[7], b[7]);
    AFC_CHAIN_ACC_TOP(O[6], O[7], O[8], O[9], O[10], O[11], O[12], O[13], O[14], a[0], a[2], a[4], a[6], b[7]);
    // t = E + (O << 32)
    t[0] = E[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;\n\t"
        : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]),
          "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]),
          "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
          "r"(O[0]), "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]),
          "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]), "r"(O[14]));
#else
    fe_mul_wide_c(t, a, b);
#endif
}

#if AFC_DEVICE_CODE && AFC_FE_PTX
// r[0..7] = a[0..3] * b[0..3] (256-bit product) with the same even/odd carry-chain scheme: 16 IMAD.WIDE + 10 adds.
__device__ __forceinline__ void mul4x4_ptx(uint32_t* r, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    uint32_t E0, E1, E2, E3, E4, E5, E6, E7, O0, O1, O2, O3, O4, O5, O6;
    // row 0
    asm("mul.lo.u32 %0, %8, %12;\n\tmul.hi.u32 %1, %8, %12;\n\tmul.lo.u32 %2, %10, %12;\n\tmul.hi.u32 %3, %10, %12;\n\t"
        "mul.lo.u32 %4, %9, %12;\n\tmul.hi.u32 %5, %9, %12;\n\tmul.lo.u32 %6, %11, %12;\n\tmul.hi.u32 %7, %11, %12;"
        : "=r"(E0), "=r"(E1), "=r"(E2), "=r"(E3), "=r"(O0), "=r"(O1), "=r"(O2), "=r"(O3)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0));
    // row 1: E (2,3) += a1 b1, (4,5) fresh a3 b1;  O (0,1) += a0 b1, (2,3) += a2 b1, carry -> O4
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %6, 0;\n\tmadc.hi.u32 %3, %5, %6, 0;"
        : "+r"(E2), "+r"(E3), "=r"(E4), "=r"(E5) : "r"(a1), "r"(a3), "r"(b1));
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\tmadc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(O0), "+r"(O1), "+r"(O2), "+r"(O3), "=r"(O4) : "r"(a0), "r"(a2), "r"(b1));
    // row 2: E (2,3) += a0 b2, (4,5) += a2 b2, carry -> E6;  O (2,3) += a1 b2, (4,5): O4 += lo, O5 fresh (a3 b2)
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\tmadc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(E2), "+r"(E3), "+r"(E4), "+r"(E5), "=r"(E6) : "r"(a0), "r"(a2), "r"(b2));
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.u32 %3, %5, %6, 0;"
        : "+r"(O2), "+r"(O3), "+r"(O4), "=r"(O5) : "r"(a1), "r"(a3), "r"(b2));
    // row 3: E (4,5) += a1 b3, (6,7): E6 += lo, E7 fresh (a3 b3);  O (2,3) += a0 b3, (4,5) += a2 b3, carry -> O6
    asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\tmadc.hi.cc.u32 %1, %4, %6, %1;\n\tmadc.lo.cc.u32 %2, %5, %6, %2;\n\tmadc.hi.u32 %3, %5, %6, 0;"
        : "+r"(E4), "+r"(E5), "+r"(E6), "=r"(E7) : "r"(a1), "r"(a3), "r"(b3));
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\tmadc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(O2), "+r"(O3), "+r"(O4), "+r"(O5), "=r"(O6) : "r"(a0), "r"(a2), "r"(b3));
    // r = E + (O << 32)
    r[0] = E0;
    asm("add.cc.u32 %0, %7, %14;\n\taddc.cc.u32 %1, %8, %15;\n\taddc.cc.u32 %2, %9, %16;\n\taddc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\taddc.cc.u32 %5, %12, %19;\n\taddc.u32 %6, %13, %20;"
        : "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(E1), "r"(E2), "r"(E3), "r"(E4), "r"(E5), "r"(E6), "r"(E7), "r"(O0), "r"(O1), "r"(O2), "r"(O3), "r"(O4), "r"(O5), "r"(O6));
}

// One-level Karatsuba: 3 x (4x4) = 48 IMAD.WIDE instead of 64, paid for with ~60 adds on the otherwise idle ALU pipe
// (ncu on sm_100a: the 64-bit multiplier pipe "fmaheavy" is the binding unit, IMAD.WIDE issues once per 4 cycles/SMSP).
__device__ __forceinline__ void fe_mul_wide_kara(uint32_t* t, const uint32_t* a, const uint32_t* b) {
    uint32_t L[8], H[8], M[9];
    mul4x4_ptx(L, a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]);
    mul4x4_ptx(H, a[4], a[5], a[6], a[7], b[4], b[5], b[6], b[7]);
    uint32_t sa0, sa1, sa2, sa3, ca, sb0, sb1, sb2, sb3, cb;
    asm("add.cc.u32 %0, %5, %9;\n\taddc.cc.u32 %1, %6, %10;\n\taddc.cc.u32 %2, %7, %11;\n\taddc.cc.u32 %3, %8, %12;\n\taddc.u32 %4, 0, 0;"
        : "=r"(sa0), "=r"(sa1), "=r"(sa2), "=r"(sa3), "=r"(ca) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
    asm("add.cc.u32 %0, %5, %9;\n\taddc.cc.u32 %1, %6, %10;\n\taddc.cc.u32 %2, %7, %11;\n\taddc.cc.u32 %3, %8, %12;\n\taddc.u32 %4, 0, 0;"
        : "=r"(sb0), "=r"(sb1), "=r"(sb2), "=r"(sb3), "=r"(cb) : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    mul4x4_ptx(M, sa0, sa1, sa2, sa3, sb0, sb1, sb2, sb3);
    // (sa + ca W)(sb + cb W) = M + (ca sb + cb sa) W + ca cb W^2      (W = 2^128)
    uint32_t ma = 0u - ca, mb = 0u - cb;
    uint32_t x0 = (sb0 & ma), x1 = (sb1 & ma), x2 = (sb2 & ma), x3 = (sb3 & ma);
    uint32_t y0 = (sa0 & mb), y1 = (sa1 & mb), y2 = (sa2 & mb), y3 = (sa3 & mb);
    uint32_t m8 = ca & cb;
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, 0;"
        : "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(m8) : "r"(x0), "r"(x1), "r"(x2), "r"(x3));
    asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, %6;\n\taddc.cc.u32 %2, %2, %7;\n\taddc.cc.u32 %3, %3, %8;\n\taddc.u32 %4, %4, 0;"
        : "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(m8) : "r"(y0), "r"(y1), "r"(y2), "r"(y3));
    M[8] = m8;
    // cross = M - L - H  (non-negative, < 2^258)
    asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, %10;\n\tsubc.cc.u32 %2, %2, %11;\n\tsubc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\tsubc.cc.u32 %5, %5, %14;\n\tsubc.cc.u32 %6, %6, %15;\n\tsubc.cc.u32 %7, %7, %16;\n\tsubc.u32 %8, %8, 0;"
        : "+r"(M[0]), "+r"(M[1]), "+r"(M[2]), "+r"(M[3]), "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8])
        : "r"(L[0]), "r"(L[1]), "r"(L[2]), "r"(L[3]), "r"(L[4]), "r"(L[5]), "r"(L[6]), "r"(L[7]));
    asm("sub.cc.u32 %0, %0, %9;\n\tsubc.cc.u32 %1, %1, %10;\n\tsubc.cc.u32 %2, %2, %11;\n\tsubc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\tsubc.cc.u32 %5, %5, %14;\n\tsubc.cc.u32 %6, %6, %15;\n\tsubc.cc.u32 %7, %7, %16;\n\tsubc.u32 %8, %8, 0;"
        : "+r"(M[0]), "+r"(M[1]), "+r"(M[2]), "+r"(M[3]), "+r"(M[4]), "+r"(M[5]), "+r"(M[6]), "+r"(M[7]), "+r"(M[8])
        : "r"(H[0]), "r"(H[1]), "r"(H[2]), "r"(H[3]), "r"(H[4]), "r"(H[5]), "r"(H[6]), "r"(H[7]));
    // t = L + cross W + H W^2
    t[0] = L[0]; t[1] = L[1]; t[2] = L[2]; t[3] = L[3];
    asm("add.cc.u32 %0, %12, %24;\n\taddc.cc.u32 %1, %13, %25;\n\taddc.cc.u32 %2, %14, %26;\n\taddc.cc.u32 %3, %15, %27;\n\t"
        "addc.cc.u32 %4, %16, %28;\n\taddc.cc.u32 %5, %17, %29;\n\taddc.cc.u32 %6, %18, %30;\n\taddc.cc.u32 %7, %19, %31;\n\t"
        "addc.cc.u32 %8, %20, %32;\n\taddc.cc.u32 %9, %21, 0;\n\taddc.cc.u32 %10, %22, 0;\n\taddc.u32 %11, %23, 0;"
        : "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]),
          "=r"(t[14]), "=r"(t[15])
        : "r"(L[4]), "r"(L[5]), "r"(L[6]), "r"(L[7]), "r"(H[0]), "r"(H[1]), "r"(H[2]), "r"(H[3]), "r"(H[4]), "r"(H[5]), "r"(H[6]), "r"(H[7]),
          "r"(M[0]), "r"(M[1]), "r"(M[2]), "r"(M[3]), "r"(M[4]), "r"(M[5]), "r"(M[6]), "r"(M[7]), "r"(M[8]));
}
#endif

#ifndef AFC_FE_KARATSUBA
#define AFC_FE_KARATSUBA 0      // measured on B200: 375 vs 332 cycles per warp-multiply — the extra adds do not overlap
#endif

AFC_HD void fe_mul(fe& h, const fe& f, const fe& g) {
    uint32_t t[16];
#if AFC_DEVICE_CODE && AFC_FE_PTX && AFC_FE_KARATSUBA
    fe_mul_wide_kara(t, f.v, g.v);
#else
    fe_mul_wide(t, f.v, g.v);
#endif
    fe_fold16(h, t);
}
// the 64-product schoolbook form, kept for the microbenchmark comparison
AFC_HD void fe_mul_schoolbook(fe& h, const fe& f, const fe& g) {
    uint32_t t[16];
    fe_mul_wide(t, f.v, g.v);
    fe_fold16(h, t);
}

#if AFC_DEVICE_CODE && AFC_FE_PTX
// Carry-chain primitives as single instructions.  `volatile` keeps their relative order; the carry flag
// lives in PTX's CC register, which only these instructions touch (the compiler never emits .cc forms).
#define AFC_MAD_LO_CC(acc, x, y)   asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(x), "r"(y))
#define AFC_MADC_LO_CC(acc, x, y)  asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(x), "r"(y))
#define AFC_MADC_HI_CC(acc, x, y)  asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(x), "r"(y))
#define AFC_MADC_HI_NEW(acc, x, y) asm volatile("madc.hi.u32 %0, %1, %2, 0;" : "=r"(acc) : "r"(x), "r"(y))
#define AFC_MADC_LO_NEW_CC(acc, x, y) asm volatile("madc.lo.cc.u32 %0, %1, %2, 0;" : "=r"(acc) : "r"(x), "r"(y))
#define AFC_ADDC_NEW(acc)          asm volatile("addc.u32 %0, 0, 0;" : "=r"(acc))
#define AFC_MUL_WIDE(lo, hi, x, y) asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(y))
#endif

// t[0..15] = a^2
AFC_HD void fe_sq_wide(uint32_t* t, const uint32_t* a) {
#if defined(AFC_FE_SQ_AS_MUL)
    fe_mul_wide(t, a, a);
#elif AFC_DEVICE_CODE && AFC_FE_PTX
    // S = sum_{i<j} a_i a_j 2^(32(i+j)) with even columns in E[k] (limb k) and odd columns in O[k] (limb k+1)
    uint32_t E[14], O[14];
    // row 0
    AFC_MUL_WIDE(O[0], O[1], a[0], a[1]); AFC_MUL_WIDE(O[2], O[3], a[0], a[3]);
    AFC_MUL_WIDE(O[4], O[5], a[0], a[5]); AFC_MUL_WIDE(O[6], O[7], a[0], a[7]);
    AFC_MUL_WIDE(E[2], E[3], a[0], a[2]); AFC_MUL_WIDE(E[4], E[5], a[0], a[4]); AFC_MUL_WIDE(E[6], E[7], a[0], a[6]);
    // row 1: O (2,3)(4,5)(6,7) += a1*{a2,a4,a6}, carry -> O[8];  E (4,5)(6,7) += a1*{a3,a5}, (8,9) fresh a1*a7
    AFC_MAD_LO_CC(O[2], a[1], a[2]); AFC_MADC_HI_CC(O[3], a[1], a[2]);
    AFC_MADC_LO_CC(O[4], a[1], a[4]); AFC_MADC_HI_CC(O[5], a[1], a[4]);
    AFC_MADC_LO_CC(O[6], a[1], a[6]); AFC_MADC_HI_CC(O[7], a[1], a[6]);
    AFC_ADDC_NEW(O[8]);
    AFC_MAD_LO_CC(E[4], a[1], a[3]); AFC_MADC_HI_CC(E[5], a[1], a[3]);
    AFC_MADC_LO_CC(E[6], a[1], a[5]); AFC_MADC_HI_CC(E[7], a[1], a[5]);
    AFC_MADC_LO_NEW_CC(E[8], a[1], a[7]); AFC_MADC_HI_NEW(E[9], a[1], a[7]);
    // row 2: O (4,5)(6,7) += a2*{a3,a5}, (8,9): O[8] += lo, O[9] fresh;  E (6,7)(8,9) += a2*{a4,a6}, carry -> E[10]
    AFC_MAD_LO_CC(O[4], a[2], a[3]); AFC_MADC_HI_CC(O[5], a[2], a[3]);
    AFC_MADC_LO_CC(O[6], a[2], a[5]); AFC_MADC_HI_CC(O[7], a[2], a[5]);
    AFC_MADC_LO_CC(O[8], a[2], a[7]); AFC_MADC_HI_NEW(O[9], a[2], a[7]);
    AFC_MAD_LO_CC(E[6], a[2], a[4]); AFC_MADC_HI_CC(E[7], a[2], a[4]);
    AFC_MADC_LO_CC(E[8], a[2], a[6]); AFC_MADC_HI_CC(E[9], a[2], a[6]);
    AFC_ADDC_NEW(E[10]);
    // row 3: O (6,7)(8,9) += a3*{a4,a6}, carry -> O[10];  E (8,9) += a3*a5, (10,11): E[10] += lo, E[11] fresh
    AFC_MAD_LO_CC(O[6], a[3], a[4]); AFC_MADC_HI_CC(O[7], a[3], a[4]);
    AFC_MADC_LO_CC(O[8], a[3], a[6]); AFC_MADC_HI_CC(O[9], a[3], a[6]);
    AFC_ADDC_NEW(O[10]);
    AFC_MAD_LO_CC(E[8], a[3], a[5]); AFC_MADC_HI_CC(E[9], a[3], a[5]);
    AFC_MADC_LO_CC(E[10], a[3], a[7]); AFC_MADC_HI_NEW(E[11], a[3], a[7]);
    // row 4: O (8,9) += a4*a5, (10,11): O[10] += lo, O[11] fresh;  E (10,11) += a4*a6, carry -> E[12]
    AFC_MAD_LO_CC(O[8], a[4], a[5]); AFC_MADC_HI_CC(O[9], a[4], a[5]);
    AFC_MADC_LO_CC(O[10], a[4], a[7]); AFC_MADC_HI_NEW(O[11], a[4], a[7]);
    AFC_MAD_LO_CC(E[10], a[4], a[6]); AFC_MADC_HI_CC(E[11], a[4], a[6]);
    AFC_ADDC_NEW(E[12]);
    // row 5: O (10,11) += a5*a6, carry -> O[12];  E (12,13): E[12] += lo, E[13] fresh
    AFC_MAD_LO_CC(O[10], a[5], a[6]); AFC_MADC_HI_CC(O[11], a[5], a[6]);
    AFC_ADDC_NEW(O[12]);
    AFC_MAD_LO_CC(E[12], a[5], a[7]); AFC_MADC_HI_NEW(E[13], a[5], a[7]);
    // row 6: O (12,13): O[12] += lo, O[13] fresh
    AFC_MAD_LO_CC(O[12], a[6], a[7]); AFC_MADC_HI_NEW(O[13], a[6], a[7]);
    // S = E + (O << 32): limbs 1..15 (E[0] = E[1] = 0)
    uint32_t S1 = O[0], S2, S3, S4, S5, S6, S7, S8, S9, S10, S11, S12, S13, S14, S15;
    asm("add.cc.u32 %0, %14, %26;\n\t"
        "addc.cc.u32 %1, %15, %27;\n\t"
        "addc.cc.u32 %2, %16, %28;\n\t"
        "addc.cc.u32 %3, %17, %29;\n\t"
        "addc.cc.u32 %4, %18, %30;\n\t"
        "addc.cc.u32 %5, %19, %31;\n\t"
        "addc.cc.u32 %6, %20, %32;\n\t"
        "addc.cc.u32 %7, %21, %33;\n\t"
        "addc.cc.u32 %8, %22, %34;\n\t"
        "addc.cc.u32 %9, %23, %35;\n\t"
        "addc.cc.u32 %10, %24, %36;\n\t"
        "addc.cc.u32 %11, %25, %37;\n\t"
        "addc.cc.u32 %12, %38, 0;\n\t"
        "addc.u32 %13, 0, 0;\n\t"
        : "=r"(S2), "=r"(S3), "=r"(S4), "=r"(S5), "=r"(S6), "=r"(S7), "=r"(S8), "=r"(S9), "=r"(S10), "=r"(S11),
          "=r"(S12), "=r"(S13), "=r"(S14), "=r"(S15)
        : "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]),
          "r"(E[12]), "r"(E[13]),
          "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]),
          "r"(O[11]), "r"(O[12]), "r"(O[13]));
    // T = 2S (funnel shifts, no carries) + diagonal squares
    uint32_t D[16];
#pragma unroll
    for (int i = 0; i < 8; i++) AFC_MUL_WIDE(D[2 * i], D[2 * i + 1], a[i], a[i]);
    uint32_t T1 = S1 << 1, T2 = __funnelshift_l(S1, S2, 1), T3 = __funnelshift_l(S2, S3, 1), T4 = __funnelshift_l(S3, S4, 1),
             T5 = __funnelshift_l(S4, S5, 1), T6 = __funnelshift_l(S5, S6, 1), T7 = __funnelshift_l(S6, S7, 1),
             T8 = __funnelshift_l(S7, S8, 1), T9 = __funnelshift_l(S8, S9, 1), T10 = __funnelshift_l(S9, S10, 1),
             T11 = __funnelshift_l(S10, S11, 1), T12 = __funnelshift_l(S11, S12, 1), T13 = __funnelshift_l(S12, S13, 1),
             T14 = __funnelshift_l(S13, S14, 1), T15 = __funnelshift_l(S14, S15, 1);
    t[0] = D[0];
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, %44;\n\t"
        : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]),
          "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
        : "r"(D[1]), "r"(D[2]), "r"(D[3]), "r"(D[4]), "r"(D[5]), "r"(D[6]), "r"(D[7]), "r"(D[8]),
          "r"(D[9]), "r"(D[10]), "r"(D[11]), "r"(D[12]), "r"(D[13]), "r"(D[14]), "r"(D[15]),
          "r"(T1), "r"(T2), "r"(T3), "r"(T4), "r"(T5), "r"(T6), "r"(T7), "r"(T8),
          "r"(T9), "r"(T10), "r"(T11), "r"(T12), "r"(T13), "r"(T14), "r"(T15));
#else
    fe_sq_wide_c(t, a);
#endif
}

AFC_HD void fe_sq(fe& h, const fe& f) {
    uint32_t t[16];
    fe_sq_wide(t, f.v);
    fe_fold16(h, t);
}

// Field-multiplication policy.  FeInline expands the ~100-instruction multiply at every call site (fastest per call,
// but the verify loop then spans ~80 KB of SASS and misses the instruction cache); FeCall routes through ONE
// out-of-line copy with operands passed in registers, which keeps the whole double-scalar loop I-cache resident.
struct FeInline {
    static AFC_HDM void mul(fe& h, const fe& f, const fe& g) { fe_mul(h, f, g); }
    static AFC_HDM void sq(fe& h, const fe& f) { fe_sq(h, f); }
};
#if !defined(AFC_HOSTSIM)
static __device__ __noinline__ fe fe_mul_call(fe f, fe g) { fe h; fe_mul(h, f, g); return h; }
static __device__ __noinline__ fe fe_sq_call(fe f) { fe h; fe_sq(h, f); return h; }
struct FeCall {
    static AFC_HDM void mul(fe& h, const fe& f, const fe& g) { h = fe_mul_call(f, g); }
    static AFC_HDM void sq(fe& h, const fe& f) { h = fe_sq_call(f); }
};
#endif



template <class F = FeInline>
AFC_HD void fe_sqn_t(fe& h, const fe& f, int n) {
    F::sq(h, f);
#pragma unroll 1
    for (int i = 1; i < n; i++) F::sq(h, h);
}

// ---------------------------------------------------------------------------------- encode / decode
// Go field.Element.SetBytes: bit 255 ignored, values >= p accepted (no canonical check).
AFC_HD void fe_frombytes_words(fe& h, const uint32_t* le_words) {
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = le_words[i];
    h.v[7] &= 0x7fffffffu;
}

// canonical little-endian words of f mod p
AFC_HD void fe_towords(uint32_t* out, const fe& f) {
    uint32_t r[8];
    uint64_t c;
    // fold bit 255: x = (x mod 2^255) + 19*(x >> 255)   -> x < 2^255 + 19
    c = (uint64_t)(f.v[7] >> 31) * 19u;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (i == 7) ? (f.v[7] & 0x7fffffffu) : f.v[i]; r[i] = (uint32_t)c; c >>= 32; }
    // y = x + 19; if y >= 2^255 then x >= p and x - p = y - 2^255
    uint32_t y[8];
    c = 19;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += r[i]; y[i] = (uint32_t)c; c >>= 32; }
    uint32_t ge = y[7] >> 31;           // 1 iff x >= p
    uint32_t m = 0u - ge;
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = (r[i] & ~m) | (y[i] & m);
    out[7] &= 0x7fffffffu;
}
AFC_HD int fe_isnegative(const fe& f) { uint32_t w[8]; fe_towords(w, f); return (int)(w[0] & 1u); }
AFC_HD int fe_iszero(const fe& f) {
    uint32_t w[8]; fe_towords(w, f);
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= w[i];
    return o == 0;
}
AFC_HD int fe_equal(const fe& a, const fe& b) { fe d; fe_sub(d, a, b); return fe_iszero(d); }

// ---------------------------------------------------------------------------------- exponentiations
// z^(2^250-1) and z^11 (shared prefix of the inversion and square-root chains)
template <class F = FeInline>
AFC_HD void fe_pow_2_250_m1(fe& out, fe& z11, const fe& z) {
    fe t0, t1, t2, t3;
    F::sq(t0, z);                               // 2
    fe_sqn_t<F>(t1, t0, 2);                          // 8
    F::mul(t1, z, t1);                          // 9
    F::mul(t0, t0, t1);                         // 11
    fe_copy(z11, t0);
    F::sq(t2, t0);                              // 22
    F::mul(t1, t1, t2);                         // 2^5-1
    fe_sqn_t<F>(t2, t1, 5);   F::mul(t1, t2, t1);    // 2^10-1
    fe_sqn_t<F>(t2, t1, 10);  F::mul(t2, t2, t1);    // 2^20-1
    fe_sqn_t<F>(t3, t2, 20);  F::mul(t2, t3, t2);    // 2^40-1
    fe_sqn_t<F>(t2, t2, 10);  F::mul(t1, t2, t1);    // 2^50-1
    fe_sqn_t<F>(t2, t1, 50);  F::mul(t2, t2, t1);    // 2^100-1
    fe_sqn_t<F>(t3, t2, 100); F::mul(t2, t3, t2);    // 2^200-1
    fe_sqn_t<F>(t2, t2, 50);  F::mul(out, t2, t1);   // 2^250-1
}
template <class F = FeInline>
AFC_HD void fe_invert(fe& out, const fe& z) {
    fe t, z11;
    fe_pow_2_250_m1<F>(t, z11, z);
    fe_sqn_t<F>(t, t, 5);
    F::mul(out, t, z11);                        // z^(2^255-21) = z^(p-2)
}
template <class F = FeInline>
AFC_HD void fe_pow22523(fe& out, const fe& z) {
    fe t, z11;
    fe_pow_2_250_m1<F>(t, z11, z);
    fe_sqn_t<F>(t, t, 2);
    F::mul(out, t, z);                          // z^(2^252-3) = z^((p-5)/8)
}

}  // namespace afc
