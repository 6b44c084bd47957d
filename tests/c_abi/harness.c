/* tests/c_abi/harness.c — a plain-C caller of libafcrypto.so, the closest thing to the cgo binding this image allows.
 *
 * The reference is Go; its adapter (go/pkg/crypto/cuda_cgo.go) reaches the library through cgo, i.e. through the
 * C compiler and nothing but include/afcrypto.h.  This program does exactly that with gcc: it includes the same header, links the
 * same .so, uses only what cgo can pass (pointers to packed buffers, sizes) and checks
 *   - RFC 8032 §7.1 TEST 1-3: public key, signature, verification; a corrupted signature is ok[i] = 0, not an error
 *     (replaces ed25519.NewKeyFromSeed / Sign / Verify at internal/services/vc_service.go:460-463,504)
 *   - RFC 4231 test case 2 (generateWebhookSignature, internal/services/webhook_dispatcher.go:470-474)
 *   - SHA-256("abc") and the streaming form of the same digest (payload_store.go:69-94)
 *   - one RFC 6962 append against the first roots of the Certificate-Transparency reference tree
 * Exit code 0 = all good; 77 = no usable GPU (afc_init -> AFC_ECUDA: the library has no CPU path, the Go adapter falls back to the
 * stdlib on its side of the boundary); anything else = a failed check.  Built by tests/c_abi/Makefile, run by tests/test_abi.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "afcrypto.h"

static int unhex(const char* s, uint8_t* out) {
    int n = 0;
    for (; s[0] && s[1]; s += 2, n++) { unsigned v; if (sscanf(s, "%2x", &v) != 1) return -1; out[n] = (uint8_t)v; }
    return n;
}
static int expect(const char* what, const uint8_t* got, const char* want_hex, int n) {
    uint8_t want[128];
    if (unhex(want_hex, want) != n || memcmp(got, want, (size_t)n) != 0) { fprintf(stderr, "MISMATCH %s\n", what); return 1; }
    return 0;
}

static const struct { const char *seed, *pk, *msg, *sig; } RFC8032[3] = {
    {"9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60", "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a", "",
     "e5564300c360ac729086e2cc806e828a84877f1eb8e5d974d873e065224901555fb8821590a33bacc61e39701cf9b46bd25bf5f0595bbe24655141438e7a100b"},
    {"4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb", "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c", "72",
     "92a009a9f0d4cab8720e820b5f642540a2b27b5416503f8fb3762223ebdb69da085ac1e43e15996e458f3613d0f11d8c387b2eaeb4302aeeb00d291612bb0c00"},
    {"c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7", "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025", "af82",
     "6291d657deec24024827e69c3abe01a30ce548a284743a445e3680d7db5ac3ac18ff9b538d16f290ae67f760984dc6594a7c15e9716ed28dc027beceea1ec40a"},
};

int main(void) {
    afc_ctx* ctx = NULL;
    int rc = afc_init(0, &ctx);
    if (rc == AFC_ECUDA) { printf("no usable GPU: %s (the library has no CPU path)\n", afc_strerror(rc)); return 77; }
    if (rc != AFC_OK) { fprintf(stderr, "afc_init: %s\n", afc_strerror(rc)); return 2; }
    int bad = 0;
    printf("%s, constant-time signing: %d\n", afc_version(), afc_sign_mode(ctx));

    /* ---- RFC 8032: packed batch of three (seeds 3 x 32, messages back to back, offsets n + 1) */
    uint8_t seeds[96], pks[96], want_pk[32], msgs[8], sigs[192], ok[3];
    uint64_t off[4] = {0, 0, 0, 0};
    for (int i = 0; i < 3; i++) {
        unhex(RFC8032[i].seed, seeds + 32 * i);
        int m = unhex(RFC8032[i].msg, msgs + off[i]);
        off[i + 1] = off[i] + (uint64_t)m;
    }
    rc = afc_ed25519_pubkey_batch(ctx, seeds, 3, pks);
    if (rc) { fprintf(stderr, "pubkey_batch: %s (%s)\n", afc_strerror(rc), afc_last_cuda_error(ctx)); return 3; }
    rc = afc_ed25519_sign_batch(ctx, seeds, msgs, off, 3, sigs);
    if (rc) { fprintf(stderr, "sign_batch: %s\n", afc_strerror(rc)); return 3; }
    for (int i = 0; i < 3; i++) {
        (void)want_pk;
        bad += expect("RFC 8032 public key", pks + 32 * i, RFC8032[i].pk, 32);
        bad += expect("RFC 8032 signature", sigs + 64 * i, RFC8032[i].sig, 64);
    }
    rc = afc_ed25519_verify_batch(ctx, pks, sigs, msgs, off, 3, ok);
    if (rc || !(ok[0] && ok[1] && ok[2])) { fprintf(stderr, "verify of the RFC signatures failed (rc %d)\n", rc); bad++; }
    sigs[64 + 40] ^= 0x10;                                   /* a bad signature is a result, not an error */
    rc = afc_ed25519_verify_batch(ctx, pks, sigs, msgs, off, 3, ok);
    if (rc || !(ok[0] && !ok[1] && ok[2])) { fprintf(stderr, "corrupted signature not rejected (rc %d)\n", rc); bad++; }
    /* offsets that run backwards are refused, nothing is dereferenced */
    uint64_t bad_off[4] = {0, 2, 1, 3};
    if (afc_sha256_batch(ctx, msgs, bad_off, 3, sigs) != AFC_EINVAL) { fprintf(stderr, "non-monotone offsets accepted\n"); bad++; }

    /* ---- RFC 4231 case 2: key "Jefe", data "what do ya want for nothing?" */
    const uint8_t key[] = "Jefe", data[] = "what do ya want for nothing?";
    uint32_t koff[2] = {0, 4};
    uint64_t doff[2] = {0, 28};
    uint8_t tag[32];
    rc = afc_hmac_sha256_batch(ctx, key, koff, data, doff, 1, tag);
    if (rc) { fprintf(stderr, "hmac: %s\n", afc_strerror(rc)); return 3; }
    bad += expect("RFC 4231 case 2", tag, "5bdcc146bf60754e6a042426089575c75a003f089d2739839dec58b964ec3843", 32);

    /* ---- SHA-256("abc"), one-shot and as a stream of one final chunk */
    uint64_t aoff[2] = {0, 3};
    uint8_t dig[32], state[AFC_SHA256_STATE_BYTES], fin = 1;
    rc = afc_sha256_batch(ctx, (const uint8_t*)"abc", aoff, 1, dig);
    bad += rc != 0;
    bad += expect("SHA-256(abc)", dig, "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad", 32);
    afc_sha256_stream_init(state, 1);
    memset(dig, 0, 32);
    rc = afc_sha256_update_batch(ctx, state, (const uint8_t*)"abc", aoff, 1, &fin, dig);
    bad += rc != 0;
    bad += expect("streaming SHA-256(abc)", dig, "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad", 32);

    /* ---- RFC 6962: the first three leaves of the Certificate-Transparency reference tree ("", 00, 10) */
    afc_merkle* log = NULL;
    rc = afc_merkle_new(ctx, &log);
    if (rc) { fprintf(stderr, "merkle_new: %s\n", afc_strerror(rc)); return 3; }
    const uint8_t leaves[2] = {0x00, 0x10};
    uint64_t loff[4] = {0, 0, 1, 2}, size = 0;
    uint8_t root[32];
    rc = afc_merkle_append(log, leaves, loff, 1, root, &size);           /* the empty leaf */
    bad += rc != 0 || size != 1;
    bad += expect("CT root, 1 leaf", root, "6e340b9cffb37a989ca544e6bb780a2c78901d3fb33738768511a30617afa01d", 32);
    rc = afc_merkle_append(log, leaves, loff + 1, 2, root, &size);       /* 00, 10 */
    bad += rc != 0 || size != 3;
    bad += expect("CT root, 3 leaves", root, "aeb6bcfe274b70a14fb067a5e5578264db0fa9b51af5e0ba159158f329e06e77", 32);
    afc_merkle_free(log);

    printf("launches: %llu\n", (unsigned long long)afc_launch_count(ctx));
    afc_destroy(ctx);
    if (bad) { fprintf(stderr, "%d check(s) failed\n", bad); return 1; }
    printf("c_abi harness ok\n");
    return 0;
}
