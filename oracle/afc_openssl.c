/* ORACLE — TEST INFRASTRUCTURE ONLY.  OpenSSL (system libcrypto 3.x) batch drivers.
 *
 * Purpose: (1) an independent RFC 8032 / RFC 2104 / FIPS 180-4 implementation to cross-check the
 * restatement in afc_oracle.c on honest and randomly-corrupted inputs; (2) the CPU stand-in for Go's
 * assembly-backed crypto packages when timing the reference's CPU path on the GPU box (no Go toolchain in
 * this image; SURVEY.md §8d "CPU baseline beside it").  One pthread per requested core over the same
 * packed inputs the GPU path takes.
 */
#include <openssl/evp.h>
#include <openssl/hmac.h>
#include <openssl/sha.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int kind; uint32_t lo, hi;
    const uint8_t *a, *b, *msgs; const uint64_t *off; const uint32_t *koff; uint8_t *out;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    EVP_MD_CTX *ctx = EVP_MD_CTX_new();
    for (uint32_t i = j->lo; i < j->hi; i++) {
        const uint8_t *m = j->msgs + j->off[i];
        size_t len = (size_t)(j->off[i + 1] - j->off[i]);
        if (j->kind == 0) {
            SHA256(m, len, j->out + 32 * (size_t)i);
        } else if (j->kind == 1) {
            unsigned int ol = 32;
            HMAC(EVP_sha256(), j->a + j->koff[i], (int)(j->koff[i + 1] - j->koff[i]), m, len, j->out + 32 * (size_t)i, &ol);
        } else if (j->kind == 2) {
            EVP_PKEY *pk = EVP_PKEY_new_raw_public_key(EVP_PKEY_ED25519, NULL, j->a + 32 * (size_t)i, 32);
            int ok = 0;
            if (pk) {
                EVP_MD_CTX_reset(ctx);
                if (EVP_DigestVerifyInit(ctx, NULL, NULL, NULL, pk) == 1)
                    ok = EVP_DigestVerify(ctx, j->b + 64 * (size_t)i, 64, m, len) == 1;
                EVP_PKEY_free(pk);
            }
            j->out[i] = (uint8_t)ok;
        } else if (j->kind == 3) {
            EVP_PKEY *sk = EVP_PKEY_new_raw_private_key(EVP_PKEY_ED25519, NULL, j->a + 32 * (size_t)i, 32);
            size_t sl = 64;
            EVP_MD_CTX_reset(ctx);
            EVP_DigestSignInit(ctx, NULL, NULL, NULL, sk);
            EVP_DigestSign(ctx, j->out + 64 * (size_t)i, &sl, m, len);
            EVP_PKEY_free(sk);
        }
    }
    EVP_MD_CTX_free(ctx);
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
void afx_sha256_batch(const uint8_t *msgs, const uint64_t *off, uint32_t n, uint8_t *out32, int nthreads) {
    job_t j; memset(&j, 0, sizeof j); j.kind = 0; j.msgs = msgs; j.off = off; j.out = out32; run(j, n, nthreads);
}
void afx_hmac_sha256_batch(const uint8_t *keys, const uint32_t *koff, const uint8_t *msgs, const uint64_t *off,
                           uint32_t n, uint8_t *out32, int nthreads) {
    job_t j; memset(&j, 0, sizeof j); j.kind = 1; j.a = keys; j.koff = koff; j.msgs = msgs; j.off = off; j.out = out32; run(j, n, nthreads);
}
void afx_ed25519_verify_batch(const uint8_t *pks, const uint8_t *sigs, const uint8_t *msgs, const uint64_t *off,
                              uint32_t n, uint8_t *ok, int nthreads) {
    job_t j; memset(&j, 0, sizeof j); j.kind = 2; j.a = pks; j.b = sigs; j.msgs = msgs; j.off = off; j.out = ok; run(j, n, nthreads);
}
void afx_ed25519_sign_batch(const uint8_t *seeds, const uint8_t *msgs, const uint64_t *off, uint32_t n,
                            uint8_t *sigs, int nthreads) {
    job_t j; memset(&j, 0, sizeof j); j.kind = 3; j.a = seeds; j.msgs = msgs; j.off = off; j.out = sigs; run(j, n, nthreads);
}
