/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not linked into libafcrypto.so; only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load it.
 *
 * Plain-C CPU restatement of the arithmetic the reference reaches through the Go standard
 * library (toolchain go1.24.2, control-plane/go.mod:3-5) on the identity-and-audit hot path:
 *
 *   crypto/ed25519   vc_service.go:460,463,504,712,715,1624  cli/vc_verification_enhanced.go:453
 *   crypto/sha256    vc_service.go:513  did_service.go:517  payload_store.go:69
 *   crypto/hmac      webhook_dispatcher.go:470-474
 *   (RFC 6962 Merkle tree hash: new component)
 *
 * The algorithm structure follows Go's crypto/internal/fips140/edwards25519: 5x51-bit field limbs,
 * extended/P1xP1/P2/cached point forms, VarTimeDoubleScalarBaseMult with width-5 NAF for A and
 * width-8 NAF for B, radix-16 fixed-base multiplication for signing.
 *
 * All batch entry points take packed buffers + offset arrays (same layout as include/afcrypto.h).
 */
#ifndef AFC_ORACLE_H
#define AFC_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

void afo_sha256(const uint8_t *msg, size_t len, uint8_t out[32]);
void afo_sha512(const uint8_t *msg, size_t len, uint8_t out[64]);
void afo_hmac_sha256(const uint8_t *key, size_t klen, const uint8_t *msg, size_t mlen, uint8_t out[32]);

void afo_ed25519_pubkey(const uint8_t seed[32], uint8_t pk[32]);
void afo_ed25519_sign(const uint8_t seed[32], const uint8_t *msg, size_t len, uint8_t sig[64]);
/* returns 1 (valid) / 0 (invalid); Go's accept/reject rules */
int afo_ed25519_verify(const uint8_t pk[32], const uint8_t *msg, size_t len, const uint8_t sig[64]);

/* RFC 6962 */
void afo_merkle_leaf_hash(const uint8_t *leaf, size_t len, uint8_t out[32]);
void afo_merkle_root_from_hashes(const uint8_t *hashes /* n x 32 */, uint64_t n, uint8_t root[32]);

/* Batch forms: packed messages, offsets[n+1]; nthreads >= 1 (pthreads). */
void afo_sha256_batch(const uint8_t *msgs, const uint64_t *off, uint32_t n, uint8_t *out32, int nthreads);
void afo_hmac_sha256_batch(const uint8_t *keys, const uint32_t *koff, const uint8_t *msgs, const uint64_t *off,
                           uint32_t n, uint8_t *out32, int nthreads);
void afo_ed25519_verify_batch(const uint8_t *pks, const uint8_t *sigs, const uint8_t *msgs, const uint64_t *off,
                              uint32_t n, uint8_t *ok, int nthreads);
void afo_ed25519_sign_batch(const uint8_t *seeds, const uint8_t *msgs, const uint64_t *off, uint32_t n,
                            uint8_t *sigs, int nthreads);
void afo_ed25519_pubkey_batch(const uint8_t *seeds, uint32_t n, uint8_t *pks, int nthreads);
void afo_merkle_root(const uint8_t *leaves, const uint64_t *off, uint32_t n, uint8_t root[32], int nthreads);

#ifdef __cplusplus
}
#endif
#endif
