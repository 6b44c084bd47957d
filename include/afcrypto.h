/* afcrypto.h — C ABI of libafcrypto.so: the B200 (sm_100a) drop-in for AgentField's cryptographic
 * identity-and-audit hot path.  Plain pointers and sizes only; no CUDA or torch types in signatures
 * (streams are passed as void*).  This is exactly what the reference-side cgo adapter binds
 * (INTEGRATION.md shows the Go stubs); each entry point cites the reference code it replaces, relative
 * to /root/reference/control-plane.
 *
 * Conventions
 *   - Batches are PACKED: one contiguous byte buffer + an offsets array of n+1 uint64 (msg i is
 *     bytes [off[i], off[i+1])).  No pointer-to-pointer (cgo cannot pass [][]byte without copying).
 *   - The caller owns every buffer; nothing is retained past return.
 *   - Return 0 on success, negative AFC_E* on failure.  An INVALID SIGNATURE IS NOT AN ERROR: it is
 *     ok[i] = 0 (mirrors how the reference turns failures into Valid:false, vc_service.go:242-289).
 *   - There is NO CPU implementation behind these calls: without a usable CUDA device afc_init fails
 *     with AFC_ECUDA and every other call fails with AFC_EINVAL (NULL ctx).  The reference-side adapter
 *     keeps Go's stdlib as ITS fallback, not this library.
 *   - Thread-safe and re-entrant: each call takes a stream + staging slot from the ctx pool.
 *   - "_dev" variants take DEVICE pointers and a cudaStream_t (as void*); they only enqueue work.
 */
#ifndef AFCRYPTO_H
#define AFCRYPTO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define AFC_OK 0
#define AFC_EINVAL (-1) /* bad argument / bad length (e.g. Go would panic on len(pk) != 32) */
#define AFC_ECUDA (-2)  /* CUDA runtime failure or no device; see afc_last_cuda_error() */
#define AFC_ENOMEM (-3) /* host or device allocation failed */
#define AFC_ENCCL (-4)  /* NCCL failure / NCCL not available */
#define AFC_ESTATE (-5) /* call not valid in the current state */

typedef struct afc_ctx afc_ctx;

/* ---- lifecycle ----------------------------------------------------------------------------------
 * device: CUDA ordinal this context drives (one context per GPU; the control plane shards its pending
 * credential queue round-robin over contexts — SURVEY.md §8e).  Builds the base-point tables on the
 * device. */
int afc_init(int device, afc_ctx** out);
void afc_destroy(afc_ctx* ctx);
const char* afc_strerror(int rc);
const char* afc_last_cuda_error(afc_ctx* ctx);
const char* afc_version(void);
/* device facts for benchmarks: SM count, SM clock kHz, global memory bytes */
int afc_device_info(afc_ctx* ctx, int* sm_count, int* clock_khz, uint64_t* mem_bytes);
/* number of kernel launches issued through this ctx since init (bench "gpu_launches") */
uint64_t afc_launch_count(afc_ctx* ctx);

/* pinned host staging for the adapter's packing buffers (cudaHostAlloc / cudaFreeHost) */
void* afc_alloc_pinned(size_t bytes);
void afc_free_pinned(void* p);
/* the same, but with the pages on the NUMA node of ctx's GPU whichever thread calls (the calling thread is bound to the CPUs in
 * /sys/bus/pci/devices/<gpu>/local_cpulist for the duration of the allocation): what a single process driving several contexts —
 * the Go control plane — uses for its packing pools; the library's own staging buffers are allocated the same way.
 * afc_numa_info: the GPU's NUMA node (-1 if unknown) and the number of CPUs next to it.  AFC_NUMA_BIND=0 disables the binding. */
void* afc_alloc_pinned_for(afc_ctx* ctx, size_t bytes);
int afc_numa_info(afc_ctx* ctx, int* numa_node, int* local_cpus);

/* ---- H1/H2: SHA-256 -----------------------------------------------------------------------------
 * replaces sha256.Sum256 in VCService.hashData (internal/services/vc_service.go:508-515), the
 * streaming digest in FilePayloadStore.SaveFromReader (internal/services/payload_store.go:69-94)
 * and the seed derivation hash (internal/services/did_service.go:515-521).  out32: n x 32 bytes. */
int afc_sha256_batch(afc_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, uint32_t n, uint8_t* out32);
int afc_sha256_batch_dev(afc_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, uint32_t n,
                         uint8_t* d_out32, void* stream);

/* H2 streaming: FilePayloadStore.SaveFromReader (internal/services/payload_store.go:45-97) hashes a payload while it streams to
 * disk, 32 KiB at a time (:154).  Many concurrent uploads = many streams advancing together: stream i absorbs chunk i into its
 * state.  A state is 108 bytes laid out like Go's crypto/sha256 (*digest).MarshalBinary ("sha\x03" || h big-endian || 64-byte
 * buffer || length big-endian) and always sits on a block boundary, so Go code can UnmarshalBinary it and carry on, and back.
 * Every chunk except a stream's last must be a multiple of 64 bytes (the reference's 32 KiB are); where final_flags[i] != 0 the
 * chunk may have any length (also 0), the digest goes to out32[i] and the state is finished.  final_flags may be NULL (none). */
#define AFC_SHA256_STATE_BYTES 108
int afc_sha256_stream_init(uint8_t* states /* n x 108 */, uint32_t n);
int afc_sha256_update_batch(afc_ctx* ctx, uint8_t* states /* n x 108, in/out */, const uint8_t* chunks, const uint64_t* chunk_off /* n+1 */,
                            uint32_t n, const uint8_t* final_flags /* n or NULL */, uint8_t* out32 /* n x 32, written where final */);
int afc_sha256_update_batch_dev(afc_ctx* ctx, uint8_t* d_states, const uint8_t* d_chunks, const uint64_t* d_chunk_off, uint32_t n,
                                const uint8_t* d_final_flags, uint8_t* d_out32, uint8_t* d_status /* n: 0 = malformed state */, void* stream);

/* ---- W1: HMAC-SHA256 ----------------------------------------------------------------------------
 * replaces generateWebhookSignature (internal/services/webhook_dispatcher.go:470-474): tag =
 * HMAC-SHA256(key = secret bytes, msg = body); the "sha256="+hex header text stays host-side.
 * Keys longer than 64 bytes are pre-hashed (RFC 2104, as Go's hmac.New).  key_off: n+1 uint32. */
int afc_hmac_sha256_batch(afc_ctx* ctx, const uint8_t* keys, const uint32_t* key_off, const uint8_t* msgs,
                          const uint64_t* msg_off, uint32_t n, uint8_t* out32);
int afc_hmac_sha256_batch_dev(afc_ctx* ctx, const uint8_t* d_keys, const uint32_t* d_key_off, const uint8_t* d_msgs,
                              const uint64_t* d_msg_off, uint32_t n, uint8_t* d_out32, void* stream);

/* ---- E2: Ed25519 verify -------------------------------------------------------------------------
 * replaces ed25519.Verify in VCService.verifyVCSignature (internal/services/vc_service.go:469-505),
 * verifyWorkflowVCSignature (:1589-1625) and EnhancedVCVerifier.verifyVCSignature
 * (internal/cli/vc_verification_enhanced.go:418-454).  pks: n x 32, sigs: n x 64, ok: n x 1 (0/1).
 * Bit-exact with Go 1.24: S must be canonical, sig[63]&0xE0 must be 0, non-canonical / small-order A
 * accepted, cofactor-less equation, R compared byte-wise with the canonical encoding.
 * (Go panics on len(pk) != 32: the adapter checks lengths before packing and re-panics.) */
int afc_ed25519_verify_batch(afc_ctx* ctx, const uint8_t* pks, const uint8_t* sigs, const uint8_t* msgs,
                             const uint64_t* msg_off, uint32_t n, uint8_t* ok);
int afc_ed25519_verify_batch_dev(afc_ctx* ctx, const uint8_t* d_pks, const uint8_t* d_sigs, const uint8_t* d_msgs,
                                 const uint64_t* d_msg_off, uint32_t n, uint8_t* d_ok, void* stream);

/* Transparent issuer-key cache behind afc_ed25519_verify_batch[_dev]: the batch's public keys are de-duplicated and counted on
 * the device and looked up in a persistent per-context cache of per-key tables (384 KB per key).  The decision is PER KEY:
 * a cached key goes through its table; a key that is not cached but signs at least 48 credentials of the batch gets a table
 * built inside the call (the least recently used tables that this call does not need are evicted when the cache is full);
 * every other credential is verified by the generic kernel in the same call.  Hot credentials are processed bucketed by
 * issuer.  Decisions are taken on the device, nothing synchronises, results are bit-identical whichever way a credential goes.
 * max_keys = 0 disables (always the generic kernel); default 4096 (1.6 GB), or env AFC_KEYCACHE_MAX_KEYS.
 * afc_keycache_info: cached_keys = tables currently held, last_mode = 1 if the last verify call sent any credential through
 * tables.  afc_keycache_stats: the split of the last call and the running totals. */
typedef struct afc_keycache_stats_t {
    uint32_t max_keys, cached_keys;
    uint32_t last_hot, last_cold;          /* credentials of the last call verified through tables / by the generic kernel */
    uint32_t last_distinct, last_built, last_evicted;
    uint32_t total_built, total_evicted, calls;
} afc_keycache_stats_t;
int afc_keycache_configure(afc_ctx* ctx, uint32_t max_keys);
int afc_keycache_info(afc_ctx* ctx, uint32_t* max_keys, uint32_t* cached_keys, uint32_t* last_mode);
int afc_keycache_stats(afc_ctx* ctx, afc_keycache_stats_t* out);
/* forget every cached table (stream-ordered: enqueued on `stream`, ordered after earlier verify calls); memory is kept */
int afc_keycache_clear(afc_ctx* ctx, void* stream);

/* ---- N1: keyed verification (the identity cache) ---------------------------------------------------------
 * The reference resolves every issuer DID from its own registry before it verifies (VCService.VerifyVC,
 * internal/services/vc_service.go:259 -> DIDService.ResolveDID, internal/services/did_service.go:368-473) and the
 * bulk audit loops over one workflow's VCs (vc_service.go:1442-1512): the verifier knows the key set in advance.
 * afc_keyset_new decodes each key once and builds its radix-256 table of -A on the device (384 KB per key);
 * afc_ed25519_verify_keyed_batch then needs no doublings and no per-credential key decoding (~4.5x fewer field
 * multiplications).  Results are bit-identical to afc_ed25519_verify_batch with pk = keys[key_index[i]]; an index
 * >= n_keys or a key that does not decode yields ok[i] = 0. */
typedef struct afc_keyset afc_keyset;
int afc_keyset_new(afc_ctx* ctx, const uint8_t* pks /* n_keys x 32 */, uint32_t n_keys, afc_keyset** out);
void afc_keyset_free(afc_keyset* ks);
int afc_keyset_info(afc_keyset* ks, uint32_t* n_keys, uint64_t* table_bytes);
int afc_ed25519_verify_keyed_batch(afc_ctx* ctx, afc_keyset* ks, const uint32_t* key_index, const uint8_t* sigs,
                                   const uint8_t* msgs, const uint64_t* msg_off, uint32_t n, uint8_t* ok);
int afc_ed25519_verify_keyed_batch_dev(afc_ctx* ctx, afc_keyset* ks, const uint32_t* d_key_index, const uint8_t* d_sigs,
                                       const uint8_t* d_msgs, const uint64_t* d_msg_off, uint32_t n, uint8_t* d_ok, void* stream);

/* ---- E1/K1: Ed25519 sign, public keys -----------------------------------------------------------
 * replaces ed25519.NewKeyFromSeed + ed25519.Sign in VCService.signVC / signWorkflowVC
 * (internal/services/vc_service.go:434-466, :686-718) and NewKeyFromSeed in
 * DIDService.derivePrivateKey (internal/services/did_service.go:515-525).
 * seeds: n x 32; sigs: n x 64 (R || S); pks: n x 32. */
int afc_ed25519_sign_batch(afc_ctx* ctx, const uint8_t* seeds, const uint8_t* msgs, const uint64_t* msg_off,
                           uint32_t n, uint8_t* sigs);
int afc_ed25519_sign_batch_dev(afc_ctx* ctx, const uint8_t* d_seeds, const uint8_t* d_msgs, const uint64_t* d_msg_off,
                               uint32_t n, uint8_t* d_sigs, void* stream);
int afc_ed25519_pubkey_batch(afc_ctx* ctx, const uint8_t* seeds, uint32_t n, uint8_t* pks);
int afc_ed25519_pubkey_batch_dev(afc_ctx* ctx, const uint8_t* d_seeds, uint32_t n, uint8_t* d_pks, void* stream);
/* expanded keys (the N1 identity cache): 96 bytes per key = clamped scalar s (32) || prefix (32) || pk (32).
 * Removes the two redundant fixed-base multiplications the reference does per issued VC
 * (did_service.go:585-599 via ResolveDID, vc_service.go:460). */
int afc_ed25519_expand_batch(afc_ctx* ctx, const uint8_t* seeds, uint32_t n, uint8_t* expanded96);
int afc_ed25519_sign_expanded_batch(afc_ctx* ctx, const uint8_t* expanded96, const uint32_t* key_index,
                                    uint32_t n_keys, const uint8_t* msgs, const uint64_t* msg_off, uint32_t n,
                                    uint8_t* sigs);
int afc_ed25519_expand_batch_dev(afc_ctx* ctx, const uint8_t* d_seeds, uint32_t n, uint8_t* d_expanded96, void* stream);
int afc_ed25519_sign_expanded_batch_dev(afc_ctx* ctx, const uint8_t* d_expanded96, const uint32_t* d_key_index,
                                        const uint8_t* d_msgs, const uint64_t* d_msg_off, uint32_t n,
                                        uint8_t* d_sigs, void* stream);
/* the same with the number of keys stated: an index >= n_keys can then not read key material out of bounds (it signs with key 0;
 * the host-buffer call rejects such a batch with AFC_EINVAL, a device-side index array cannot be inspected without a sync) */
int afc_ed25519_sign_expanded_keys_batch_dev(afc_ctx* ctx, const uint8_t* d_expanded96, uint32_t n_keys, const uint32_t* d_key_index,
                                             const uint8_t* d_msgs, const uint64_t* d_msg_off, uint32_t n, uint8_t* d_sigs, void* stream);

/* Secret scalars (the signing nonce r, and the private scalar when keys are expanded) are multiplied in CONSTANT TIME by default,
 * as Go's crypto/ed25519 does: signed radix 16, a 48 KB table staged in shared memory, every entry of a row read for every digit
 * and one kept by mask — no branch and no address depends on key material.  afc_sign_configure(ctx, 0) (or AFC_SIGN_CT=0) selects
 * the fast variable-time path (radix-65536 table gathers: 16 additions instead of 64) for deployments where nothing untrusted
 * shares the GPU.  Results are identical either way.  Seeds and expanded keys staged by the host-buffer calls are wiped from the
 * device slot and the pinned bounce buffer when the call completes.  afc_sign_mode: 1 = constant time. */
int afc_sign_configure(afc_ctx* ctx, int constant_time);
int afc_sign_mode(afc_ctx* ctx);

/* ---- M1: RFC 6962 Merkle audit log (NEW — the reference's chain check is a stub,
 * internal/cli/vc_verification_enhanced.go:531-534; nearest code generateWorkflowVCDocument,
 * internal/services/vc_service.go:525-632) -----------------------------------------------------------
 * leaf hash = SHA-256(0x00 || leaf), node = SHA-256(0x01 || l || r), split at the largest power of two
 * < n.  The log state (leaf count + <= 64 frontier hashes) lives in a afc_merkle handle on the device.
 * afc_merkle_append adds n leaves and returns the new root and size. */
typedef struct afc_merkle afc_merkle;
int afc_merkle_new(afc_ctx* ctx, afc_merkle** out);
void afc_merkle_free(afc_merkle* m);
int afc_merkle_append(afc_merkle* m, const uint8_t* leaves, const uint64_t* leaf_off, uint32_t n,
                      uint8_t root32[32], uint64_t* tree_size);
int afc_merkle_append_dev(afc_merkle* m, const uint8_t* d_leaves, const uint64_t* d_leaf_off, uint32_t n, void* stream);
/* append n already-hashed nodes (32 bytes each) as leaves-of-this-tree: used to fold per-GPU subtree
 * roots after the all-gather, and to resume a log from stored leaf hashes */
int afc_merkle_append_hashes(afc_merkle* m, const uint8_t* hashes32, uint32_t n, uint8_t root32[32], uint64_t* tree_size);
int afc_merkle_append_hashes_dev(afc_merkle* m, const uint8_t* d_hashes32, uint32_t n, void* stream);
int afc_merkle_root(afc_merkle* m, uint8_t root32[32], uint64_t* tree_size);
int afc_merkle_root_dev(afc_merkle* m, uint8_t* d_root32, void* stream);
/* checkpoint / resume: frontier = up to 64 (height, hash) entries; buffer of 8 + 64*32 bytes */
#define AFC_MERKLE_STATE_BYTES (8 + 64 * 32)
int afc_merkle_save(afc_merkle* m, uint8_t* state);
int afc_merkle_load(afc_merkle* m, const uint8_t* state);
/* one-shot: leaf hashes of a batch (SHA-256(0x00 || leaf)), n x 32 */
int afc_merkle_leaf_hashes_dev(afc_ctx* ctx, const uint8_t* d_leaves, const uint64_t* d_leaf_off, uint32_t n,
                               uint8_t* d_out32, void* stream);

/* ---- N4: audit proofs for bulk offline verification (`af vc verify`, internal/cli/vc.go:159-331; export bundles,
 * internal/handlers/ui/did.go:533-865) ------------------------------------------------------------------------
 * afc_merkle_tree materialises every level over a fixed set of leaf hashes (device memory, ~2n x 32 B) so that RFC 6962
 * §2.1.1 audit paths can be read out in bulk; afc_merkle_verify_inclusion_batch checks many paths against one root on
 * the device (RFC 9162 §2.1.3.2).  proof_off: m+1 offsets in units of 32-byte nodes. */
typedef struct afc_merkle_tree afc_merkle_tree;
int afc_merkle_tree_build(afc_ctx* ctx, const uint8_t* leaf_hashes32, uint64_t n, afc_merkle_tree** out);
int afc_merkle_tree_build_dev(afc_ctx* ctx, const uint8_t* d_leaf_hashes32, uint64_t n, afc_merkle_tree** out);
void afc_merkle_tree_free(afc_merkle_tree* t);
int afc_merkle_tree_root(afc_merkle_tree* t, uint8_t root32[32], uint64_t* n_leaves, uint32_t* max_proof_nodes);
int afc_merkle_tree_inclusion_proofs(afc_merkle_tree* t, const uint64_t* indices, uint32_t m, uint8_t* proofs /* m x max_proof_nodes x 32 */,
                                     uint32_t* proof_lens /* m, in nodes */);
int afc_merkle_verify_inclusion_batch(afc_ctx* ctx, const uint8_t* leaf_hashes32, const uint64_t* indices, uint64_t tree_size,
                                      const uint8_t* proofs, const uint32_t* proof_off, const uint8_t root32[32], uint32_t m, uint8_t* ok);
/* RFC 6962 §2.1.2 consistency proof between the first `first` leaves (0 < first <= n) and the whole tree: at most
 * 2 * max_proof_nodes nodes are written to `proof`, their number to *n_nodes.  afc_merkle_verify_consistency_batch checks
 * m earlier checkpoints (first_sizes[i], first_roots32[i]) against one later (second_size, second_root32) on the device
 * (RFC 9162 §2.1.4.2); ok[i] = 0 for first_sizes[i] == 0 or > second_size. */
int afc_merkle_tree_consistency_proof(afc_merkle_tree* t, uint64_t first, uint8_t* proof, uint32_t* n_nodes);
int afc_merkle_verify_consistency_batch(afc_ctx* ctx, const uint64_t* first_sizes, const uint8_t* first_roots32, uint64_t second_size,
                                        const uint8_t second_root32[32], const uint8_t* proofs, const uint32_t* proof_off, uint32_t m, uint8_t* ok);

/* ---- multi-GPU: independent shards, one exchange step --------------------------------
 * Each rank appends its contiguous, 2^k-aligned leaf range to its own afc_merkle; the 32-byte subtree
 * roots are all-gathered (NCCL over NVLink) and every rank folds the top levels redundantly.
 * NCCL is dlopen'ed lazily (libnccl.so.2); unique_id is ncclUniqueId (128 bytes). */
int afc_comm_unique_id(uint8_t id128[128]);
int afc_comm_init(afc_ctx* ctx, int nranks, int rank, const uint8_t id128[128]);
int afc_comm_allgather_roots(afc_ctx* ctx, const uint8_t local_root32[32], uint8_t* all_roots /* nranks x 32 */);
int afc_comm_destroy(afc_ctx* ctx);

/* ---- N3: text codecs either side of the kernels (device pointers) ------------------------------------------------------
 * base64.RawURLEncoding of signatures / digests (internal/services/vc_service.go:465,514) and hex.EncodeToString of the
 * webhook tag (internal/services/webhook_dispatcher.go:473) for n fixed-size records of item_bytes each.
 * out: n x ceil(item_bytes*4/3) characters (no padding), resp. n x 2*item_bytes lowercase hex characters. */
int afc_b64url_encode_fixed_dev(afc_ctx* ctx, const uint8_t* d_in, uint32_t item_bytes, uint32_t n, uint8_t* d_out, void* stream);
int afc_hex_encode_fixed_dev(afc_ctx* ctx, const uint8_t* d_in, uint32_t item_bytes, uint32_t n, uint8_t* d_out, void* stream);

/* ---- canonical form on the device ---------------------------------------------------------------
 * What the reference signs is json.Marshal of a fixed struct (VCDocument pkg/types/did_types.go:135-220; marshalled at
 * internal/services/vc_service.go:436-439 for signing and :201 for storage; ExecutionWebhookPayload pkg/types/webhook.go:42-53,
 * marshalled at internal/services/webhook_dispatcher.go:286-300): constant text interleaved with a fixed number of values.
 * n documents out[i] = seg[0] || v(i,0) || seg[1] || ... || v(i,F-1) || seg[F]; v(i,f) = d_fields[d_field_off[i*F+f] .. [i*F+f+1])
 * written according to d_kinds[f]: AFC_JSON_STRING = Go encoding/json string escaping (HTML-safe; U+2028/9; invalid UTF-8
 * becomes \ufffd byte by byte) without the quotes — the template supplies them; AFC_JSON_RAW = copied (numbers, pre-rendered
 * optional members).  d_seg_off: F+2 offsets into d_segs.  afc_json_fill_sizes_dev leaves the n+1 output offsets in
 * d_out_off (and, if total_bytes is not NULL, synchronises and returns the total); afc_json_fill_dev writes the bytes. */
#define AFC_JSON_STRING 0
#define AFC_JSON_RAW 1
int afc_json_fill_sizes_dev(afc_ctx* ctx, const uint8_t* d_segs, const uint32_t* d_seg_off, const uint8_t* d_kinds, uint32_t n_fields,
                            const uint8_t* d_fields, const uint64_t* d_field_off, uint32_t n, uint64_t* d_out_off, uint64_t* total_bytes,
                            void* stream);
int afc_json_fill_dev(afc_ctx* ctx, const uint8_t* d_segs, const uint32_t* d_seg_off, const uint8_t* d_kinds, uint32_t n_fields,
                      const uint8_t* d_fields, const uint64_t* d_field_off, uint32_t n, const uint64_t* d_out_off, uint8_t* d_out,
                      void* stream);

/* ---- N2: batching ingest dispatcher (BASELINE.json configs[4]: sustained mixed ingest) ------------------------------
 * One "agent action" = Ed25519 signature over the credential bytes (expanded key `key_index` of the identity cache) +
 * HMAC-SHA256 webhook tag + one audit-log leaf (the signature) appended to an RFC 6962 log owned by the dispatcher.
 * Replaces the reference's one-per-request issuing (internal/handlers/did_handlers.go:192 ->
 * internal/services/vc_service.go:138) and its 4-goroutine webhook worker pool (internal/services/webhook_dispatcher.go:
 * 97-128, 261-320) with GPU batches: flush at `batch_max` actions or after `linger_us`.  Thread-safe; submit blocks only
 * when both staging buffers are full (back-pressure). */
typedef struct afc_ingest afc_ingest;
typedef struct afc_ingest_stats {
    uint64_t submitted, completed, batches;
    double avg_batch;
    uint32_t p50_us, p99_us, max_us;       /* submit -> results-on-host latency */
    uint64_t log_size;
    uint8_t log_root[32];
    int last_error;
} afc_ingest_stats;
int afc_ingest_new(afc_ctx* ctx, const uint8_t* expanded96, uint32_t n_keys, uint32_t batch_max, uint32_t linger_us,
                   uint32_t max_msg, uint32_t max_key, uint32_t max_body, afc_ingest** out);
void afc_ingest_free(afc_ingest* g);
int afc_ingest_submit(afc_ingest* g, uint32_t key_index, const uint8_t* msg, uint32_t msg_len, const uint8_t* hkey, uint32_t hkey_len,
                      const uint8_t* body, uint32_t body_len, uint64_t* ticket);
int afc_ingest_wait(afc_ingest* g, uint64_t ticket, uint8_t sig64[64], uint8_t tag32[32]);
int afc_ingest_flush(afc_ingest* g);
int afc_ingest_stats_get(afc_ingest* g, afc_ingest_stats* out);
/* native open-loop load generator: Poisson arrivals at rate_per_s for `seconds` from `producers` threads (32-byte HMAC keys) */
int afc_ingest_soak(afc_ingest* g, double rate_per_s, double seconds, uint32_t producers, uint32_t msg_len, uint32_t body_len,
                    uint64_t seed, afc_ingest_stats* out, double* achieved_rate, uint64_t* late_submits);

/* ---- diagnostics --------------------------------------------------------------------------------
 * afc_selftest: runs the PTX field arithmetic against the portable reference ON THE DEVICE for `iters`
 * random operands (returns number of mismatches, or negative error).
 * afc_microbench: register-only throughput of a primitive; which = 0 fe_mul, 1 fe_sq, 2 fe_add,
 * 3 sha256 compress, 4 sha512 compress; returns primitive-ops per second (whole GPU) in *ops_per_s. */
int afc_selftest(afc_ctx* ctx, uint32_t iters);
/* Per-kernel CUDA-event timing (tracing hook): between begin and end every kernel this ctx launches is
 * bracketed by two events on its own stream; afc_profile_end synchronises the device and returns one aggregated entry
 * per kernel name (returns the number of entries, or a negative error). */
typedef struct afc_profile_entry {
    char name[48];
    uint32_t count;
    double total_ms, min_ms, max_ms;
} afc_profile_entry;
int afc_profile_begin(afc_ctx* ctx, int max_launches);
int afc_profile_end(afc_ctx* ctx, afc_profile_entry* out, int cap);
int afc_microbench(afc_ctx* ctx, int which, uint32_t iters, double* ops_per_s, double* ms);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* AFCRYPTO_H */
