// Internal launch API between the C-ABI layer (afcrypto.cu) and the kernel translation units
// (k_hash.cu, k_ed25519.cu).  All pointers are DEVICE pointers; every function only enqueues work on
// `s` and returns cudaGetLastError().  Each launch is counted (and optionally timed) through the LaunchLog.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace afc {
namespace launch {

// Launch accounting + optional per-kernel CUDA-event timing (afc_profile_*): when `profile` is set every kernel
// launch is bracketed by two events on its own stream; afcrypto.cu aggregates them by kernel name after a sync.
struct LaunchRec { const char* name; cudaEvent_t e0, e1; bool done; };
struct LaunchLog {
    unsigned long long n = 0;
    bool profile = false;
    LaunchRec* pool = nullptr;       // the context's pool of event pairs; one is claimed per launch (never more than are used)
    int cap = 0;
    int* next = nullptr;             // shared claim counter (atomic on the host)
    LaunchRec* cur = nullptr;
    cudaEvent_t begin(const char* name, cudaStream_t s) {
        ++n;
        if (!profile) return nullptr;
        const int idx = __atomic_fetch_add(next, 1, __ATOMIC_RELAXED);
        if (idx >= cap) return nullptr;
        cur = &pool[idx];
        cur->name = name; cur->done = false;
        cudaEventRecord(cur->e0, s);
        return cur->e1;
    }
    void end(cudaEvent_t e1, cudaStream_t s) { if (e1) { cudaEventRecord(e1, s); cur->done = true; } }
};
#define AFC_LAUNCH(lg, name, stream, ...)            \
    do {                                             \
        cudaEvent_t _e1 = (lg)->begin(name, stream); \
        __VA_ARGS__;                                 \
        (lg)->end(_e1, stream);                      \
    } while (0)

// ---- k_hash.cu
cudaError_t sha256_batch(const uint8_t* msgs, const uint64_t* off, uint32_t n, uint8_t* out32, cudaStream_t s, LaunchLog* lg);
cudaError_t hmac_sha256_batch(const uint8_t* keys, const uint32_t* koff, const uint8_t* msgs, const uint64_t* off, uint32_t n,
                              uint8_t* out32, cudaStream_t s, LaunchLog* lg);
// streaming SHA-256: 108-byte states (Go MarshalBinary layout) in/out, chunk i into state i, digest where final_flags[i]
cudaError_t sha256_update(uint8_t* states, const uint8_t* chunks, const uint64_t* off, uint32_t n, const uint8_t* final_flags, uint8_t* out32,
                          uint8_t* status, cudaStream_t s, LaunchLog* lg);
cudaError_t merkle_leaf_hashes(const uint8_t* leaves, const uint64_t* off, uint32_t n, uint8_t* out32, cudaStream_t s, LaunchLog* lg);
// One tree level.  in: node array whose element 0 has level-index s_idx; npairs pairs starting at element `first`
// (first = 1 when s_idx was odd and the left orphan in[0] merges with frontier[h]); out[carry..] receives the parents.
//   left_merge  : out[0] = H(frontier[h] || in[0])                         (then pairs go to out[1..])
//   right_orphan: frontier[h] = in[first + 2*npairs]
cudaError_t merkle_level(const uint8_t* in, uint8_t* out, uint64_t npairs, int left_merge, int right_orphan,
                         uint8_t* frontier, int h, cudaStream_t s, LaunchLog* lg);
// root from frontier (bits of `size` say which heights are present); size == 0 -> SHA-256("")
cudaError_t merkle_root(const uint8_t* frontier, uint64_t size, uint8_t* out32, cudaStream_t s, LaunchLog* lg);
cudaError_t merkle_level_promote(const uint8_t* in, uint64_t n_in, uint8_t* out, cudaStream_t s, LaunchLog* lg);
cudaError_t merkle_gather_proofs(const uint8_t* const* levels, const uint64_t* sizes, int n_levels, const uint64_t* indices, uint32_t m,
                                 uint8_t* out, uint32_t* lens, cudaStream_t s, LaunchLog* lg);
cudaError_t merkle_verify_inclusion(const uint8_t* leaf_hashes, const uint64_t* indices, uint64_t tree_size, const uint8_t* proofs,
                                    const uint32_t* proof_off, const uint8_t* root, uint32_t m, uint8_t* ok, cudaStream_t s, LaunchLog* lg);
cudaError_t merkle_verify_consistency(const uint64_t* first_sizes, const uint8_t* first_roots, uint64_t second_size, const uint8_t* second_root,
                                      const uint8_t* proofs, const uint32_t* proof_off, uint32_t m, uint8_t* ok, cudaStream_t s, LaunchLog* lg);
size_t json_scan_scratch_bytes(uint32_t n);
cudaError_t json_fill_sizes(const uint8_t* segs, const uint32_t* seg_off, const uint8_t* kinds, uint32_t F, const uint8_t* fields,
                            const uint64_t* field_off, uint32_t n, uint64_t* out_off, uint64_t* scan_scratch, cudaStream_t s, LaunchLog* lg);
cudaError_t json_fill(const uint8_t* segs, const uint32_t* seg_off, const uint8_t* kinds, uint32_t F, const uint8_t* fields,
                      const uint64_t* field_off, uint32_t n, const uint64_t* out_off, uint8_t* out, cudaStream_t s, LaunchLog* lg);
cudaError_t b64url_encode(const uint8_t* in, uint32_t item, uint32_t n, uint8_t* out, cudaStream_t s, LaunchLog* lg);
cudaError_t hex_encode(const uint8_t* in, uint64_t total, uint8_t* out, cudaStream_t s, LaunchLog* lg);
cudaError_t microbench_hash(int which, uint32_t iters, uint32_t blocks, uint32_t threads, uint32_t* sink, cudaStream_t s, LaunchLog* lg);

// ---- k_ed25519.cu
// comb: 32 x 128 ge_precomp (row 0 = the 128 small multiples of B)
size_t ed_tables_bytes();
cudaError_t ed_build_tables(void* comb, cudaStream_t s, LaunchLog* lg);
// Device buffers of the transparent issuer-key cache used by ed_verify_batch (see k_ed25519.cu); owned by afcrypto.cu.
// state words (uint32 each)
enum KcState {
    KS_HIGH = 0,       // ids ever handed out (high-water mark, <= max_keys)
    KS_NHOT = 1,       // credentials of the current call that go through tables
    KS_NCOLD = 2,      // credentials of the current call left to the generic kernel
    KS_DISTINCT = 3,   // distinct public keys in the current call
    KS_NBUILD = 4,     // tables being built in the current call
    KS_EPOCH = 5,      // call counter (LRU stamps)
    KS_NCAND = 6,      // distinct keys of the call that are not cached and are used often enough to deserve a table
    KS_NFREE = 7,      // entries of free_list
    KS_REBUILD = 8,    // the persistent hash table must be rebuilt (something was evicted)
    KS_EVICTED = 9,    // tables evicted in the current call
    KS_TOTAL_BUILT = 10, KS_TOTAL_EVICTED = 11, KS_CALLS = 12,
    KS_QTILE = 13,     // next position of the order the four-lane kernel hands to a warp
    KS_WORDS = 16
};
struct KeyCache {
    // persistent
    uint32_t* slots;        // open-addressing table: slot -> cache id, 0xffffffff = empty   (slot_mask + 1 entries)
    uint32_t slot_mask;
    uint32_t max_keys;
    uint8_t* cpks;          // max_keys x 32
    uint8_t* valid;         // max_keys: key decodes
    void* tabs;             // max_keys x 32 x 128 ge_precomp
    uint32_t* stamp;        // max_keys: epoch of the last call that used the key, 0 = free id
    uint32_t* free_list;    // max_keys: evicted ids waiting for reuse
    uint32_t* bucket;       // max_keys + 1: credentials per id in the current call, then the scatter cursors
    uint32_t* state;        // KS_WORDS
    uint32_t* build_list;   // max_keys: ids whose tables are built in the current call
    void* bases;            // max_keys x 32 x 3 ge_p3: chain output of the tables being built (indexed by build position)
    // per call (capacity >= n)
    uint32_t* bslots;       // batch de-duplication table (bmask + 1 entries, >= 2n)
    uint32_t bmask;
    uint32_t* rep;          // credential -> representative credential (first with the same public key)
    uint32_t* kid;          // representative -> cache id | KC_COLD
    uint32_t* cnt;          // representative -> number of credentials with that key
    uint32_t* dlist;        // distinct representatives
    uint32_t* cand;         // representatives that want a table (at most n / KC_AMORTISE)
    uint32_t* perm;         // hot credentials bucketed by cache id (KS_NHOT entries)
    uint32_t* cold;         // cold credentials (KS_NCOLD entries)
    void* pts;              // 96 bytes per hot position: the projective result of the four-lane kernel, until k_ed_quad_finish
    // the cache work of a call runs on `side` (a high-priority stream) and `side2` between ev_fork and ev_join, while the
    // caller's stream computes H(R||A||M), which does not depend on the cache
    cudaStream_t side, side2;
    cudaEvent_t ev_fork, ev_join, ev_plan, ev_rows, ev_chain[4];
};
cudaError_t ed_keycache_clear(const KeyCache& kc, cudaStream_t s, LaunchLog* lg);     // empties the persistent table (slots, counters)
// scratch_k: n * 32 bytes; kc: nullable (nullptr = always the generic Straus kernel)
cudaError_t ed_verify_batch(const void* comb, const uint8_t* pks, const uint8_t* sigs, const uint8_t* msgs, const uint64_t* off,
                            uint32_t n, uint8_t* ok, uint32_t* scratch_k, const KeyCache* kc, cudaStream_t s, LaunchLog* lg);
size_t ed_key_table_bytes(uint32_t n_keys);
size_t ed_key_bases_bytes(uint32_t n_keys);     // scratch for ed_build_key_tables
cudaError_t ed_build_key_tables(const uint8_t* pks, uint32_t n_keys, void* tabs, uint8_t* valid, void* bases_scratch, cudaStream_t s, LaunchLog* lg);
size_t ed_keyed_scratch_bytes(uint32_t n_keys, uint32_t n);
size_t ed_verify_pts_bytes(uint64_t n);         // KeyCache::pts for a call of n credentials (0 unless AFC_VERIFY_QUAD is set)
// scratch_perm: ed_keyed_scratch_bytes() bytes, 32-byte aligned: projective results, then the issuer-bucketed order (nullptr = credential order, one-thread kernel)
cudaError_t ed_verify_keyed_batch(const void* comb, const void* tabs, const uint8_t* valid, const uint8_t* key_pks, uint32_t n_keys,
                                  const uint32_t* key_index, const uint8_t* sigs, const uint8_t* msgs, const uint64_t* off, uint32_t n,
                                  uint8_t* ok, uint32_t* scratch_k, uint32_t* scratch_perm, cudaStream_t s, LaunchLog* lg);
// ct16: the 48 KB constant-time table (ed_build_ct_table) or nullptr for the fast, variable-time fixed-base multiplication
size_t ed_ct_table_bytes();
cudaError_t ed_build_ct_table(const void* comb, void* ct16, cudaStream_t s, LaunchLog* lg);
cudaError_t ed_sign_batch(const void* comb, const void* ct16, const uint8_t* seeds, const uint8_t* msgs, const uint64_t* off, uint32_t n,
                          uint8_t* sigs, cudaStream_t s, LaunchLog* lg);
// n_keys bounds key_index on the device (an index >= n_keys signs with key 0 instead of reading out of bounds; the host entry
// point has already rejected such a batch, the device-pointer one cannot look)
cudaError_t ed_sign_expanded_batch(const void* comb, const void* ct16, const uint8_t* expanded96, uint32_t n_keys, const uint32_t* key_index,
                                   const uint8_t* msgs, const uint64_t* off, uint32_t n, uint8_t* sigs, cudaStream_t s, LaunchLog* lg);
cudaError_t ed_expand_batch(const void* comb, const void* ct16, const uint8_t* seeds, uint32_t n, uint8_t* expanded96, uint8_t* pks_only,
                            cudaStream_t s, LaunchLog* lg);
// returns number of mismatches in *d_mismatch (device uint32)
cudaError_t ed_selftest(uint32_t iters, uint32_t* d_mismatch, cudaStream_t s, LaunchLog* lg);
cudaError_t microbench_fe(int which, uint32_t iters, uint32_t blocks, uint32_t threads, uint32_t* sink, cudaStream_t s, LaunchLog* lg);

}  // namespace launch
}  // namespace afc
