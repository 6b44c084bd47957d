// libafcrypto.so — C-ABI layer (include/afcrypto.h) over the sm_100a kernels.
//
// Host-side responsibilities only: context + stream pool, pinned/device staging, chunked
// copy/compute overlap for host-buffer calls, the RFC 6962 level schedule, optional NCCL (dlopen).
// There is deliberately NO CPU implementation of any primitive in this library: if CUDA is missing or
// fails, calls return AFC_ECUDA and the reference-side adapter decides what to do.
#include "../../include/afcrypto.h"
#include "afc_internal.h"
#include "afc_launch.h"

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <sched.h>
#include <mutex>
#include <string>
#include <vector>

using namespace afc;

namespace {

constexpr int kLanes = 4;                 // concurrent host-buffer calls per context
constexpr int kSlots = 4;                 // chunks in flight per call: copies keep streaming while an earlier chunk computes
constexpr uint32_t kChunkItems = 1u << 17;  // items per pipeline chunk for host-buffer calls (one full wave of the table-driven verify at G = 2)
constexpr size_t kChunkBytes = 96u << 20;   // and at most this many message bytes per chunk
constexpr uint32_t kMinChunkItems = 1u << 16;   // measured on 1 M x 512 B: 16 k -> 12.41 ms, 64 k -> 12.06 ms per call (fewer, larger copies)

struct DevBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc((void**)&p, want);
        // zeroed once: the kernels read whole aligned words and mask what lies past a message, so the slack behind the last
        // staged byte is read (never used); this keeps those reads defined without a memset per call
        if (e == cudaSuccess) e = cudaMemset(p, 0, want);
        // the memset runs on the legacy default stream and returns before it has happened; the buffer's users are NON-BLOCKING
        // streams, which do not order themselves after it: without this wait a copy enqueued next could be zeroed afterwards
        // (seen once as a wrong Merkle root right after a log was created)
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n, afc_ctx* ctx) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        p = (uint8_t*)afc_internal_pinned_alloc(ctx, want);        // on the NUMA node of the context's GPU
        if (!p) return cudaErrorMemoryAllocation;
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// One stage of a lane's ring of in-flight chunks
struct Slot {
    cudaStream_t stream = nullptr;
    DevBuf msgs, off, a, b, k, out, koff;   // a: pks/seeds/keys  b: sigs
    PinBuf h_in, h_out;                      // bounce buffers when the caller's memory is not pinned
};
struct Lane {
    Slot slot[kSlots];
    bool busy = false;
};

}  // namespace

struct afc_ctx {
    int device = 0;
    cudaDeviceProp prop{};
    void* comb = nullptr;
    void* ct16 = nullptr;               // 48 KB constant-time signing table (radix 16), gathered from comb at init
    bool sign_ct = true;                // secret scalars go through the constant-time fixed-base multiplication (AFC_SIGN_CT=0: fast path)
    Lane lanes[kLanes];
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<unsigned long long> launches{0};
    std::string last_error;
    // CPUs next to this GPU (sysfs local_cpulist of its PCI device): pinned staging is allocated from a thread bound to them
    cpu_set_t local_cpus;
    bool has_local_cpus = false;
    int numa_node = -1;
    // transparent issuer-key cache behind afc_ed25519_verify_batch (see k_ed25519.cu)
    launch::KeyCache kc{};
    bool kc_ready = false;
    uint32_t kc_max_keys = 4096;        // AFC_KEYCACHE_MAX_KEYS / afc_keycache_configure; 0 disables (4096 keys = 1.6 GB of the 180 GB)
    uint32_t kc_call_cap = 0;           // per-call scratch capacity (items)
    std::mutex kc_mu;
    cudaEvent_t kc_event = nullptr;     // orders successive users of the cache across streams
    // per-kernel CUDA-event profiling (afc_profile_begin / afc_profile_end)
    bool profile = false;
    std::vector<launch::LaunchRec> prof_recs;
    int prof_next = 0;                  // claimed with __atomic_fetch_add (launch::LaunchLog::begin)
    // NCCL (lazy)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int nranks = 0, rank = 0;
    cudaStream_t comm_stream = nullptr;
    uint8_t* d_comm_buf = nullptr;
};

struct afc_keyset {
    afc_ctx* ctx = nullptr;
    uint32_t n_keys = 0;
    uint8_t* d_pks = nullptr;      // n_keys x 32
    uint8_t* d_valid = nullptr;    // n_keys
    void* d_tabs = nullptr;        // n_keys x 32 x 128 ge_precomp
    size_t tab_bytes = 0;
};

struct afc_merkle_tree {
    afc_ctx* ctx = nullptr;
    uint64_t n = 0;
    std::vector<uint8_t*> levels;        // device pointers, level 0 = leaf hashes
    std::vector<uint64_t> sizes;
    uint8_t** d_levels = nullptr;        // the same table on the device
    uint64_t* d_sizes = nullptr;
};

struct afc_merkle {
    afc_ctx* ctx = nullptr;
    uint64_t size = 0;
    uint8_t* d_frontier = nullptr;   // 64 x 32
    uint8_t* d_root = nullptr;       // 32
    DevBuf lv[2], leaves, off;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;        // last device-side work enqueued on a caller's stream (orders _dev calls vs host calls)
    std::mutex mu;
};

namespace {

#define CK(expr)                                                           \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) { set_err(ctx, _e, #expr); return AFC_ECUDA; } \
    } while (0)

void set_err(afc_ctx* ctx, cudaError_t e, const char* what) {
    if (!ctx) return;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->last_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + what;
}

// Per-call launch log: counts launches; when profiling is on every launch claims one event pair from the ctx pool.
struct CallLog {
    afc_ctx* ctx;
    launch::LaunchLog lg;
    explicit CallLog(afc_ctx* c, int = 0) : ctx(c) {
        if (ctx->profile) { lg.profile = true; lg.pool = ctx->prof_recs.data(); lg.cap = (int)ctx->prof_recs.size(); lg.next = &ctx->prof_next; }
    }
    ~CallLog() { ctx->launches += lg.n; }
    operator launch::LaunchLog*() { return &lg; }
};


// ---- issuer-key cache management ------------------------------------------------------------------------------
void kc_free_call_buffers(launch::KeyCache& k) {
    void* ps[] = {k.bslots, k.rep, k.kid, k.cnt, k.dlist, k.cand, k.perm, k.cold, k.pts};
    for (void* p : ps) if (p) cudaFree(p);
    k.bslots = k.rep = k.kid = k.cnt = k.dlist = k.cand = k.perm = k.cold = nullptr; k.pts = nullptr;
}
void kc_free(afc_ctx* ctx) {
    launch::KeyCache& k = ctx->kc;
    void* ps[] = {k.slots, k.cpks, k.valid, k.tabs, k.stamp, k.free_list, k.bucket, k.state, k.build_list, k.bases};
    for (void* p : ps) if (p) cudaFree(p);
    kc_free_call_buffers(k);
    if (k.side) cudaStreamDestroy(k.side);
    if (k.side2) cudaStreamDestroy(k.side2);
    cudaEvent_t evs[] = {k.ev_fork, k.ev_join, k.ev_plan, k.ev_rows, k.ev_chain[0], k.ev_chain[1], k.ev_chain[2], k.ev_chain[3]};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    k = launch::KeyCache{};
    ctx->kc_ready = false; ctx->kc_call_cap = 0;
}
// Makes the cache usable for a call of n items (allocating / growing as needed).  Returns nullptr when disabled or when
// memory cannot be had — callers then run the generic kernel.  Caller holds kc_mu.
const launch::KeyCache* kc_prepare(afc_ctx* ctx, uint32_t n) {
    if (ctx->kc_max_keys == 0 || n < 64) return nullptr;
    launch::KeyCache& k = ctx->kc;
    auto dmalloc = [](auto** p, size_t bytes) { return cudaMalloc((void**)p, bytes ? bytes : 1) == cudaSuccess; };
    if (!ctx->kc_ready) {
        uint32_t cap = 1; while (cap < 4 * (uint64_t)ctx->kc_max_keys) cap <<= 1;
        k.slot_mask = cap - 1; k.max_keys = ctx->kc_max_keys;
        const size_t mk = k.max_keys;
        int prio_least = 0, prio_greatest = 0;
        bool ok = dmalloc(&k.slots, (size_t)cap * 4) && dmalloc(&k.cpks, mk * 32) && dmalloc(&k.valid, mk) &&
                  dmalloc(&k.tabs, launch::ed_key_table_bytes(k.max_keys)) && dmalloc(&k.stamp, mk * 4) && dmalloc(&k.free_list, mk * 4) &&
                  dmalloc(&k.bucket, (mk + 1) * 4) && dmalloc(&k.state, launch::KS_WORDS * 4) && dmalloc(&k.build_list, mk * 4) &&
                  dmalloc(&k.bases, launch::ed_key_bases_bytes(k.max_keys)) &&
                  cudaMemset(k.slots, 0xff, (size_t)cap * 4) == cudaSuccess && cudaMemset(k.state, 0, launch::KS_WORDS * 4) == cudaSuccess &&
                  cudaMemset(k.valid, 0, mk) == cudaSuccess && cudaMemset(k.stamp, 0, mk * 4) == cudaSuccess &&
                  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&k.side, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&k.side2, cudaStreamNonBlocking, prio_greatest) == cudaSuccess &&
                  (ctx->kc_event || cudaEventCreateWithFlags(&ctx->kc_event, cudaEventDisableTiming) == cudaSuccess);
        cudaEvent_t* evs[] = {&k.ev_fork, &k.ev_join, &k.ev_plan, &k.ev_rows, &k.ev_chain[0], &k.ev_chain[1], &k.ev_chain[2], &k.ev_chain[3]};
        for (cudaEvent_t* e : evs) ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaStreamSynchronize(cudaStreamLegacy) == cudaSuccess;        // the memsets above vs the non-blocking side streams
        if (!ok) { cudaGetLastError(); kc_free(ctx); ctx->kc_max_keys = 0; return nullptr; }     // no room: stay generic
        ctx->kc_ready = true;
    }
    if (n > ctx->kc_call_cap) {
        // earlier calls may still be using the per-call buffers on their streams: wait for the last cache user before freeing
        if (ctx->kc_call_cap) { cudaEventSynchronize(ctx->kc_event); }
        kc_free_call_buffers(k);
        const uint64_t want = (uint64_t)n + n / 4;
        uint64_t cap = 1; while (cap < 2 * want) cap <<= 1;
        if (cap > 0x80000000ull) { ctx->kc_call_cap = 0; return nullptr; }                    // beyond the 32-bit tables: generic kernel
        bool ok = dmalloc(&k.bslots, (size_t)cap * 4) && dmalloc(&k.rep, want * 4) && dmalloc(&k.kid, want * 4) && dmalloc(&k.cnt, want * 4) &&
                  dmalloc(&k.dlist, want * 4) && dmalloc(&k.cand, want * 4) && dmalloc(&k.perm, want * 4) && dmalloc(&k.cold, want * 4) &&
                  (launch::ed_verify_pts_bytes(want) == 0 || dmalloc(&k.pts, launch::ed_verify_pts_bytes(want)));
        if (!ok) { cudaGetLastError(); kc_free_call_buffers(k); ctx->kc_call_cap = 0; return nullptr; }
        k.bmask = (uint32_t)(cap - 1); ctx->kc_call_cap = (uint32_t)(want > 0xffffffffull ? 0xffffffffu : want);
    }
    return &k;
}
// verify through the cache when possible; serialises cache users across streams with an event
cudaError_t verify_with_cache(afc_ctx* ctx, const uint8_t* pks, const uint8_t* sigs, const uint8_t* msgs, const uint64_t* off, uint32_t n,
                              uint8_t* ok, uint32_t* scratch_k, cudaStream_t st, launch::LaunchLog* lg) {
    std::lock_guard<std::mutex> g(ctx->kc_mu);
    const launch::KeyCache* kc = kc_prepare(ctx, n);
    if (kc) { cudaError_t e = cudaStreamWaitEvent(st, ctx->kc_event, 0); if (e != cudaSuccess) return e; }
    cudaError_t e = launch::ed_verify_batch(ctx->comb, pks, sigs, msgs, off, n, ok, scratch_k, kc, st, lg);
    if (kc && e == cudaSuccess) e = cudaEventRecord(ctx->kc_event, st);
    return e;
}

struct LaneGuard {
    afc_ctx* ctx; int idx;
    LaneGuard(afc_ctx* c) : ctx(c), idx(-1) {
        std::unique_lock<std::mutex> lk(ctx->mu);
        for (;;) {
            for (int i = 0; i < kLanes; i++) if (!ctx->lanes[i].busy) { idx = i; break; }
            if (idx >= 0) break;
            ctx->cv.wait(lk);
        }
        ctx->lanes[idx].busy = true;
    }
    ~LaneGuard() {
        { std::lock_guard<std::mutex> g(ctx->mu); ctx->lanes[idx].busy = false; }
        ctx->cv.notify_one();
    }
    Lane& lane() { return ctx->lanes[idx]; }
};

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// H2D of [src, src+bytes) into dst on `st`; pageable sources go through the slot's pinned bounce buffer
// (offset `*bounce_used` within it, advanced) so the copy is truly asynchronous.
cudaError_t h2d(Slot& sl, void* dst, const void* src, size_t bytes, bool pinned, size_t* bounce_used) {
    if (bytes == 0) return cudaSuccess;
    if (pinned) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sl.stream);
    memcpy(sl.h_in.p + *bounce_used, src, bytes);
    cudaError_t e = cudaMemcpyAsync(dst, sl.h_in.p + *bounce_used, bytes, cudaMemcpyHostToDevice, sl.stream);
    *bounce_used += (bytes + 255) & ~(size_t)255;
    return e;
}

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

enum Op { OP_SHA256, OP_HMAC, OP_VERIFY, OP_SIGN, OP_SIGN_EXP, OP_LEAF, OP_VERIFY_KEYED };

struct BatchArgs {
    Op op;
    const uint8_t* msgs; const uint64_t* off; uint32_t n;
    const uint8_t* a = nullptr; size_t a_item = 0;     // per-item fixed-size input (pks / seeds)
    const uint8_t* b = nullptr; size_t b_item = 0;     // second per-item input (sigs)
    const uint8_t* keys = nullptr; const uint32_t* koff = nullptr;   // HMAC keys
    const uint8_t* d_expanded = nullptr; const uint32_t* key_index = nullptr;  // sign-expanded / keyed verify
    const afc_keyset* keyset = nullptr;
    uint8_t* out; size_t out_item;
};

// Generic chunked pipeline for host-buffer calls: a ring of kSlots chunks in flight so the H2D of later chunks overlaps the kernels of earlier ones.
// offsets must be non-decreasing: a caller's slip would otherwise turn into an out-of-bounds copy from its own buffer
template <class T>
bool monotone(const T* off, uint32_t n) {
    T bad = 0;
    for (uint32_t i = 0; i < n; i++) bad |= (T)(off[i + 1] < off[i]);
    return bad == 0;
}
// On ANY exit of a host-buffer call — error paths included — nothing may still be reading the caller's buffers or writing its
// results: drain every stream of the lane before it is handed to the next caller.
struct LaneDrain {
    Lane& lane; bool armed = true;
    explicit LaneDrain(Lane& l) : lane(l) {}
    ~LaneDrain() { if (armed) for (int w = 0; w < kSlots; w++) cudaStreamSynchronize(lane.slot[w].stream); }
};

int run_host_batch(afc_ctx* ctx, const BatchArgs& A) {
    if (A.n == 0) return AFC_OK;
    if (A.koff && !monotone(A.koff, A.n)) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    LaneGuard lg(ctx);
    Lane& lane = lg.lane();
    LaneDrain drain_on_exit(lane);
    const bool pin_msgs = is_pinned(A.msgs), pin_off = is_pinned(A.off), pin_out = is_pinned(A.out);
    const bool pin_a = A.a ? is_pinned(A.a) : true, pin_b = A.b ? is_pinned(A.b) : true;
    const bool pin_keys = A.keys ? is_pinned(A.keys) : true, pin_koff = A.koff ? is_pinned(A.koff) : true;
    const bool pin_ki = A.key_index ? is_pinned(A.key_index) : true;
    CallLog lc(ctx, 12 * (int)(A.n / kChunkItems + 8) + 8);
    // AFC_CHUNK_ITEMS / AFC_MIN_CHUNK_ITEMS: tuning knobs for experiments
    static const uint32_t max_chunk = [] { const char* e = getenv("AFC_CHUNK_ITEMS"); uint32_t v = e ? (uint32_t)strtoul(e, nullptr, 10) : 0; return v ? v : kChunkItems; }();
    static const uint32_t min_chunk = [] { const char* e = getenv("AFC_MIN_CHUNK_ITEMS"); uint32_t v = e ? (uint32_t)strtoul(e, nullptr, 10) : 0; return v ? v : kMinChunkItems; }();
    uint32_t i0 = 0;
    int which = 0;
    struct Pending { uint32_t i0, cnt; bool active; } pend[kSlots] = {};
    auto drain = [&](int w) -> int {
        if (!pend[w].active) return AFC_OK;
        Slot& sl = lane.slot[w];
        CK(cudaStreamSynchronize(sl.stream));
        if (!pin_out) memcpy(A.out + (size_t)pend[w].i0 * A.out_item, sl.h_out.p, (size_t)pend[w].cnt * A.out_item);
        if (A.op == OP_SIGN && !pin_a && sl.h_in.p) memset(sl.h_in.p, 0, sl.h_in.cap);         // nor in the pinned bounce buffer
        pend[w].active = false;
        return AFC_OK;
    };
    while (i0 < A.n) {
        uint32_t i1 = i0;
        uint64_t base = A.off[i0];
        // chunks taper towards the end of the call (each takes at most half of what is left, down to kMinChunkItems): what
        // remains to be done after the last H2D copy lands — the part of a transfer-bound call that nothing overlaps — is small
        const uint32_t left = A.n - i0;
        uint32_t limit = left <= min_chunk ? left : (left + 1) / 2;
        if (limit < min_chunk) limit = min_chunk;
        if (limit > max_chunk) limit = max_chunk;
        // the walk that sizes the chunk also checks that the offsets never run backwards (no separate pass over n offsets)
        while (i1 < A.n && (i1 - i0) < limit) {
            if (A.off[i1 + 1] < A.off[i1]) return AFC_EINVAL;
            if (A.off[i1 + 1] - base > kChunkBytes && i1 != i0) break;
            i1++;
        }
        uint32_t cnt = i1 - i0;
        uint64_t mbytes = A.off[i1] - base;
        Slot& sl = lane.slot[which];
        int rc = drain(which);
        if (rc != AFC_OK) return rc;
        // size buffers
        CK(sl.msgs.reserve(mbytes + 32));
        CK(sl.off.reserve((size_t)(cnt + 1) * 8));
        CK(sl.out.reserve((size_t)cnt * A.out_item + 16));
        if (A.a) CK(sl.a.reserve((size_t)cnt * A.a_item + 16));
        if (A.b) CK(sl.b.reserve((size_t)cnt * A.b_item + 16));
        if (A.op == OP_VERIFY) CK(sl.k.reserve((size_t)cnt * 32));
        if (A.op == OP_VERIFY_KEYED) CK(sl.k.reserve((size_t)cnt * 32 + launch::ed_keyed_scratch_bytes(A.keyset->n_keys, cnt)));
        uint32_t kbase = 0; size_t kbytes = 0;
        if (A.op == OP_HMAC) {
            kbase = A.koff[i0]; kbytes = A.koff[i1] - kbase;
            CK(sl.a.reserve(kbytes + 16));
            CK(sl.koff.reserve((size_t)(cnt + 1) * 4));
        }
        if ((A.op == OP_SIGN_EXP || A.op == OP_VERIFY_KEYED) && A.key_index) CK(sl.koff.reserve((size_t)cnt * 4));
        size_t bounce = 0;
        if (!pin_msgs) bounce += (mbytes + 255) & ~(size_t)255;
        if (!pin_off) bounce += ((size_t)(cnt + 1) * 8 + 255) & ~(size_t)255;
        if (A.a && !pin_a) bounce += ((size_t)cnt * A.a_item + 255) & ~(size_t)255;
        if (A.b && !pin_b) bounce += ((size_t)cnt * A.b_item + 255) & ~(size_t)255;
        if (A.op == OP_HMAC) { if (!pin_keys) bounce += (kbytes + 255) & ~(size_t)255; if (!pin_koff) bounce += ((size_t)(cnt + 1) * 4 + 255) & ~(size_t)255; }
        if ((A.op == OP_SIGN_EXP || A.op == OP_VERIFY_KEYED) && A.key_index && !pin_ki) bounce += ((size_t)cnt * 4 + 255) & ~(size_t)255;
        CK(sl.h_in.reserve(bounce + 256, ctx));
        if (!pin_out) CK(sl.h_out.reserve((size_t)cnt * A.out_item, ctx));
        size_t used = 0;
        // keep the device copy of the messages congruent mod 16 with the absolute offsets so that the kernels see the
        // same alignment whatever the chunking (d_base + off[i] == sl.msgs.p + pad + off[i] - base)
        size_t pad = (size_t)(base & 15);
        CK(h2d(sl, sl.msgs.p + pad, A.msgs + base, mbytes, pin_msgs, &used));
        CK(h2d(sl, sl.off.p, A.off + i0, (size_t)(cnt + 1) * 8, pin_off, &used));
        if (A.a) CK(h2d(sl, sl.a.p, A.a + (size_t)i0 * A.a_item, (size_t)cnt * A.a_item, pin_a, &used));
        if (A.b) CK(h2d(sl, sl.b.p, A.b + (size_t)i0 * A.b_item, (size_t)cnt * A.b_item, pin_b, &used));
        const uint8_t* d_base = sl.msgs.p + pad - base;     // absolute offsets stay valid
        const uint64_t* d_off = (const uint64_t*)sl.off.p;
        cudaError_t e = cudaSuccess;
        switch (A.op) {
        case OP_SHA256: e = launch::sha256_batch(d_base, d_off, cnt, sl.out.p, sl.stream, lc); break;
        case OP_LEAF: e = launch::merkle_leaf_hashes(d_base, d_off, cnt, sl.out.p, sl.stream, lc); break;
        case OP_HMAC: {
            CK(h2d(sl, sl.a.p + (kbase & 3), A.keys + kbase, kbytes, pin_keys, &used));
            CK(h2d(sl, sl.koff.p, A.koff + i0, (size_t)(cnt + 1) * 4, pin_koff, &used));
            e = launch::hmac_sha256_batch(sl.a.p + (kbase & 3) - kbase, (const uint32_t*)sl.koff.p, d_base, d_off, cnt, sl.out.p, sl.stream, lc);
            break;
        }
        case OP_VERIFY:
            e = verify_with_cache(ctx, sl.a.p, sl.b.p, d_base, d_off, cnt, sl.out.p, (uint32_t*)sl.k.p, sl.stream, lc);
            break;
        case OP_VERIFY_KEYED: {
            CK(h2d(sl, sl.koff.p, A.key_index + i0, (size_t)cnt * 4, pin_ki, &used));
            e = launch::ed_verify_keyed_batch(ctx->comb, A.keyset->d_tabs, A.keyset->d_valid, A.keyset->d_pks, A.keyset->n_keys,
                                              (const uint32_t*)sl.koff.p, sl.b.p, d_base, d_off, cnt, sl.out.p, (uint32_t*)sl.k.p,
                                              (uint32_t*)(sl.k.p + (size_t)cnt * 32), sl.stream, lc);
            break;
        }
        case OP_SIGN: e = launch::ed_sign_batch(ctx->comb, afc_internal_sign_table(ctx), sl.a.p, d_base, d_off, cnt, sl.out.p, sl.stream, lc); break;
        case OP_SIGN_EXP: {
            const uint32_t* d_ki = nullptr;
            if (A.key_index) { CK(h2d(sl, sl.koff.p, A.key_index + i0, (size_t)cnt * 4, pin_ki, &used)); d_ki = (const uint32_t*)sl.koff.p; }
            e = launch::ed_sign_expanded_batch(ctx->comb, afc_internal_sign_table(ctx), A.d_expanded + (A.key_index ? 0 : (size_t)i0 * 96), 0xffffffffu, d_ki, d_base, d_off, cnt, sl.out.p, sl.stream, lc);
            break;
        }
        }
        CK(e);
        if (A.op == OP_SIGN) CK(cudaMemsetAsync(sl.a.p, 0, (size_t)cnt * A.a_item, sl.stream));   // seeds do not stay behind in the staging slot
        CK(cudaMemcpyAsync(pin_out ? A.out + (size_t)i0 * A.out_item : sl.h_out.p, sl.out.p, (size_t)cnt * A.out_item,
                           cudaMemcpyDeviceToHost, sl.stream));
        pend[which] = {i0, cnt, true};
        which = (which + 1) % kSlots;
        i0 = i1;
    }
    for (int w = 0; w < kSlots; w++) {      // oldest first
        int rc = drain((which + w) % kSlots);
        if (rc != AFC_OK) return rc;
    }
    drain_on_exit.armed = false;            // every slot was drained above
    return AFC_OK;
}

}  // namespace

// Binds the CALLING THREAD to the CPUs next to ctx's GPU for the lifetime of the scope: pages allocated (and pinned) meanwhile
// land on that GPU's NUMA node.  One process driving 8 contexts (the Go control plane) otherwise stages half of its batches
// through the far socket: measured in round 1 as 37 instead of 55 GB/s per GPU at 8 GPUs.
struct NumaScope {
    cpu_set_t old; bool active = false;
    explicit NumaScope(afc_ctx* ctx) {
        static const bool off = [] { const char* e = getenv("AFC_NUMA_BIND"); return e && atoi(e) == 0; }();
        if (!ctx || !ctx->has_local_cpus || off) return;
        if (sched_getaffinity(0, sizeof old, &old) != 0) return;
        cpu_set_t want; CPU_AND(&want, &old, &ctx->local_cpus);
        if (CPU_COUNT(&want) == 0 || CPU_EQUAL(&want, &old)) return;
        active = sched_setaffinity(0, sizeof want, &want) == 0;
    }
    ~NumaScope() { if (active) sched_setaffinity(0, sizeof old, &old); }
};
void* afc_internal_pinned_alloc(afc_ctx* ctx, size_t bytes) {
    NumaScope ns(ctx);
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
static void numa_discover(afc_ctx* ctx) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, ctx->device) != cudaSuccess) { cudaGetLastError(); return; }
    for (char* c = bus; *c; c++) if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');
    std::string base = std::string("/sys/bus/pci/devices/") + bus;
    if (FILE* f = fopen((base + "/numa_node").c_str(), "r")) { if (fscanf(f, "%d", &ctx->numa_node) != 1) ctx->numa_node = -1; fclose(f); }
    FILE* f = fopen((base + "/local_cpulist").c_str(), "r");
    if (!f) return;
    char line[4096] = {0};
    if (fgets(line, sizeof line, f)) {
        CPU_ZERO(&ctx->local_cpus);
        int n = 0;
        for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &ctx->local_cpus); n++; } }
            else if (sscanf(tok, "%d", &a) == 1 && a < CPU_SETSIZE) { CPU_SET(a, &ctx->local_cpus); n++; }
        }
        ctx->has_local_cpus = n > 0;
    }
    fclose(f);
}

int afc_internal_device(afc_ctx* ctx) { return ctx->device; }
const void* afc_internal_comb(afc_ctx* ctx) { return ctx->comb; }
const void* afc_internal_sign_table(afc_ctx* ctx) { return ctx->sign_ct ? ctx->ct16 : nullptr; }
void afc_internal_add_launches(afc_ctx* ctx, unsigned long long n) { ctx->launches += n; }

extern "C" {

const char* afc_version(void) { return "afcrypto-b200 0.1.0 (sm_100a)"; }

const char* afc_strerror(int rc) {
    switch (rc) {
    case AFC_OK: return "ok";
    case AFC_EINVAL: return "invalid argument";
    case AFC_ECUDA: return "CUDA failure (see afc_last_cuda_error)";
    case AFC_ENOMEM: return "out of memory";
    case AFC_ENCCL: return "NCCL failure or NCCL unavailable";
    case AFC_ESTATE: return "invalid state";
    default: return "unknown error";
    }
}

const char* afc_last_cuda_error(afc_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }

int afc_init(int device, afc_ctx** out) {
    if (!out) return AFC_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return AFC_ECUDA; }
    if (device < 0 || device >= ndev) return AFC_EINVAL;
    afc_ctx* ctx = new (std::nothrow) afc_ctx();
    if (!ctx) return AFC_ENOMEM;
    ctx->device = device;
    if (const char* e = getenv("AFC_KEYCACHE_MAX_KEYS")) ctx->kc_max_keys = (uint32_t)strtoul(e, nullptr, 10);
    auto fail = [&](cudaError_t e, const char* what) { set_err(ctx, e, what); fprintf(stderr, "afc_init: %s\n", ctx->last_error.c_str()); afc_destroy(ctx); return AFC_ECUDA; };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    numa_discover(ctx);
    {   // per-call scratch comes from the stream-ordered pool: keep freed blocks across synchronisations instead of returning them to
        // the driver (the default threshold of 0 re-maps the 32 MB of a 1 M-credential call after every sync)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    for (int l = 0; l < kLanes; l++)
        for (int s = 0; s < kSlots; s++)
            if ((e = cudaStreamCreateWithFlags(&ctx->lanes[l].slot[s].stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaMalloc(&ctx->comb, launch::ed_tables_bytes())) != cudaSuccess) return fail(e, "cudaMalloc(tables)");
    cudaStream_t s0 = ctx->lanes[0].slot[0].stream;
    if ((e = cudaMalloc(&ctx->ct16, launch::ed_ct_table_bytes())) != cudaSuccess) return fail(e, "cudaMalloc(ct table)");
    if (const char* ev = getenv("AFC_SIGN_CT")) ctx->sign_ct = atoi(ev) != 0;
    {   // the launch log must be gone before fail() can destroy the context it points into
        CallLog lc(ctx);
        e = launch::ed_build_tables(ctx->comb, s0, lc);
        if (e == cudaSuccess) e = launch::ed_build_ct_table(ctx->comb, ctx->ct16, s0, lc);
    }
    if (e != cudaSuccess) return fail(e, "ed_build_tables");
    if ((e = cudaStreamSynchronize(s0)) != cudaSuccess) return fail(e, "ed_build_tables sync");
    *out = ctx;
    return AFC_OK;
}

void afc_destroy(afc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    afc_comm_destroy(ctx);
    for (int l = 0; l < kLanes; l++)
        for (int s = 0; s < kSlots; s++) {
            Slot& sl = ctx->lanes[l].slot[s];
            if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
            sl.msgs.release(); sl.off.release(); sl.a.release(); sl.b.release(); sl.k.release(); sl.out.release(); sl.koff.release();
            sl.h_in.release(); sl.h_out.release();
        }
    kc_free(ctx);
    if (ctx->kc_event) cudaEventDestroy(ctx->kc_event);
    if (ctx->comb) cudaFree(ctx->comb);
    if (ctx->ct16) cudaFree(ctx->ct16);
    for (auto& r : ctx->prof_recs) { if (r.e0) cudaEventDestroy(r.e0); if (r.e1) cudaEventDestroy(r.e1); }
    delete ctx;
}

int afc_device_info(afc_ctx* ctx, int* sm_count, int* clock_khz, uint64_t* mem_bytes) {
    if (!ctx) return AFC_EINVAL;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (clock_khz) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, ctx->device); *clock_khz = v; }
    if (mem_bytes) *mem_bytes = ctx->prop.totalGlobalMem;
    return AFC_OK;
}

int afc_sign_configure(afc_ctx* ctx, int constant_time) {
    if (!ctx) return AFC_EINVAL;
    ctx->sign_ct = constant_time != 0;
    return AFC_OK;
}
int afc_sign_mode(afc_ctx* ctx) { return ctx ? (ctx->sign_ct ? 1 : 0) : AFC_EINVAL; }

uint64_t afc_launch_count(afc_ctx* ctx) { return ctx ? (uint64_t)ctx->launches.load() : 0; }

void* afc_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void afc_free_pinned(void* p) { if (p) cudaFreeHost(p); }
void* afc_alloc_pinned_for(afc_ctx* ctx, size_t bytes) { return ctx ? afc_internal_pinned_alloc(ctx, bytes) : nullptr; }
int afc_numa_info(afc_ctx* ctx, int* numa_node, int* local_cpus) {
    if (!ctx) return AFC_EINVAL;
    if (numa_node) *numa_node = ctx->numa_node;
    if (local_cpus) *local_cpus = ctx->has_local_cpus ? CPU_COUNT(&ctx->local_cpus) : 0;
    return AFC_OK;
}

// ---------------------------------------------------------------------------------- host-buffer batch calls
int afc_sha256_batch(afc_ctx* ctx, const uint8_t* msgs, const uint64_t* offsets, uint32_t n, uint8_t* out32) {
    if (!ctx || !offsets || (!out32 && n) || (!msgs && n && offsets[n] != offsets[0])) return AFC_EINVAL;
    BatchArgs A{}; A.op = OP_SHA256; A.msgs = msgs; A.off = offsets; A.n = n; A.out = out32; A.out_item = 32;
    return run_host_batch(ctx, A);
}
int afc_hmac_sha256_batch(afc_ctx* ctx, const uint8_t* keys, const uint32_t* key_off, const uint8_t* msgs,
                          const uint64_t* msg_off, uint32_t n, uint8_t* out32) {
    if (!ctx || !key_off || !msg_off || (!out32 && n)) return AFC_EINVAL;
    BatchArgs A{}; A.op = OP_HMAC; A.msgs = msgs; A.off = msg_off; A.n = n; A.keys = keys; A.koff = key_off; A.out = out32; A.out_item = 32;
    return run_host_batch(ctx, A);
}
int afc_ed25519_verify_batch(afc_ctx* ctx, const uint8_t* pks, const uint8_t* sigs, const uint8_t* msgs,
                             const uint64_t* msg_off, uint32_t n, uint8_t* ok) {
    if (!ctx || !msg_off || (n && (!pks || !sigs || !ok))) return AFC_EINVAL;
    BatchArgs A{}; A.op = OP_VERIFY; A.msgs = msgs; A.off = msg_off; A.n = n; A.a = pks; A.a_item = 32; A.b = sigs; A.b_item = 64; A.out = ok; A.out_item = 1;
    return run_host_batch(ctx, A);
}
int afc_ed25519_sign_batch(afc_ctx* ctx, const uint8_t* seeds, const uint8_t* msgs, const uint64_t* msg_off,
                           uint32_t n, uint8_t* sigs) {
    if (!ctx || !msg_off || (n && (!seeds || !sigs))) return AFC_EINVAL;
    BatchArgs A{}; A.op = OP_SIGN; A.msgs = msgs; A.off = msg_off; A.n = n; A.a = seeds; A.a_item = 32; A.out = sigs; A.out_item = 64;
    return run_host_batch(ctx, A);
}

// ---------------------------------------------------------------------------------- H2: streaming SHA-256
int afc_sha256_stream_init(uint8_t* states, uint32_t n) {
    if (!states && n) return AFC_EINVAL;
    static const uint32_t iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    for (uint32_t i = 0; i < n; i++) {
        uint8_t* s = states + (size_t)AFC_SHA256_STATE_BYTES * i;
        memset(s, 0, AFC_SHA256_STATE_BYTES);
        s[0] = 's'; s[1] = 'h'; s[2] = 'a'; s[3] = 3;
        for (int w = 0; w < 8; w++) { s[4 + 4 * w] = (uint8_t)(iv[w] >> 24); s[5 + 4 * w] = (uint8_t)(iv[w] >> 16); s[6 + 4 * w] = (uint8_t)(iv[w] >> 8); s[7 + 4 * w] = (uint8_t)iv[w]; }
    }
    return AFC_OK;
}
int afc_sha256_update_batch(afc_ctx* ctx, uint8_t* states, const uint8_t* chunks, const uint64_t* chunk_off, uint32_t n, const uint8_t* final_flags,
                            uint8_t* out32) {
    if (!ctx || !chunk_off || (n && !states)) return AFC_EINVAL;
    if (n == 0) return AFC_OK;
    bool any_final = false;
    for (uint32_t i = 0; i < n; i++) {
        if (chunk_off[i + 1] < chunk_off[i]) return AFC_EINVAL;
        const bool fin = final_flags && final_flags[i];
        any_final |= fin;
        if (!fin && ((chunk_off[i + 1] - chunk_off[i]) & 63)) return AFC_EINVAL;      // only a stream's last chunk may be ragged
    }
    if (any_final && !out32) return AFC_EINVAL;
    const uint64_t base = chunk_off[0], bytes = chunk_off[n] - base;
    if (bytes && !chunks) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    LaneGuard lg(ctx);
    Slot& sl = lg.lane().slot[0];
    LaneDrain drain_on_exit(lg.lane());
    const size_t pad = (size_t)(base & 15);
    CK(sl.msgs.reserve(bytes + 32)); CK(sl.off.reserve((size_t)(n + 1) * 8)); CK(sl.a.reserve((size_t)n * AFC_SHA256_STATE_BYTES));
    CK(sl.b.reserve((size_t)n * 2 + 16)); CK(sl.out.reserve((size_t)n * 32 + 16));
    if (bytes) CK(cudaMemcpyAsync(sl.msgs.p + pad, chunks + base, bytes, cudaMemcpyHostToDevice, sl.stream));
    CK(cudaMemcpyAsync(sl.off.p, chunk_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, sl.stream));
    CK(cudaMemcpyAsync(sl.a.p, states, (size_t)n * AFC_SHA256_STATE_BYTES, cudaMemcpyHostToDevice, sl.stream));
    uint8_t* d_fin = nullptr;
    if (final_flags) { d_fin = sl.b.p; CK(cudaMemcpyAsync(d_fin, final_flags, n, cudaMemcpyHostToDevice, sl.stream)); }
    uint8_t* d_status = sl.b.p + n;
    {
        CallLog lc(ctx);
        CK(launch::sha256_update(sl.a.p, sl.msgs.p + pad - base, (const uint64_t*)sl.off.p, n, d_fin, sl.out.p, d_status, sl.stream, lc));
    }
    std::vector<uint8_t> status(n), dig(any_final ? (size_t)n * 32 : 0);
    CK(cudaMemcpyAsync(states, sl.a.p, (size_t)n * AFC_SHA256_STATE_BYTES, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaMemcpyAsync(status.data(), d_status, n, cudaMemcpyDeviceToHost, sl.stream));
    if (any_final) CK(cudaMemcpyAsync(dig.data(), sl.out.p, (size_t)n * 32, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaStreamSynchronize(sl.stream));
    drain_on_exit.armed = false;
    for (uint32_t i = 0; i < n; i++) {
        if (!status[i]) return AFC_EINVAL;                                             // a state that is not ours (or not on a block boundary)
        if (final_flags && final_flags[i]) memcpy(out32 + 32ull * i, dig.data() + 32ull * i, 32);
    }
    return AFC_OK;
}
int afc_sha256_update_batch_dev(afc_ctx* ctx, uint8_t* d_states, const uint8_t* d_chunks, const uint64_t* d_chunk_off, uint32_t n,
                                const uint8_t* d_final_flags, uint8_t* d_out32, uint8_t* d_status, void* stream) {
    if (!ctx) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CallLog lc(ctx);
    CK(launch::sha256_update(d_states, d_chunks, d_chunk_off, n, d_final_flags, d_out32, d_status, (cudaStream_t)stream, lc));
    return AFC_OK;
}

static int expand_common(afc_ctx* ctx, const uint8_t* seeds, uint32_t n, uint8_t* out, size_t item, bool expanded) {
    if (!ctx || (n && (!seeds || !out))) return AFC_EINVAL;
    if (n == 0) return AFC_OK;
    CK(cudaSetDevice(ctx->device));
    LaneGuard lg(ctx);
    Slot& sl = lg.lane().slot[0];
    CK(sl.a.reserve((size_t)n * 32)); CK(sl.out.reserve((size_t)n * item));
    CK(cudaMemcpyAsync(sl.a.p, seeds, (size_t)n * 32, cudaMemcpyHostToDevice, sl.stream));
    CallLog lc(ctx);
    CK(launch::ed_expand_batch(ctx->comb, afc_internal_sign_table(ctx), sl.a.p, n, expanded ? sl.out.p : nullptr, expanded ? nullptr : sl.out.p, sl.stream, lc));
    CK(cudaMemcpyAsync(out, sl.out.p, (size_t)n * item, cudaMemcpyDeviceToHost, sl.stream));
    CK(cudaMemsetAsync(sl.a.p, 0, (size_t)n * 32, sl.stream));                       // seeds and expanded private keys are wiped from the slot
    if (expanded) CK(cudaMemsetAsync(sl.out.p, 0, (size_t)n * item, sl.stream));
    CK(cudaStreamSynchronize(sl.stream));
    return AFC_OK;
}
int afc_ed25519_pubkey_batch(afc_ctx* ctx, const uint8_t* seeds, uint32_t n, uint8_t* pks) { return expand_common(ctx, seeds, n, pks, 32, false); }
int afc_ed25519_expand_batch(afc_ctx* ctx, const uint8_t* seeds, uint32_t n, uint8_t* expanded96) { return expand_common(ctx, seeds, n, expanded96, 96, true); }

int afc_ed25519_sign_expanded_batch(afc_ctx* ctx, const uint8_t* expanded96, const uint32_t* key_index, uint32_t n_keys,
                                    const uint8_t* msgs, const uint64_t* msg_off, uint32_t n, uint8_t* sigs) {
    if (!ctx || !msg_off || (n && (!expanded96 || !sigs))) return AFC_EINVAL;
    if (!key_index && n_keys < n) return AFC_EINVAL;
    if (key_index) for (uint32_t i = 0; i < n; i++) if (key_index[i] >= n_keys) return AFC_EINVAL;
    if (n == 0) return AFC_OK;
    CK(cudaSetDevice(ctx->device));
    uint8_t* d_keys = nullptr;
    CK(cudaMalloc(&d_keys, (size_t)n_keys * 96));
    cudaError_t e = cudaMemcpy(d_keys, expanded96, (size_t)n_keys * 96, cudaMemcpyHostToDevice);
    int rc = AFC_ECUDA;
    if (e == cudaSuccess) {
        BatchArgs A{}; A.op = OP_SIGN_EXP; A.msgs = msgs; A.off = msg_off; A.n = n; A.d_expanded = d_keys; A.key_index = key_index; A.out = sigs; A.out_item = 64;
        rc = run_host_batch(ctx, A);
    } else set_err(ctx, e, "cudaMemcpy(expanded keys)");
    cudaMemset(d_keys, 0, (size_t)n_keys * 96);       // expanded private keys do not outlive the call on the device
    cudaFree(d_keys);
    return rc;
}

// ---------------------------------------------------------------------------------- device-pointer variants
#define DEV_PROLOGUE()                                   \
    if (!ctx) return AFC_EINVAL;                         \
    CK(cudaSetDevice(ctx->device));                      \
    CallLog lc(ctx);                                     \
    cudaStream_t st = (cudaStream_t)stream;
#define DEV_EPILOGUE() return AFC_OK;

int afc_sha256_batch_dev(afc_ctx* ctx, const uint8_t* d_msgs, const uint64_t* d_offsets, uint32_t n, uint8_t* d_out32, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_out32)) return AFC_EINVAL;
    CK(launch::sha256_batch(d_msgs, d_offsets, n, d_out32, st, lc));
    DEV_EPILOGUE();
}
int afc_hmac_sha256_batch_dev(afc_ctx* ctx, const uint8_t* d_keys, const uint32_t* d_key_off, const uint8_t* d_msgs,
                              const uint64_t* d_msg_off, uint32_t n, uint8_t* d_out32, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_out32)) return AFC_EINVAL;
    CK(launch::hmac_sha256_batch(d_keys, d_key_off, d_msgs, d_msg_off, n, d_out32, st, lc));
    DEV_EPILOGUE();
}
int afc_ed25519_verify_batch_dev(afc_ctx* ctx, const uint8_t* d_pks, const uint8_t* d_sigs, const uint8_t* d_msgs,
                                 const uint64_t* d_msg_off, uint32_t n, uint8_t* d_ok, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_pks) || !aligned16(d_sigs)) return AFC_EINVAL;
    // k scratch: per-call allocation from the stream-ordered pool (freed in stream order)
    uint32_t* d_k = nullptr;
    if (n) CK(cudaMallocAsync((void**)&d_k, (size_t)n * 32, st));
    cudaError_t e = verify_with_cache(ctx, d_pks, d_sigs, d_msgs, d_msg_off, n, d_ok, d_k, st, lc);
    if (n) cudaFreeAsync(d_k, st);
    CK(e);
    DEV_EPILOGUE();
}
int afc_ed25519_sign_batch_dev(afc_ctx* ctx, const uint8_t* d_seeds, const uint8_t* d_msgs, const uint64_t* d_msg_off,
                               uint32_t n, uint8_t* d_sigs, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_seeds) || !aligned16(d_sigs)) return AFC_EINVAL;
    CK(launch::ed_sign_batch(ctx->comb, afc_internal_sign_table(ctx), d_seeds, d_msgs, d_msg_off, n, d_sigs, st, lc));
    DEV_EPILOGUE();
}
int afc_ed25519_pubkey_batch_dev(afc_ctx* ctx, const uint8_t* d_seeds, uint32_t n, uint8_t* d_pks, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_seeds) || !aligned16(d_pks)) return AFC_EINVAL;
    CK(launch::ed_expand_batch(ctx->comb, afc_internal_sign_table(ctx), d_seeds, n, nullptr, d_pks, st, lc));
    DEV_EPILOGUE();
}
int afc_ed25519_expand_batch_dev(afc_ctx* ctx, const uint8_t* d_seeds, uint32_t n, uint8_t* d_expanded96, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_seeds) || !aligned16(d_expanded96)) return AFC_EINVAL;
    CK(launch::ed_expand_batch(ctx->comb, afc_internal_sign_table(ctx), d_seeds, n, d_expanded96, nullptr, st, lc));
    DEV_EPILOGUE();
}
int afc_ed25519_sign_expanded_batch_dev(afc_ctx* ctx, const uint8_t* d_expanded96, const uint32_t* d_key_index,
                                        const uint8_t* d_msgs, const uint64_t* d_msg_off, uint32_t n, uint8_t* d_sigs, void* stream) {
    return afc_ed25519_sign_expanded_keys_batch_dev(ctx, d_expanded96, 0xffffffffu, d_key_index, d_msgs, d_msg_off, n, d_sigs, stream);
}
int afc_ed25519_sign_expanded_keys_batch_dev(afc_ctx* ctx, const uint8_t* d_expanded96, uint32_t n_keys, const uint32_t* d_key_index,
                                             const uint8_t* d_msgs, const uint64_t* d_msg_off, uint32_t n, uint8_t* d_sigs, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_expanded96) || !aligned16(d_sigs) || n_keys == 0) return AFC_EINVAL;
    CK(launch::ed_sign_expanded_batch(ctx->comb, afc_internal_sign_table(ctx), d_expanded96, n_keys, d_key_index, d_msgs, d_msg_off, n, d_sigs, st, lc));
    DEV_EPILOGUE();
}
int afc_merkle_leaf_hashes_dev(afc_ctx* ctx, const uint8_t* d_leaves, const uint64_t* d_leaf_off, uint32_t n, uint8_t* d_out32, void* stream) {
    DEV_PROLOGUE();
    if (!aligned16(d_out32)) return AFC_EINVAL;
    CK(launch::merkle_leaf_hashes(d_leaves, d_leaf_off, n, d_out32, st, lc));
    DEV_EPILOGUE();
}


int afc_b64url_encode_fixed_dev(afc_ctx* ctx, const uint8_t* d_in, uint32_t item_bytes, uint32_t n, uint8_t* d_out, void* stream) {
    DEV_PROLOGUE();
    if (n && item_bytes && (!d_in || !d_out)) return AFC_EINVAL;
    CK(launch::b64url_encode(d_in, item_bytes, n, d_out, st, lc));
    DEV_EPILOGUE();
}
int afc_hex_encode_fixed_dev(afc_ctx* ctx, const uint8_t* d_in, uint32_t item_bytes, uint32_t n, uint8_t* d_out, void* stream) {
    DEV_PROLOGUE();
    if (n && item_bytes && (!d_in || !d_out)) return AFC_EINVAL;
    CK(launch::hex_encode(d_in, (uint64_t)item_bytes * n, d_out, st, lc));
    DEV_EPILOGUE();
}


int afc_json_fill_sizes_dev(afc_ctx* ctx, const uint8_t* d_segs, const uint32_t* d_seg_off, const uint8_t* d_kinds, uint32_t n_fields,
                            const uint8_t* d_fields, const uint64_t* d_field_off, uint32_t n, uint64_t* d_out_off, uint64_t* total_bytes, void* stream) {
    DEV_PROLOGUE();
    if (!d_seg_off || !d_out_off || (n_fields && !d_kinds) || (n && n_fields && !d_field_off)) return AFC_EINVAL;
    uint64_t* d_scratch = nullptr;
    CK(cudaMallocAsync((void**)&d_scratch, launch::json_scan_scratch_bytes(n), st));
    cudaError_t e = launch::json_fill_sizes(d_segs, d_seg_off, d_kinds, n_fields, d_fields, d_field_off, n, d_out_off, d_scratch, st, lc);
    cudaFreeAsync(d_scratch, st);
    CK(e);
    if (total_bytes) {
        CK(cudaMemcpyAsync(total_bytes, d_out_off + n, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    DEV_EPILOGUE();
}
int afc_json_fill_dev(afc_ctx* ctx, const uint8_t* d_segs, const uint32_t* d_seg_off, const uint8_t* d_kinds, uint32_t n_fields,
                      const uint8_t* d_fields, const uint64_t* d_field_off, uint32_t n, const uint64_t* d_out_off, uint8_t* d_out, void* stream) {
    DEV_PROLOGUE();
    if (!d_seg_off || !d_out_off || (n && !d_out) || (n_fields && !d_kinds) || (n && n_fields && !d_field_off)) return AFC_EINVAL;
    CK(launch::json_fill(d_segs, d_seg_off, d_kinds, n_fields, d_fields, d_field_off, n, d_out_off, d_out, st, lc));
    DEV_EPILOGUE();
}


int afc_keycache_configure(afc_ctx* ctx, uint32_t max_keys) {
    if (!ctx) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->kc_mu);
    CK(cudaDeviceSynchronize());
    kc_free(ctx);
    ctx->kc_max_keys = max_keys;
    return AFC_OK;
}
int afc_keycache_clear(afc_ctx* ctx, void* stream) {
    if (!ctx) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->kc_mu);
    if (!ctx->kc_ready) return AFC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaStreamWaitEvent(st, ctx->kc_event, 0));
    {
        CallLog lc(ctx);
        CK(launch::ed_keycache_clear(ctx->kc, st, lc));
    }
    CK(cudaEventRecord(ctx->kc_event, st));
    return AFC_OK;
}
static int kc_read_state(afc_ctx* ctx, uint32_t* st) {
    memset(st, 0, launch::KS_WORDS * 4);
    if (ctx->kc_ready) { CK(cudaDeviceSynchronize()); CK(cudaMemcpy(st, ctx->kc.state, launch::KS_WORDS * 4, cudaMemcpyDeviceToHost)); }
    return AFC_OK;
}
int afc_keycache_info(afc_ctx* ctx, uint32_t* max_keys, uint32_t* cached_keys, uint32_t* last_mode) {
    if (!ctx) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->kc_mu);
    uint32_t st[launch::KS_WORDS];
    int rc = kc_read_state(ctx, st);
    if (rc != AFC_OK) return rc;
    if (max_keys) *max_keys = ctx->kc_max_keys;
    if (cached_keys) *cached_keys = st[launch::KS_HIGH] - st[launch::KS_NFREE];
    if (last_mode) *last_mode = st[launch::KS_NHOT] ? 1u : 0u;
    return AFC_OK;
}
int afc_keycache_stats(afc_ctx* ctx, afc_keycache_stats_t* out) {
    if (!ctx || !out) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->kc_mu);
    uint32_t st[launch::KS_WORDS];
    int rc = kc_read_state(ctx, st);
    if (rc != AFC_OK) return rc;
    out->max_keys = ctx->kc_max_keys; out->cached_keys = st[launch::KS_HIGH] - st[launch::KS_NFREE];
    out->last_hot = st[launch::KS_NHOT]; out->last_cold = st[launch::KS_NCOLD]; out->last_distinct = st[launch::KS_DISTINCT];
    out->last_built = st[launch::KS_NBUILD]; out->last_evicted = st[launch::KS_EVICTED];
    out->total_built = st[launch::KS_TOTAL_BUILT]; out->total_evicted = st[launch::KS_TOTAL_EVICTED]; out->calls = st[launch::KS_CALLS];
    return AFC_OK;
}

// ---------------------------------------------------------------------------------- keyed verification
int afc_keyset_new(afc_ctx* ctx, const uint8_t* pks, uint32_t n_keys, afc_keyset** out) {
    if (!ctx || !out || !pks || n_keys == 0) return AFC_EINVAL;
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    afc_keyset* ks = new (std::nothrow) afc_keyset();
    if (!ks) return AFC_ENOMEM;
    ks->ctx = ctx; ks->n_keys = n_keys; ks->tab_bytes = launch::ed_key_table_bytes(n_keys);
    cudaError_t e = cudaMalloc((void**)&ks->d_pks, (size_t)n_keys * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ks->d_valid, n_keys);
    if (e == cudaSuccess) e = cudaMalloc(&ks->d_tabs, ks->tab_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(ks->d_pks, pks, (size_t)n_keys * 32, cudaMemcpyHostToDevice);
    void* d_bases = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&d_bases, launch::ed_key_bases_bytes(n_keys));
    if (e == cudaSuccess) {
        CallLog lc(ctx);
        e = launch::ed_build_key_tables(ks->d_pks, n_keys, ks->d_tabs, ks->d_valid, d_bases, 0, lc);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (d_bases) cudaFree(d_bases);
    if (e != cudaSuccess) {
        set_err(ctx, e, "afc_keyset_new");
        afc_keyset_free(ks);
        return e == cudaErrorMemoryAllocation ? AFC_ENOMEM : AFC_ECUDA;
    }
    *out = ks;
    return AFC_OK;
}
void afc_keyset_free(afc_keyset* ks) {
    if (!ks) return;
    cudaSetDevice(ks->ctx->device);
    if (ks->d_pks) cudaFree(ks->d_pks);
    if (ks->d_valid) cudaFree(ks->d_valid);
    if (ks->d_tabs) cudaFree(ks->d_tabs);
    delete ks;
}
int afc_keyset_info(afc_keyset* ks, uint32_t* n_keys, uint64_t* table_bytes) {
    if (!ks) return AFC_EINVAL;
    if (n_keys) *n_keys = ks->n_keys;
    if (table_bytes) *table_bytes = ks->tab_bytes;
    return AFC_OK;
}
int afc_ed25519_verify_keyed_batch(afc_ctx* ctx, afc_keyset* ks, const uint32_t* key_index, const uint8_t* sigs,
                                   const uint8_t* msgs, const uint64_t* msg_off, uint32_t n, uint8_t* ok) {
    if (!ctx || !ks || ks->ctx != ctx || !msg_off || (n && (!key_index || !sigs || !ok))) return AFC_EINVAL;
    BatchArgs A{}; A.op = OP_VERIFY_KEYED; A.msgs = msgs; A.off = msg_off; A.n = n; A.b = sigs; A.b_item = 64; A.key_index = key_index;
    A.keyset = ks; A.out = ok; A.out_item = 1;
    return run_host_batch(ctx, A);
}
int afc_ed25519_verify_keyed_batch_dev(afc_ctx* ctx, afc_keyset* ks, const uint32_t* d_key_index, const uint8_t* d_sigs,
                                       const uint8_t* d_msgs, const uint64_t* d_msg_off, uint32_t n, uint8_t* d_ok, void* stream) {
    if (!ctx || !ks || ks->ctx != ctx) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (!aligned16(d_sigs)) return AFC_EINVAL;
    CallLog lc(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* d_k = nullptr;
    if (n) CK(cudaMallocAsync((void**)&d_k, (size_t)n * 32 + launch::ed_keyed_scratch_bytes(ks->n_keys, n), st));
    cudaError_t e = launch::ed_verify_keyed_batch(ctx->comb, ks->d_tabs, ks->d_valid, ks->d_pks, ks->n_keys, d_key_index, d_sigs, d_msgs,
                                                  d_msg_off, n, d_ok, d_k, n ? d_k + (size_t)n * 8 : nullptr, st, lc);
    if (n) cudaFreeAsync(d_k, st);
    CK(e);
    return AFC_OK;
}

// ---------------------------------------------------------------------------------- Merkle log
int afc_merkle_new(afc_ctx* ctx, afc_merkle** out) {
    if (!ctx || !out) return AFC_EINVAL;
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    afc_merkle* m = new (std::nothrow) afc_merkle();
    if (!m) return AFC_ENOMEM;
    m->ctx = ctx;
    cudaError_t e = cudaMalloc((void**)&m->d_frontier, 64 * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_root, 32);
    if (e == cudaSuccess) e = cudaMemset(m->d_frontier, 0, 64 * 32);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);     // the log's own stream is non-blocking: it would not wait for that memset
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(m->ev, m->stream);
    if (e != cudaSuccess) { set_err(ctx, e, "afc_merkle_new"); afc_merkle_free(m); return AFC_ECUDA; }
    *out = m;
    return AFC_OK;
}
void afc_merkle_free(afc_merkle* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    if (m->ev) cudaEventDestroy(m->ev);
    if (m->d_frontier) cudaFree(m->d_frontier);
    if (m->d_root) cudaFree(m->d_root);
    m->lv[0].release(); m->lv[1].release(); m->leaves.release(); m->off.release();
    delete m;
}

// Level schedule (host decides structure from (size, n) only; every hash runs on the device).
// d_h: n node hashes occupying leaf positions [size, size+n).  May alias m->lv[1] but not m->lv[0].
static int merkle_append_hashes_locked(afc_merkle* m, const uint8_t* d_h, uint32_t n, cudaStream_t st) {
    afc_ctx* ctx = m->ctx;
    if (n == 0) return AFC_OK;
    if (!aligned16(d_h)) return AFC_EINVAL;
    CallLog lc(ctx, 72);
    uint64_t s = m->size, e = m->size + n;
    const uint8_t* cur = d_h;
    int h = 0, pp = 0;
    // d_h aliasing lv[1]: first output must go to lv[0]
    CK(m->lv[0].reserve(((size_t)n / 2 + 2) * 32));
    while (e > s) {
        int left = (int)(s & 1);
        uint64_t s2 = s + left;
        int right = (int)(e & 1);
        uint64_t e2 = e - right;
        uint64_t npairs = (e2 - s2) / 2;
        DevBuf& ob = m->lv[pp];
        CK(ob.reserve(((size_t)npairs + 2) * 32));
        CK(launch::merkle_level(cur, ob.p, npairs, left, right, m->d_frontier, h, st, lc));
        cur = ob.p;
        s = s2 / 2 - left;
        e = e2 / 2;
        pp ^= 1;
        h++;
        if (h > 63) return AFC_ESTATE;
    }
    m->size += n;
    return AFC_OK;
}

int afc_merkle_append_hashes_dev(afc_merkle* m, const uint8_t* d_hashes32, uint32_t n, void* stream) {
    if (!m) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    // the log state and level buffers are shared across calls: order this stream after whatever touched them last
    CK(cudaStreamWaitEvent((cudaStream_t)stream, m->ev, 0));
    int rc = merkle_append_hashes_locked(m, d_hashes32, n, (cudaStream_t)stream);
    CK(cudaEventRecord(m->ev, (cudaStream_t)stream));
    return rc;
}
int afc_merkle_append_dev(afc_merkle* m, const uint8_t* d_leaves, const uint64_t* d_leaf_off, uint32_t n, void* stream) {
    if (!m) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    if (n == 0) return AFC_OK;
    CallLog lc(ctx);
    CK(cudaStreamWaitEvent((cudaStream_t)stream, m->ev, 0));
    int rc = AFC_ECUDA;
    cudaError_t e = m->lv[1].reserve((size_t)n * 32);
    if (e == cudaSuccess) e = launch::merkle_leaf_hashes(d_leaves, d_leaf_off, n, m->lv[1].p, (cudaStream_t)stream, lc);
    if (e == cudaSuccess) rc = merkle_append_hashes_locked(m, m->lv[1].p, n, (cudaStream_t)stream);
    else set_err(ctx, e, "afc_merkle_append_dev");
    cudaEventRecord(m->ev, (cudaStream_t)stream);      // whatever was enqueued must order the next user of the log, also after a failure
    return rc;
}
int afc_merkle_root_dev(afc_merkle* m, uint8_t* d_root32, void* stream) {
    if (!m || !d_root32) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    CallLog lc(ctx);
    if (!aligned16(d_root32)) return AFC_EINVAL;
    CK(cudaStreamWaitEvent((cudaStream_t)stream, m->ev, 0));
    CK(launch::merkle_root(m->d_frontier, m->size, d_root32, (cudaStream_t)stream, lc));
    CK(cudaEventRecord(m->ev, (cudaStream_t)stream));
    return AFC_OK;
}
static int merkle_root_host_locked(afc_merkle* m, uint8_t root32[32], uint64_t* tree_size) {
    afc_ctx* ctx = m->ctx;
    CallLog lc(ctx);
    CK(cudaStreamWaitEvent(m->stream, m->ev, 0));
    CK(launch::merkle_root(m->d_frontier, m->size, m->d_root, m->stream, lc));
    if (root32) CK(cudaMemcpyAsync(root32, m->d_root, 32, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    if (tree_size) *tree_size = m->size;
    return AFC_OK;
}
int afc_merkle_root(afc_merkle* m, uint8_t root32[32], uint64_t* tree_size) {
    if (!m) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    return merkle_root_host_locked(m, root32, tree_size);
}
int afc_merkle_append(afc_merkle* m, const uint8_t* leaves, const uint64_t* leaf_off, uint32_t n, uint8_t root32[32], uint64_t* tree_size) {
    if (!m || !leaf_off) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    CK(cudaStreamWaitEvent(m->stream, m->ev, 0));
    if (n) {
        if (!monotone(leaf_off, n)) return AFC_EINVAL;
        uint64_t base = leaf_off[0], bytes = leaf_off[n] - base;
        if (bytes && !leaves) return AFC_EINVAL;
        size_t pad = (size_t)(base & 15);
        CK(m->leaves.reserve(bytes + 32)); CK(m->off.reserve((size_t)(n + 1) * 8)); CK(m->lv[1].reserve((size_t)n * 32));
        if (bytes) CK(cudaMemcpyAsync(m->leaves.p + pad, leaves + base, bytes, cudaMemcpyHostToDevice, m->stream));
        CK(cudaMemcpyAsync(m->off.p, leaf_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, m->stream));
        CallLog lc(ctx);
        CK(launch::merkle_leaf_hashes(m->leaves.p + pad - base, (const uint64_t*)m->off.p, n, m->lv[1].p, m->stream, lc));
            int rc = merkle_append_hashes_locked(m, m->lv[1].p, n, m->stream);
        if (rc != AFC_OK) return rc;
    }
    return merkle_root_host_locked(m, root32, tree_size);
}
int afc_merkle_append_hashes(afc_merkle* m, const uint8_t* hashes32, uint32_t n, uint8_t root32[32], uint64_t* tree_size) {
    if (!m || (n && !hashes32)) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    CK(cudaStreamWaitEvent(m->stream, m->ev, 0));
    if (n) {
        CK(m->lv[1].reserve((size_t)n * 32));
        CK(cudaMemcpyAsync(m->lv[1].p, hashes32, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
        int rc = merkle_append_hashes_locked(m, m->lv[1].p, n, m->stream);
        if (rc != AFC_OK) return rc;
    }
    return merkle_root_host_locked(m, root32, tree_size);
}
int afc_merkle_save(afc_merkle* m, uint8_t* state) {
    if (!m || !state) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    CK(cudaStreamWaitEvent(m->stream, m->ev, 0));
    CK(cudaStreamSynchronize(m->stream));
    memcpy(state, &m->size, 8);
    CK(cudaMemcpy(state + 8, m->d_frontier, 64 * 32, cudaMemcpyDeviceToHost));
    for (int h = 0; h < 64; h++) if (!((m->size >> h) & 1)) memset(state + 8 + 32 * h, 0, 32);
    return AFC_OK;
}
int afc_merkle_load(afc_merkle* m, const uint8_t* state) {
    if (!m || !state) return AFC_EINVAL;
    afc_ctx* ctx = m->ctx;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(m->mu);
    CK(cudaStreamWaitEvent(m->stream, m->ev, 0));
    CK(cudaStreamSynchronize(m->stream));
    memcpy(&m->size, state, 8);
    CK(cudaMemcpy(m->d_frontier, state + 8, 64 * 32, cudaMemcpyHostToDevice));
    return AFC_OK;
}


// ---------------------------------------------------------------------------------- materialised tree + audit proofs
static int tree_build_common(afc_ctx* ctx, const uint8_t* src, bool src_is_device, uint64_t n, afc_merkle_tree** out) {
    if (!ctx || !out || (n && !src)) return AFC_EINVAL;
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    afc_merkle_tree* t = new (std::nothrow) afc_merkle_tree();
    if (!t) return AFC_ENOMEM;
    t->ctx = ctx; t->n = n;
    CallLog lc(ctx, 72);
    cudaError_t e = cudaSuccess;
    uint64_t cur = n;
    while (e == cudaSuccess) {
        uint8_t* p = nullptr;
        e = cudaMalloc((void**)&p, (size_t)(cur ? cur : 1) * 32);
        if (e != cudaSuccess) break;
        t->levels.push_back(p); t->sizes.push_back(cur);
        size_t h = t->levels.size() - 1;
        if (h == 0) {
            if (n) e = cudaMemcpy(p, src, (size_t)n * 32, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
        } else {
            e = launch::merkle_level_promote(t->levels[h - 1], t->sizes[h - 1], p, 0, lc);
        }
        if (cur <= 1) break;
        cur = (cur + 1) / 2;
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&t->d_levels, t->levels.size() * sizeof(uint8_t*));
    if (e == cudaSuccess) e = cudaMalloc((void**)&t->d_sizes, t->sizes.size() * 8);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_levels, t->levels.data(), t->levels.size() * sizeof(uint8_t*), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_sizes, t->sizes.data(), t->sizes.size() * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_err(ctx, e, "afc_merkle_tree_build"); afc_merkle_tree_free(t); return e == cudaErrorMemoryAllocation ? AFC_ENOMEM : AFC_ECUDA; }
    *out = t;
    return AFC_OK;
}
int afc_merkle_tree_build(afc_ctx* ctx, const uint8_t* leaf_hashes32, uint64_t n, afc_merkle_tree** out) { return tree_build_common(ctx, leaf_hashes32, false, n, out); }
int afc_merkle_tree_build_dev(afc_ctx* ctx, const uint8_t* d_leaf_hashes32, uint64_t n, afc_merkle_tree** out) { return tree_build_common(ctx, d_leaf_hashes32, true, n, out); }
void afc_merkle_tree_free(afc_merkle_tree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    for (uint8_t* p : t->levels) if (p) cudaFree(p);
    if (t->d_levels) cudaFree(t->d_levels);
    if (t->d_sizes) cudaFree(t->d_sizes);
    delete t;
}
int afc_merkle_tree_root(afc_merkle_tree* t, uint8_t root32[32], uint64_t* n_leaves, uint32_t* max_proof_nodes) {
    if (!t) return AFC_EINVAL;
    afc_ctx* ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    if (n_leaves) *n_leaves = t->n;
    if (max_proof_nodes) *max_proof_nodes = (uint32_t)(t->levels.size() > 1 ? t->levels.size() - 1 : 1);
    if (root32) {
        if (t->n == 0) {                                   // MTH({}) = SHA-256("") — computed on the device like everything else
            CallLog lc(ctx);
            uint8_t* d = nullptr;
            CK(cudaMalloc((void**)&d, 32));
            cudaError_t e = launch::merkle_root(d, 0, d, 0, lc);
            if (e == cudaSuccess) e = cudaMemcpy(root32, d, 32, cudaMemcpyDeviceToHost);
            cudaFree(d);
            CK(e);
        } else {
            CK(cudaMemcpy(root32, t->levels.back(), 32, cudaMemcpyDeviceToHost));
        }
    }
    return AFC_OK;
}
int afc_merkle_tree_inclusion_proofs(afc_merkle_tree* t, const uint64_t* indices, uint32_t m, uint8_t* proofs, uint32_t* proof_lens) {
    if (!t || (m && (!indices || !proofs || !proof_lens))) return AFC_EINVAL;
    afc_ctx* ctx = t->ctx;
    if (m == 0) return AFC_OK;
    CK(cudaSetDevice(ctx->device));
    for (uint32_t i = 0; i < m; i++) if (indices[i] >= t->n) return AFC_EINVAL;
    const size_t depth = t->levels.size() > 1 ? t->levels.size() - 1 : 1;
    uint64_t* d_idx = nullptr; uint8_t* d_out = nullptr; uint32_t* d_len = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_idx, (size_t)m * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, (size_t)m * depth * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_len, (size_t)m * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_idx, indices, (size_t)m * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_out, 0, (size_t)m * depth * 32);             // slots past a path's length come back as zeros
    if (e == cudaSuccess) {
        CallLog lc(ctx);
        e = launch::merkle_gather_proofs((const uint8_t* const*)t->d_levels, t->d_sizes, (int)t->levels.size(), d_idx, m, d_out, d_len, 0, lc);
    }
    if (e == cudaSuccess) e = cudaMemcpy(proofs, d_out, (size_t)m * depth * 32, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(proof_lens, d_len, (size_t)m * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_idx); cudaFree(d_out); cudaFree(d_len);
    CK(e);
    return AFC_OK;
}
int afc_merkle_verify_inclusion_batch(afc_ctx* ctx, const uint8_t* leaf_hashes32, const uint64_t* indices, uint64_t tree_size,
                                      const uint8_t* proofs, const uint32_t* proof_off, const uint8_t root32[32], uint32_t m, uint8_t* ok) {
    if (!ctx || (m && (!leaf_hashes32 || !indices || !proof_off || !root32 || !ok))) return AFC_EINVAL;
    if (m == 0) return AFC_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t pn = proof_off[m];
    if (pn && !proofs) return AFC_EINVAL;
    uint8_t *d_lh = nullptr, *d_pr = nullptr, *d_root = nullptr, *d_ok = nullptr; uint64_t* d_idx = nullptr; uint32_t* d_po = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_lh, (size_t)m * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_pr, (pn ? pn : 1) * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_root, 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_ok, m);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_idx, (size_t)m * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_po, (size_t)(m + 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_lh, leaf_hashes32, (size_t)m * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && pn) e = cudaMemcpy(d_pr, proofs, pn * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_root, root32, 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_idx, indices, (size_t)m * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_po, proof_off, (size_t)(m + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        CallLog lc(ctx);
        e = launch::merkle_verify_inclusion(d_lh, d_idx, tree_size, d_pr, d_po, d_root, m, d_ok, 0, lc);
    }
    if (e == cudaSuccess) e = cudaMemcpy(ok, d_ok, m, cudaMemcpyDeviceToHost);
    cudaFree(d_lh); cudaFree(d_pr); cudaFree(d_root); cudaFree(d_ok); cudaFree(d_idx); cudaFree(d_po);
    CK(e);
    return AFC_OK;
}

// RFC 6962 §2.1.2 PROOF(first, D[n]) read out of the materialised levels.  Every node of the proof is MTH(D[a:b]) with a
// a multiple of 2^l and b = min(a + 2^l, n): exactly node (level l, index a >> l) of the tree (a lone right node is promoted
// unchanged, which is what the RFC's largest-power-of-two split yields for a ragged right edge).
namespace {
void consistency_nodes(uint64_t m, uint64_t lo, uint64_t hi, bool whole, std::vector<std::pair<int, uint64_t>>& out) {
    const uint64_t n = hi - lo;
    auto node = [&](uint64_t a, uint64_t width_pow2) { int l = 0; while ((1ull << l) < width_pow2) l++; out.emplace_back(l, a >> l); };
    auto pow2_ceil = [](uint64_t w) { uint64_t p = 1; while (p < w) p <<= 1; return p; };
    if (m == n) { if (!whole) node(lo, pow2_ceil(n)); return; }
    uint64_t k = 1; while (k * 2 < n) k *= 2;
    if (m <= k) { consistency_nodes(m, lo, lo + k, whole, out); node(lo + k, k); }
    else { consistency_nodes(m - k, lo + k, hi, false, out); node(lo, k); }
}
}  // namespace
int afc_merkle_tree_consistency_proof(afc_merkle_tree* t, uint64_t first, uint8_t* proof, uint32_t* n_nodes) {
    if (!t || !n_nodes || first == 0 || first > t->n) return AFC_EINVAL;
    afc_ctx* ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    std::vector<std::pair<int, uint64_t>> nodes;
    consistency_nodes(first, 0, t->n, true, nodes);
    if (nodes.size() && !proof) return AFC_EINVAL;
    for (size_t i = 0; i < nodes.size(); i++) {
        const int l = nodes[i].first; const uint64_t j = nodes[i].second;
        if (l >= (int)t->levels.size() || j >= t->sizes[l]) return AFC_ESTATE;
        CK(cudaMemcpy(proof + 32 * i, t->levels[l] + 32 * j, 32, cudaMemcpyDeviceToHost));
    }
    *n_nodes = (uint32_t)nodes.size();
    return AFC_OK;
}
int afc_merkle_verify_consistency_batch(afc_ctx* ctx, const uint64_t* first_sizes, const uint8_t* first_roots32, uint64_t second_size,
                                        const uint8_t second_root32[32], const uint8_t* proofs, const uint32_t* proof_off, uint32_t m, uint8_t* ok) {
    if (!ctx || (m && (!first_sizes || !first_roots32 || !second_root32 || !proof_off || !ok))) return AFC_EINVAL;
    if (m == 0) return AFC_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t pn = proof_off[m];
    if (pn && !proofs) return AFC_EINVAL;
    uint8_t *d_fr = nullptr, *d_pr = nullptr, *d_sr = nullptr, *d_ok = nullptr; uint64_t* d_fs = nullptr; uint32_t* d_po = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_fr, (size_t)m * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_pr, (pn ? pn : 1) * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_sr, 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_ok, m);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_fs, (size_t)m * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_po, (size_t)(m + 1) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_fr, first_roots32, (size_t)m * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && pn) e = cudaMemcpy(d_pr, proofs, pn * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_sr, second_root32, 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_fs, first_sizes, (size_t)m * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_po, proof_off, (size_t)(m + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        CallLog lc(ctx);
        e = launch::merkle_verify_consistency(d_fs, d_fr, second_size, d_sr, d_pr, d_po, m, d_ok, 0, lc);
    }
    if (e == cudaSuccess) e = cudaMemcpy(ok, d_ok, m, cudaMemcpyDeviceToHost);
    cudaFree(d_fr); cudaFree(d_pr); cudaFree(d_sr); cudaFree(d_ok); cudaFree(d_fs); cudaFree(d_po);
    CK(e);
    return AFC_OK;
}

// ---------------------------------------------------------------------------------- NCCL (dlopen, optional)
namespace {
struct nccl_id_t { char internal[128]; };
typedef int (*fn_get_uid)(nccl_id_t*);
typedef int (*fn_init_rank)(void**, int, nccl_id_t, int);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
void* nccl_open() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // reuse the copy torch already loaded, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return h;
}
}  // namespace

int afc_comm_unique_id(uint8_t id128[128]) {
    void* h = nccl_open();
    if (!h) return AFC_ENCCL;
    fn_get_uid f = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
    if (!f) return AFC_ENCCL;
    nccl_id_t id;
    if (f(&id) != 0) return AFC_ENCCL;
    memcpy(id128, id.internal, 128);
    return AFC_OK;
}
int afc_comm_init(afc_ctx* ctx, int nranks, int rank, const uint8_t id128[128]) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return AFC_EINVAL;
    if (ctx->nccl_comm) return AFC_ESTATE;
    CK(cudaSetDevice(ctx->device));
    ctx->nccl_lib = nccl_open();
    if (!ctx->nccl_lib) return AFC_ENCCL;
    fn_init_rank f = (fn_init_rank)dlsym(ctx->nccl_lib, "ncclCommInitRank");
    if (!f) return AFC_ENCCL;
    nccl_id_t id; memcpy(id.internal, id128, 128);
    if (f(&ctx->nccl_comm, nranks, id, rank) != 0) { ctx->nccl_comm = nullptr; return AFC_ENCCL; }
    ctx->nranks = nranks; ctx->rank = rank;
    CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    CK(cudaMalloc((void**)&ctx->d_comm_buf, (size_t)(nranks + 1) * 32));
    return AFC_OK;
}
int afc_comm_allgather_roots(afc_ctx* ctx, const uint8_t local_root32[32], uint8_t* all_roots) {
    if (!ctx || !local_root32 || !all_roots) return AFC_EINVAL;
    if (!ctx->nccl_comm) return AFC_ESTATE;
    CK(cudaSetDevice(ctx->device));
    fn_allgather f = (fn_allgather)dlsym(ctx->nccl_lib, "ncclAllGather");
    if (!f) return AFC_ENCCL;
    uint8_t* send = ctx->d_comm_buf + (size_t)ctx->nranks * 32;
    CK(cudaMemcpyAsync(send, local_root32, 32, cudaMemcpyHostToDevice, ctx->comm_stream));
    if (f(send, ctx->d_comm_buf, 32, /*ncclUint8*/ 1, ctx->nccl_comm, ctx->comm_stream) != 0) return AFC_ENCCL;
    CK(cudaMemcpyAsync(all_roots, ctx->d_comm_buf, (size_t)ctx->nranks * 32, cudaMemcpyDeviceToHost, ctx->comm_stream));
    CK(cudaStreamSynchronize(ctx->comm_stream));
    return AFC_OK;
}
int afc_comm_destroy(afc_ctx* ctx) {
    if (!ctx) return AFC_EINVAL;
    if (ctx->nccl_comm) {
        fn_destroy f = (fn_destroy)dlsym(ctx->nccl_lib, "ncclCommDestroy");
        if (f) f(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
    if (ctx->d_comm_buf) { cudaFree(ctx->d_comm_buf); ctx->d_comm_buf = nullptr; }
    return AFC_OK;
}


// ---------------------------------------------------------------------------------- per-kernel event profiling
int afc_profile_begin(afc_ctx* ctx, int max_launches) {
    if (!ctx || max_launches <= 0) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> g(ctx->mu);
    if (ctx->profile) return AFC_ESTATE;
    size_t old = ctx->prof_recs.size();
    if ((size_t)max_launches > old) {
        ctx->prof_recs.resize(max_launches);
        for (size_t i = old; i < ctx->prof_recs.size(); i++) {
            ctx->prof_recs[i].name = nullptr; ctx->prof_recs[i].done = false;
            if (cudaEventCreate(&ctx->prof_recs[i].e0) != cudaSuccess || cudaEventCreate(&ctx->prof_recs[i].e1) != cudaSuccess) return AFC_ECUDA;
        }
    }
    for (auto& r : ctx->prof_recs) r.done = false;
    __atomic_store_n(&ctx->prof_next, 0, __ATOMIC_RELAXED);
    ctx->profile = true;
    return AFC_OK;
}
// Returns the number of distinct kernel names (entries written to out, at most cap).  Launches beyond the pool given to
// afc_profile_begin were not timed: they are reported in the entry named "(untimed)" (count only) so that a short pool is visible.
int afc_profile_end(afc_ctx* ctx, afc_profile_entry* out, int cap) {
    if (!ctx || (cap > 0 && !out)) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->profile) return AFC_ESTATE;
    ctx->profile = false;
    int n_out = 0;
    const int claimed = __atomic_load_n(&ctx->prof_next, __ATOMIC_RELAXED);
    const int timed = claimed < (int)ctx->prof_recs.size() ? claimed : (int)ctx->prof_recs.size();
    for (int i = 0; i < timed; i++) {
        launch::LaunchRec& r = ctx->prof_recs[i];
        if (!r.done) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
        int k = 0;
        for (; k < n_out; k++) if (strncmp(out[k].name, r.name, sizeof(out[k].name) - 1) == 0) break;
        if (k == n_out) {
            if (n_out >= cap) continue;
            memset(&out[k], 0, sizeof(out[k]));
            strncpy(out[k].name, r.name, sizeof(out[k].name) - 1);
            out[k].min_ms = ms; out[k].max_ms = ms;
            n_out++;
        }
        out[k].count++;
        out[k].total_ms += ms;
        if (ms < out[k].min_ms) out[k].min_ms = ms;
        if (ms > out[k].max_ms) out[k].max_ms = ms;
    }
    if (claimed > timed && n_out < cap) {
        memset(&out[n_out], 0, sizeof(out[n_out]));
        strncpy(out[n_out].name, "(untimed)", sizeof(out[n_out].name) - 1);
        out[n_out].count = (uint32_t)(claimed - timed);
        n_out++;
    }
    return n_out;
}

// ---------------------------------------------------------------------------------- diagnostics
int afc_selftest(afc_ctx* ctx, uint32_t iters) {
    if (!ctx) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    uint32_t* d = nullptr;
    CK(cudaMalloc((void**)&d, 4));
    CK(cudaMemset(d, 0, 4));
    CallLog lc(ctx);
    cudaError_t e = launch::ed_selftest(iters, d, 0, lc);
    uint32_t bad = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&bad, d, 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    CK(e);
    return (int)bad;
}
int afc_microbench(afc_ctx* ctx, int which, uint32_t iters, double* ops_per_s, double* ms_out) {
    if (!ctx || !ops_per_s) return AFC_EINVAL;
    CK(cudaSetDevice(ctx->device));
    uint32_t* d = nullptr;
    CK(cudaMalloc((void**)&d, 4));
    const uint32_t threads = 128;
    const uint32_t blocks = (uint32_t)ctx->prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CallLog lc(ctx);
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 2 && e == cudaSuccess; rep++) {      // first pass warms up
        cudaEventRecord(e0, 0);
        if (which == 3 || which == 4) e = launch::microbench_hash(which, iters, blocks, threads, d, 0, lc);
        else e = launch::microbench_fe(which, iters, blocks, threads, d, 0, lc);
        cudaEventRecord(e1, 0);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    }
    float ms = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    CK(e);
    double per_thread = (which == 3 || which == 4) ? 1.0 : (which >= 20 ? 1.0 : (which >= 10 ? 16.0 : 2.0));   // fe probes: 2 ops / iteration; pipe probes: 16 instr
    *ops_per_s = (double)blocks * threads * iters * per_thread / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return AFC_OK;
}

}  // extern "C"
