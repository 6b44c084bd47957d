// Batching ingest dispatcher (BASELINE.json configs[4]): turns the reference's one-action-at-a-time
// calls into GPU batches.  The reference issues one VC per HTTP request (internal/handlers/did_handlers.go:192 ->
// internal/services/vc_service.go:138) and signs one webhook per worker goroutine (internal/services/webhook_dispatcher.go:
// 261-320: 4 workers, queue 256); here every "agent action" = one Ed25519 signature over the credential bytes (expanded
// key from the identity cache), one HMAC-SHA256 webhook tag, and one audit leaf (the 64-byte signature) appended to the
// RFC 6962 log.
//
// Producers (any threads) copy an action into the open batch under a mutex; one worker thread flushes when `batch_max`
// actions are pending or the oldest has waited `linger_us`, runs H2D -> sign -> HMAC -> Merkle append -> D2H on its own
// stream and completes the tickets.  Two pinned batch buffers alternate so producers fill one while the GPU works on the
// other (back-pressure when both are full).  afc_ingest_soak is the native open-loop (Poisson) load generator.
#include "../../include/afcrypto.h"
#include "afc_internal.h"
#include "afc_launch.h"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

using namespace afc;
using clk = std::chrono::steady_clock;

namespace {

constexpr uint32_t kHistBuckets = 1u << 16;      // 4 us per bucket -> 262 ms, last bucket = overflow
constexpr uint32_t kHistShift = 2;

struct Batch {
    uint8_t *h_msgs = nullptr, *h_keys = nullptr, *h_bodies = nullptr, *h_sigs = nullptr, *h_tags = nullptr;
    uint64_t *h_moff = nullptr, *h_boff = nullptr;
    uint32_t *h_ki = nullptr, *h_koff = nullptr;
    uint8_t *d_msgs = nullptr, *d_keys = nullptr, *d_bodies = nullptr, *d_sigs = nullptr, *d_tags = nullptr;
    uint64_t *d_moff = nullptr, *d_boff = nullptr;
    uint32_t *d_ki = nullptr, *d_koff = nullptr;
    uint32_t count = 0;
    uint64_t msg_bytes = 0, key_bytes = 0, body_bytes = 0, first_ticket = 0;
    clk::time_point first_ts;
    std::vector<clk::time_point> ts;
};

}  // namespace

struct afc_ingest {
    afc_ctx* ctx = nullptr;
    afc_merkle* log = nullptr;
    uint32_t n_keys = 0, batch_max = 0, linger_us = 0, max_msg = 0, max_key = 0, max_body = 0;
    uint8_t* d_expanded = nullptr;
    uint64_t* d_sigoff = nullptr;          // 0, 64, 128, ...: leaf offsets of the signature array
    cudaStream_t stream = nullptr;
    Batch b[2];
    int open = 0;
    bool busy[2] = {false, false};
    std::mutex mu;
    std::condition_variable cv_work, cv_space, cv_done;
    std::thread worker;
    bool stop = false, force = false;
    uint64_t next_ticket = 0, completed = 0;
    uint32_t ring = 0;
    uint8_t *r_sig = nullptr, *r_tag = nullptr;
    uint64_t batches = 0;
    std::vector<uint32_t> hist;
    uint64_t lat_max_us = 0;
    int last_rc = 0;
};

namespace {

bool alloc_batch(afc_ctx* ctx, Batch& B, uint32_t bm, uint32_t mm, uint32_t mk, uint32_t mb) {
    // staging on the NUMA node of the dispatcher's GPU, whichever thread creates it
    auto pin = [ctx](void* p, size_t n) { return (*(void**)p = afc_internal_pinned_alloc(ctx, n)) != nullptr; };
    auto dev = [](void* p, size_t n) { return cudaMalloc((void**)p, n) == cudaSuccess; };
    size_t n1 = (size_t)bm + 1;
    B.ts.resize(bm);
    return pin(&B.h_msgs, (size_t)bm * mm + 16) && pin(&B.h_moff, n1 * 8) && pin(&B.h_ki, (size_t)bm * 4) && pin(&B.h_keys, (size_t)bm * mk + 16) &&
           pin(&B.h_koff, n1 * 4) && pin(&B.h_bodies, (size_t)bm * mb + 16) && pin(&B.h_boff, n1 * 8) && pin(&B.h_sigs, (size_t)bm * 64) &&
           pin(&B.h_tags, (size_t)bm * 32) && dev(&B.d_msgs, (size_t)bm * mm + 16) && dev(&B.d_moff, n1 * 8) && dev(&B.d_ki, (size_t)bm * 4) &&
           dev(&B.d_keys, (size_t)bm * mk + 16) && dev(&B.d_koff, n1 * 4) && dev(&B.d_bodies, (size_t)bm * mb + 16) && dev(&B.d_boff, n1 * 8) &&
           dev(&B.d_sigs, (size_t)bm * 64) && dev(&B.d_tags, (size_t)bm * 32);
}
void free_batch(Batch& B) {
    void* hp[] = {B.h_msgs, B.h_moff, B.h_ki, B.h_keys, B.h_koff, B.h_bodies, B.h_boff, B.h_sigs, B.h_tags};
    for (void* p : hp) if (p) cudaFreeHost(p);
    void* dp[] = {B.d_msgs, B.d_moff, B.d_ki, B.d_keys, B.d_koff, B.d_bodies, B.d_boff, B.d_sigs, B.d_tags};
    for (void* p : dp) if (p) cudaFree(p);
}

int process(afc_ingest* g, Batch& B) {
    cudaStream_t st = g->stream;
    const uint32_t n = B.count;
    B.h_moff[n] = B.msg_bytes; B.h_koff[n] = (uint32_t)B.key_bytes; B.h_boff[n] = B.body_bytes;
    launch::LaunchLog lg;
    cudaError_t e = cudaSuccess;
    auto cp = [&](void* d, const void* h, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st); };
    cp(B.d_msgs, B.h_msgs, B.msg_bytes); cp(B.d_moff, B.h_moff, (size_t)(n + 1) * 8); cp(B.d_ki, B.h_ki, (size_t)n * 4);
    cp(B.d_keys, B.h_keys, B.key_bytes); cp(B.d_koff, B.h_koff, (size_t)(n + 1) * 4);
    cp(B.d_bodies, B.h_bodies, B.body_bytes); cp(B.d_boff, B.h_boff, (size_t)(n + 1) * 8);
    if (e == cudaSuccess) e = launch::ed_sign_expanded_batch(afc_internal_comb(g->ctx), afc_internal_sign_table(g->ctx), g->d_expanded, g->n_keys, B.d_ki, B.d_msgs, B.d_moff, n, B.d_sigs, st, &lg);
    if (e == cudaSuccess) e = launch::hmac_sha256_batch(B.d_keys, B.d_koff, B.d_bodies, B.d_boff, n, B.d_tags, st, &lg);
    afc_internal_add_launches(g->ctx, lg.n);
    if (e != cudaSuccess) return AFC_ECUDA;
    int rc = afc_merkle_append_dev(g->log, B.d_sigs, g->d_sigoff, n, st);
    if (rc != AFC_OK) return rc;
    e = cudaMemcpyAsync(B.h_sigs, B.d_sigs, (size_t)n * 64, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(B.h_tags, B.d_tags, (size_t)n * 32, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? AFC_OK : AFC_ECUDA;
}

void worker_main(afc_ingest* g) {
    cudaSetDevice(afc_internal_device(g->ctx));
    std::unique_lock<std::mutex> lk(g->mu);
    while (true) {
        Batch& B = g->b[g->open];
        if (B.count == 0) {
            if (g->stop) break;
            g->force = false;
            g->cv_done.notify_all();
            g->cv_work.wait(lk);
            continue;
        }
        auto deadline = B.first_ts + std::chrono::microseconds(g->linger_us);
        if (B.count < g->batch_max && !g->force && !g->stop && clk::now() < deadline) {
            g->cv_work.wait_until(lk, deadline);
            continue;
        }
        int idx = g->open;
        g->busy[idx] = true;
        g->open = idx ^ 1;                      // the other buffer is free: this single worker finished it before
        g->cv_space.notify_all();
        lk.unlock();
        int rc = process(g, g->b[idx]);
        auto done = clk::now();
        lk.lock();
        Batch& D = g->b[idx];
        for (uint32_t i = 0; i < D.count; i++) {
            uint64_t t = D.first_ticket + i;
            memcpy(g->r_sig + (size_t)(t % g->ring) * 64, D.h_sigs + (size_t)i * 64, 64);
            memcpy(g->r_tag + (size_t)(t % g->ring) * 32, D.h_tags + (size_t)i * 32, 32);
            uint64_t us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(done - D.ts[i]).count();
            uint32_t bk = (uint32_t)std::min<uint64_t>(us >> kHistShift, kHistBuckets - 1);
            g->hist[bk]++;
            if (us > g->lat_max_us) g->lat_max_us = us;
        }
        if (rc != AFC_OK) g->last_rc = rc;
        g->completed += D.count;
        g->batches++;
        D.count = 0; D.msg_bytes = D.key_bytes = D.body_bytes = 0;
        g->busy[idx] = false;
        g->cv_done.notify_all();
        g->cv_space.notify_all();
    }
    g->cv_done.notify_all();
}

uint32_t percentile(const std::vector<uint32_t>& h, uint64_t total, double q) {
    if (!total) return 0;
    uint64_t want = (uint64_t)(q * (double)total), acc = 0;
    for (uint32_t i = 0; i < h.size(); i++) { acc += h[i]; if (acc > want) return (i << kHistShift) + (1u << kHistShift) / 2; }
    return (uint32_t)(h.size() << kHistShift);
}

}  // namespace

extern "C" {

int afc_ingest_new(afc_ctx* ctx, const uint8_t* expanded96, uint32_t n_keys, uint32_t batch_max, uint32_t linger_us, uint32_t max_msg,
                   uint32_t max_key, uint32_t max_body, afc_ingest** out) {
    if (!ctx || !out || !expanded96 || !n_keys || !batch_max || !max_msg || !max_body) return AFC_EINVAL;
    *out = nullptr;
    if (cudaSetDevice(afc_internal_device(ctx)) != cudaSuccess) return AFC_ECUDA;
    afc_ingest* g = new (std::nothrow) afc_ingest();
    if (!g) return AFC_ENOMEM;
    g->ctx = ctx; g->n_keys = n_keys; g->batch_max = batch_max; g->linger_us = linger_us;
    g->max_msg = max_msg; g->max_key = max_key ? max_key : 1; g->max_body = max_body;
    g->ring = std::max<uint32_t>(4 * batch_max, 1u << 16);
    g->hist.assign(kHistBuckets, 0);
    bool ok = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc((void**)&g->d_expanded, (size_t)n_keys * 96) == cudaSuccess &&
              cudaMemcpy(g->d_expanded, expanded96, (size_t)n_keys * 96, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMalloc((void**)&g->d_sigoff, ((size_t)batch_max + 1) * 8) == cudaSuccess &&
              (g->r_sig = (uint8_t*)afc_internal_pinned_alloc(ctx, (size_t)g->ring * 64)) != nullptr &&
              (g->r_tag = (uint8_t*)afc_internal_pinned_alloc(ctx, (size_t)g->ring * 32)) != nullptr &&
              alloc_batch(ctx, g->b[0], batch_max, max_msg, g->max_key, max_body) && alloc_batch(ctx, g->b[1], batch_max, max_msg, g->max_key, max_body) &&
              afc_merkle_new(ctx, &g->log) == AFC_OK;
    if (ok) {
        std::vector<uint64_t> so((size_t)batch_max + 1);
        for (size_t i = 0; i < so.size(); i++) so[i] = 64 * i;
        ok = cudaMemcpy(g->d_sigoff, so.data(), so.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) { cudaGetLastError(); afc_ingest_free(g); return AFC_ECUDA; }
    g->worker = std::thread(worker_main, g);
    *out = g;
    return AFC_OK;
}

void afc_ingest_free(afc_ingest* g) {
    if (!g) return;
    if (g->worker.joinable()) {
        { std::lock_guard<std::mutex> lk(g->mu); g->stop = true; }
        g->cv_work.notify_all(); g->cv_space.notify_all();
        g->worker.join();
    }
    cudaSetDevice(afc_internal_device(g->ctx));
    free_batch(g->b[0]); free_batch(g->b[1]);
    if (g->log) afc_merkle_free(g->log);
    if (g->d_expanded) cudaFree(g->d_expanded);
    if (g->d_sigoff) cudaFree(g->d_sigoff);
    if (g->r_sig) cudaFreeHost(g->r_sig);
    if (g->r_tag) cudaFreeHost(g->r_tag);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
}

int afc_ingest_submit(afc_ingest* g, uint32_t key_index, const uint8_t* msg, uint32_t msg_len, const uint8_t* hkey, uint32_t hkey_len,
                      const uint8_t* body, uint32_t body_len, uint64_t* ticket) {
    if (!g || key_index >= g->n_keys || msg_len > g->max_msg || hkey_len > g->max_key || body_len > g->max_body ||
        (msg_len && !msg) || (hkey_len && !hkey) || (body_len && !body)) return AFC_EINVAL;
    auto now = clk::now();
    std::unique_lock<std::mutex> lk(g->mu);
    while (!g->stop && g->b[g->open].count >= g->batch_max) g->cv_space.wait(lk);     // back-pressure: both buffers full
    if (g->stop) return AFC_ESTATE;
    Batch& B = g->b[g->open];
    uint32_t i = B.count;
    if (i == 0) { B.first_ts = now; B.first_ticket = g->next_ticket; }
    B.h_moff[i] = B.msg_bytes; B.h_koff[i] = (uint32_t)B.key_bytes; B.h_boff[i] = B.body_bytes; B.h_ki[i] = key_index;
    if (msg_len) memcpy(B.h_msgs + B.msg_bytes, msg, msg_len);
    if (hkey_len) memcpy(B.h_keys + B.key_bytes, hkey, hkey_len);
    if (body_len) memcpy(B.h_bodies + B.body_bytes, body, body_len);
    B.msg_bytes += msg_len; B.key_bytes += hkey_len; B.body_bytes += body_len;
    B.ts[i] = now;
    B.count = i + 1;
    uint64_t t = g->next_ticket++;
    if (ticket) *ticket = t;
    bool wake = (B.count == 1) || (B.count >= g->batch_max);
    lk.unlock();
    if (wake) g->cv_work.notify_one();
    return AFC_OK;
}

int afc_ingest_wait(afc_ingest* g, uint64_t ticket, uint8_t sig64[64], uint8_t tag32[32]) {
    if (!g) return AFC_EINVAL;
    std::unique_lock<std::mutex> lk(g->mu);
    if (ticket >= g->next_ticket) return AFC_EINVAL;
    while (g->completed <= ticket && !g->stop) g->cv_done.wait(lk);
    if (g->completed <= ticket) return AFC_ESTATE;
    if (g->completed - ticket > g->ring) return AFC_ESTATE;      // result already overwritten
    if (sig64) memcpy(sig64, g->r_sig + (size_t)(ticket % g->ring) * 64, 64);
    if (tag32) memcpy(tag32, g->r_tag + (size_t)(ticket % g->ring) * 32, 32);
    return g->last_rc;
}

int afc_ingest_flush(afc_ingest* g) {
    if (!g) return AFC_EINVAL;
    std::unique_lock<std::mutex> lk(g->mu);
    uint64_t target = g->next_ticket;
    g->force = true;
    g->cv_work.notify_all();
    while (g->completed < target && !g->stop) { g->force = true; g->cv_work.notify_all(); g->cv_done.wait_for(lk, std::chrono::milliseconds(1)); }
    return g->last_rc;
}

int afc_ingest_stats_get(afc_ingest* g, afc_ingest_stats* s) {
    if (!g || !s) return AFC_EINVAL;
    memset(s, 0, sizeof *s);
    {
        std::lock_guard<std::mutex> lk(g->mu);
        s->submitted = g->next_ticket; s->completed = g->completed; s->batches = g->batches;
        s->avg_batch = g->batches ? (double)g->completed / (double)g->batches : 0.0;
        s->p50_us = percentile(g->hist, g->completed, 0.50); s->p99_us = percentile(g->hist, g->completed, 0.99);
        s->max_us = (uint32_t)std::min<uint64_t>(g->lat_max_us, 0xffffffffu);
        s->last_error = g->last_rc;
    }
    return afc_merkle_root(g->log, s->log_root, &s->log_size);
}

// Open-loop Poisson load: `producers` threads, each submitting synthetic actions at rate/producers with exponential gaps.
int afc_ingest_soak(afc_ingest* g, double rate_per_s, double seconds, uint32_t producers, uint32_t msg_len, uint32_t body_len, uint64_t seed,
                    afc_ingest_stats* out, double* achieved_rate, uint64_t* late_submits) {
    if (!g || rate_per_s <= 0 || seconds <= 0 || !producers || msg_len > g->max_msg || body_len > g->max_body || g->max_key < 32) return AFC_EINVAL;
    std::vector<std::thread> th;
    std::vector<uint64_t> late(producers, 0), sent(producers, 0);
    auto t0 = clk::now() + std::chrono::milliseconds(5);
    auto t_end = t0 + std::chrono::duration_cast<clk::duration>(std::chrono::duration<double>(seconds));
    for (uint32_t p = 0; p < producers; p++) {
        th.emplace_back([&, p]() {
            std::mt19937_64 rng(seed * 1315423911ull + p);
            std::exponential_distribution<double> gap(rate_per_s / producers);
            std::vector<uint8_t> msg(msg_len), body(body_len), key(32);
            for (auto& x : msg) x = (uint8_t)rng();
            for (auto& x : body) x = (uint8_t)rng();
            for (auto& x : key) x = (uint8_t)rng();
            auto next = t0 + std::chrono::duration_cast<clk::duration>(std::chrono::duration<double>(gap(rng)));
            while (next < t_end) {
                auto now = clk::now();
                if (now < next) {
                    if (next - now > std::chrono::microseconds(200)) std::this_thread::sleep_until(next - std::chrono::microseconds(100));
                    while (clk::now() < next) { /* spin */ }
                } else if (now - next > std::chrono::milliseconds(1)) {
                    late[p]++;
                }
                uint64_t r = rng();
                if (msg_len >= 8) memcpy(msg.data(), &r, 8);           // every action differs
                if (body_len >= 8) memcpy(body.data(), &r, 8);
                if (afc_ingest_submit(g, (uint32_t)(r % g->n_keys), msg.data(), msg_len, key.data(), 32, body.data(), body_len, nullptr) != AFC_OK) break;
                sent[p]++;
                next += std::chrono::duration_cast<clk::duration>(std::chrono::duration<double>(gap(rng)));
            }
        });
    }
    for (auto& t : th) t.join();
    int rc = afc_ingest_flush(g);
    double el = std::chrono::duration<double>(clk::now() - t0).count();
    uint64_t total = 0, l = 0;
    for (uint32_t p = 0; p < producers; p++) { total += sent[p]; l += late[p]; }
    if (achieved_rate) *achieved_rate = (double)total / el;
    if (late_submits) *late_submits = l;
    if (out) { int r2 = afc_ingest_stats_get(g, out); if (rc == AFC_OK) rc = r2; }
    return rc;
}

}  // extern "C"
