// Concurrency stress of the single-call front end (rmqtt_b200/csrc/batcher.cpp) WITHOUT a device: batcher.cpp only uses the
// public C ABI, so the entry points it calls are stubbed here (gm_match_batch answers every topic with ids derived from the
// topic's bytes, after a short sleep standing in for the device).  Checked: every submitted cookie gets exactly one callback
// with the ids of ITS topic, from several producer threads, for several (max_batch, max_wait_us, dispatchers) settings,
// including an output that outgrows the dispatcher's buffer (GM_ERR_CAPACITY protocol) and a failing batch.
// Built with -fsanitize=thread and -fsanitize=address,undefined by tests/test_sanitizers.py.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpumqtt.h"

static std::atomic<int> g_fail_next{0};
static std::atomic<uint64_t> g_batches{0}, g_capacity_retries{0};

static uint32_t fnv(const char* s, uint32_t n) { uint32_t h = 2166136261u; for (uint32_t i = 0; i < n; ++i) h = (h ^ static_cast<unsigned char>(s[i])) * 16777619u; return h; }
static uint32_t ids_of(uint32_t h) { return h % 7 == 0 ? 300u : h % 5; }        // a few topics produce long lists

extern "C" {
struct gm_engine { int dummy; };
void* gm_host_alloc_near(gm_engine*, uint64_t bytes) { return std::malloc(bytes ? bytes : 1); }
void gm_host_free(void* p) { std::free(p); }
const char* gm_last_error(gm_engine*) { return "stub"; }
// (referenced by the churn probe that lives in the same translation unit; never called here)
int32_t gm_sub_add(gm_engine*, const char*, uint32_t, uint32_t, int32_t*) { return GM_ERR_INTERNAL; }
int32_t gm_sub_remove(gm_engine*, const char*, uint32_t, uint32_t, int32_t*) { return GM_ERR_INTERNAL; }
int32_t gm_flush(gm_engine*) { return GM_ERR_INTERNAL; }
int32_t gm_match_batch(gm_engine*, const char* blob, const uint32_t* offs, uint64_t n, gm_span* spans, uint32_t* ids, uint64_t cap, uint64_t* needed, int32_t* status) {
    g_batches.fetch_add(1);
    std::this_thread::sleep_for(std::chrono::microseconds(50));
    if (g_fail_next.exchange(0)) return GM_ERR_CUDA;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) total += ids_of(fnv(blob + offs[i], offs[i + 1] - offs[i]));
    *needed = total;
    if (total > cap) { g_capacity_retries.fetch_add(1); return GM_ERR_CAPACITY; }
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t len = offs[i + 1] - offs[i], h = fnv(blob + offs[i], len), c = ids_of(h);
        status[i] = (len && blob[offs[i]] == '!') ? GM_ERR_INVALID_TOPIC : GM_OK;
        spans[i] = gm_span{static_cast<uint32_t>(w), c};
        for (uint32_t k = 0; k < c; ++k) ids[w++] = h + k;
    }
    return GM_OK;
}
}

struct Sink {
    std::vector<std::atomic<uint32_t>> seen;
    std::vector<std::string> topics;
    std::atomic<uint64_t> bad{0}, failed{0};
    explicit Sink(size_t n) : seen(n), topics(n) {}
};

static void on_match(void* user, uint64_t cookie, int32_t status, const uint32_t* ids, uint32_t n_ids) {
    Sink* s = static_cast<Sink*>(user);
    s->seen[cookie].fetch_add(1);
    const std::string& t = s->topics[cookie];
    if (status == GM_ERR_CUDA) { s->failed.fetch_add(1); return; }                  // the whole batch failed: reported per topic
    const uint32_t h = fnv(t.data(), static_cast<uint32_t>(t.size()));
    if (!t.empty() && t[0] == '!') { if (status != GM_ERR_INVALID_TOPIC || n_ids != 0) s->bad.fetch_add(1); return; }
    if (status != GM_OK || n_ids != ids_of(h)) { s->bad.fetch_add(1); return; }
    for (uint32_t k = 0; k < n_ids; ++k) if (ids[k] != h + k) { s->bad.fetch_add(1); return; }
}

static int run(uint32_t max_batch, uint32_t wait_us, uint32_t dispatchers, int producers, int per_producer, bool inject_failure) {
    gm_engine eng{0};
    const size_t total = static_cast<size_t>(producers) * per_producer;
    Sink sink(total);
    for (size_t i = 0; i < total; ++i) sink.topics[i] = (i % 97 == 0 ? "!bad/" : "t/") + std::to_string(i * 2654435761u % 100003) + "/x";
    gm_batcher_config cfg{};
    cfg.struct_size = sizeof(cfg); cfg.max_batch = max_batch; cfg.max_wait_us = wait_us; cfg.dispatchers = dispatchers; cfg.on_match = on_match; cfg.user = &sink;
    gm_batcher* b = nullptr;
    if (gm_batcher_create(&eng, &cfg, &b) != GM_OK) return 1;
    if (inject_failure) g_fail_next.store(1);
    std::vector<std::thread> th;
    for (int p = 0; p < producers; ++p)
        th.emplace_back([&, p] {
            for (int i = 0; i < per_producer; ++i) {
                const size_t c = static_cast<size_t>(p) * per_producer + i;
                if (gm_submit(b, sink.topics[c].data(), static_cast<uint32_t>(sink.topics[c].size()), c) != GM_OK) std::abort();
                if (i % 1000 == 999) gm_batcher_drain(b);                              // drains interleaved with other producers' submits
            }
        });
    for (auto& t : th) t.join();
    gm_batcher_drain(b);
    size_t once = 0;
    for (size_t i = 0; i < total; ++i) once += sink.seen[i].load() == 1;
    gm_batcher_destroy(b);
    const bool ok = once == total && sink.bad.load() == 0 && (inject_failure ? sink.failed.load() > 0 : sink.failed.load() == 0);
    printf("batcher max_batch %u wait %u us dispatchers %u producers %d: %zu/%zu delivered once, bad %llu, failed %llu -> %s\n", max_batch, wait_us, dispatchers, producers,
           once, total, (unsigned long long)sink.bad.load(), (unsigned long long)sink.failed.load(), ok ? "ok:" : "FAILED");
    return ok ? 0 : 1;
}

int main() {
    int rc = 0;
    rc |= run(1, 0, 1, 1, 2000, false);
    rc |= run(64, 200, 2, 4, 5000, false);
    rc |= run(4096, 500, 3, 6, 8000, false);          // big batches: the id list outgrows the dispatcher's first buffer (capacity protocol)
    rc |= run(256, 100, 2, 3, 3000, true);            // one failing batch: its topics get the error, the rest are served
    printf("batches %llu, capacity retries %llu\n", (unsigned long long)g_batches.load(), (unsigned long long)g_capacity_retries.load());
    return rc;
}
