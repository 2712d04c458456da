// Host-side fork/join helpers for the bulk paths (gm_bulk_load, gm_retain_bulk_load, the first flush after them):
// plain std::thread, no pool — these run a handful of times per process, over millions of items each.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <thread>
#include <vector>

namespace gm {

// worker threads for a bulk phase: the host's hardware threads, capped (GM_HOST_THREADS overrides; 1 = serial)
inline unsigned host_threads() {
    if (const char* ev = getenv("GM_HOST_THREADS")) { const int v = atoi(ev); if (v >= 1 && v <= 1024) return static_cast<unsigned>(v); }
    static const unsigned hw = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    return hw;
}

// smallest job a bulk phase spreads over the threads (below it the fork/join costs more than it saves); GM_HOST_PAR_MIN
// lowers it so that tests reach the parallel paths with small inputs
inline size_t host_par_min(size_t dflt) {
    if (const char* ev = getenv("GM_HOST_PAR_MIN")) { const long long v = atoll(ev); if (v >= 1) return static_cast<size_t>(v); }
    return dflt;
}

// fn(tid, begin, end) over [0, n) cut into `threads` contiguous chunks (chunk t belongs to tid t: callers keep per-tid state)
template <class F>
inline void parallel_chunks(size_t n, unsigned threads, F&& fn) {
    threads = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(threads, n)));
    if (threads == 1) { fn(0u, size_t(0), n); return; }
    const size_t chunk = (n + threads - 1) / threads;
    std::vector<std::thread> th;
    th.reserve(threads - 1);
    for (unsigned t = 1; t < threads; ++t) {
        const size_t b = std::min(n, t * chunk), e = std::min(n, b + chunk);
        th.emplace_back([&fn, t, b, e] { fn(t, b, e); });
    }
    fn(0u, size_t(0), std::min(n, chunk));
    for (auto& t : th) t.join();
}

// fn(tid) once per thread
template <class F>
inline void parallel_threads(unsigned threads, F&& fn) {
    if (threads <= 1) { fn(0u); return; }
    std::vector<std::thread> th;
    th.reserve(threads - 1);
    for (unsigned t = 1; t < threads; ++t) th.emplace_back([&fn, t] { fn(t); });
    fn(0u);
    for (auto& t : th) t.join();
}

}  // namespace gm
