// Host-side fork/join helpers for the bulk paths (gm_bulk_load, gm_retain_bulk_load, the first flush after them):
// plain std::thread, no pool — these run a handful of times per process, over millions of items each.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
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

// out = the sorted, duplicate-free union of per-thread lists that are each sorted and duplicate-free.  The key space is cut
// at sampled splitters; the thread of a range pulls its part out of every list (binary searches), sorts and de-duplicates it.
inline void merge_sorted_unique(const std::vector<std::vector<uint64_t>>& lists, unsigned T, std::vector<uint64_t>& out) {
    out.clear();
    std::vector<uint64_t> samples;
    for (const auto& l : lists) { constexpr size_t S = 64; if (!l.empty()) for (size_t k = 0; k < S; ++k) samples.push_back(l[(l.size() * (2 * k + 1)) / (2 * S)]); }
    if (samples.empty()) return;
    T = std::max(1u, T);
    std::sort(samples.begin(), samples.end());
    std::vector<uint64_t> split(T + 1, 0);             // range r = [split[r], split[r + 1]), the last one unbounded above
    for (unsigned r = 1; r < T; ++r) split[r] = samples[(samples.size() * r) / T];
    std::vector<std::vector<uint64_t>> seg(T);
    parallel_threads(T, [&](unsigned r) {
        std::vector<uint64_t> mine;
        for (const auto& l : lists) {
            auto a = r == 0 ? l.begin() : std::lower_bound(l.begin(), l.end(), split[r]);
            auto b = r + 1 == T ? l.end() : std::lower_bound(l.begin(), l.end(), split[r + 1]);
            if (a < b) mine.insert(mine.end(), a, b);
        }
        std::sort(mine.begin(), mine.end());
        mine.erase(std::unique(mine.begin(), mine.end()), mine.end());
        seg[r].swap(mine);
    });
    std::vector<size_t> base(T + 1, 0);
    for (unsigned r = 0; r < T; ++r) base[r + 1] = base[r] + seg[r].size();
    out.resize(base[T]);
    parallel_threads(T, [&](unsigned r) { std::copy(seg[r].begin(), seg[r].end(), out.begin() + base[r]); });
}

}  // namespace gm
