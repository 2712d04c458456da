// Deterministic synthetic workload generator `iot6`.  Same bytes feed the oracle, the
// GPU engine and the benchmark.  Not part of the matching path; built as libgmworkload.so.
//
//   topic  = reg-%02d/site-%04d/dev-%07d/sen-%d/met-%d/ch-%d       (6 levels)
//   site and device numbers are GLOBAL indices (site = r*S+s, dev = site*D+d) so the level
//   dictionary is realistically large (C3: ~1.05 M distinct level strings).
//
// Every item i draws from its own splitmix64 stream seeded by (seed, i): generation is order-free.
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed, uint64_t i, uint64_t stream) : s(seed ^ (i * 0x9E3779B97F4A7C15ull) ^ (stream * 0xD1B54A32D192ED03ull)) { next(); }
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return static_cast<uint64_t>((static_cast<unsigned __int128>(next()) * n) >> 64); }
};

struct Params {
    uint32_t R, S, D, K, M, F;
    double p_plus, p_hash, p_root_plus;
    uint64_t seed;
};

struct Tp { uint32_t r, site, dev, k, m, f; };

inline uint64_t space(const Params& p) { return 1ull * p.R * p.S * p.D * p.K * p.M * p.F; }

inline Tp decode(const Params& p, uint64_t x) {
    Tp t;
    t.f = x % p.F; x /= p.F;
    t.m = x % p.M; x /= p.M;
    t.k = x % p.K; x /= p.K;
    uint32_t d = x % p.D; x /= p.D;
    uint32_t s = x % p.S; x /= p.S;
    t.r = static_cast<uint32_t>(x);
    t.site = t.r * p.S + s;
    t.dev = t.site * p.D + d;
    return t;
}

inline char* put_num(char* o, uint32_t v, int width) {
    char tmp[12]; int n = 0;
    do { tmp[n++] = '0' + v % 10; v /= 10; } while (v);
    for (int i = n; i < width; ++i) *o++ = '0';
    while (n) *o++ = tmp[--n];
    return o;
}
inline char* put_str(char* o, const char* s) { while (*s) *o++ = *s++; return o; }

// writes level `lv` (0..5) of topic t
inline char* put_level(char* o, const Tp& t, int lv) {
    switch (lv) {
        case 0: o = put_str(o, "reg-"); return put_num(o, t.r, 2);
        case 1: o = put_str(o, "site-"); return put_num(o, t.site, 4);
        case 2: o = put_str(o, "dev-"); return put_num(o, t.dev, 7);
        case 3: o = put_str(o, "sen-"); return put_num(o, t.k, 1);
        case 4: o = put_str(o, "met-"); return put_num(o, t.m, 1);
        default: o = put_str(o, "ch-"); return put_num(o, t.f, 1);
    }
}

// mask bit lv set => level replaced by '+'; depth<6 => truncate to `depth` levels and append '#'
inline char* put_filter(char* o, const Tp& t, uint32_t plus_mask, int depth) {
    for (int lv = 0; lv < depth; ++lv) {
        if (lv) *o++ = '/';
        if (plus_mask >> lv & 1) *o++ = '+'; else o = put_level(o, t, lv);
    }
    if (depth < 6) { *o++ = '/'; *o++ = '#'; }
    return o;
}

// concrete topic index of subscription i (first draw of its stream)
inline uint64_t sub_topic_index(const Params& p, uint64_t i) { Rng g(p.seed, i, 1); return g.below(space(p)); }

inline uint64_t pick_region_restricted(const Params& p, Rng& g, const uint32_t* regions, uint32_t nreg) {
    // uniform over the topics whose region is in `regions`
    uint64_t per_region = space(p) / p.R;
    uint32_t r = regions[g.below(nreg)];
    return 1ull * r * per_region + g.below(per_region);
}

// Feistel permutation over [0, 2^bits) with cycle walking down to [0, n): sampling without replacement.
inline uint64_t permute(uint64_t x, uint64_t n, uint64_t seed) {
    int bits = 1; while ((1ull << bits) < n) ++bits;
    if (bits & 1) ++bits;
    int half = bits / 2; uint64_t mask = (1ull << half) - 1;
    do {
        uint64_t l = x >> half, r = x & mask;
        for (int round = 0; round < 4; ++round) {
            uint64_t z = r + seed + 0x9E3779B97F4A7C15ull * (round + 1);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            uint64_t nl = r, nr = l ^ (z & mask);
            l = nl; r = nr;
        }
        x = (l << half) | r;
    } while (x >= n);
    return x;
}

constexpr int kMaxLen = 64;

}  // namespace

// ---- the big generators on all host threads -------------------------------------------------------------------------
// Items are order-free (own random stream per index): every thread generates a contiguous share into private buffers, the
// shares are stitched together at their byte offsets — the same bytes the serial loop writes.
template <class Gen>   // gen(first, count, blob, offs, values or nullptr) -> items kept; *bytes
static uint64_t stitched(uint64_t first, uint64_t n, char* blob, uint32_t* offs, uint32_t* values, uint64_t* bytes_out, Gen gen) {
    const unsigned T = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(std::min(64u, std::max(1u, std::thread::hardware_concurrency())), n / 100000)));
    if (T <= 1) { uint64_t b = 0; const uint64_t k = gen(first, n, blob, offs, values, &b); if (bytes_out) *bytes_out = b; return k; }
    struct Part { std::vector<char> blob; std::vector<uint32_t> offs, values; uint64_t kept = 0, bytes = 0, b = 0, e = 0; };
    std::vector<Part> parts(T);
    const uint64_t chunk = (n + T - 1) / T;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) {
        Part& pt = parts[t];
        pt.b = std::min<uint64_t>(n, t * chunk); pt.e = std::min<uint64_t>(n, pt.b + chunk);
        th.emplace_back([&pt, &gen, first, values] {
            const uint64_t cnt = pt.e - pt.b;
            pt.blob.resize(cnt * kMaxLen + 16); pt.offs.resize(cnt + 1);
            if (values) pt.values.resize(cnt);
            pt.kept = gen(first + pt.b, cnt, pt.blob.data(), pt.offs.data(), values ? pt.values.data() : nullptr, &pt.bytes);
        });
    }
    for (auto& x : th) x.join();
    th.clear();
    std::vector<uint64_t> kbase(T + 1, 0), bbase(T + 1, 0);
    for (unsigned t = 0; t < T; ++t) { kbase[t + 1] = kbase[t] + parts[t].kept; bbase[t + 1] = bbase[t] + parts[t].bytes; }
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            const Part& pt = parts[t];
            std::memcpy(blob + bbase[t], pt.blob.data(), pt.bytes);
            for (uint64_t k = 0; k < pt.kept; ++k) offs[kbase[t] + k] = pt.offs[k] + static_cast<uint32_t>(bbase[t]);
            if (values) std::memcpy(values + kbase[t], pt.values.data(), pt.kept * sizeof(uint32_t));
        });
    for (auto& x : th) x.join();
    offs[kbase[T]] = static_cast<uint32_t>(bbase[T]);
    if (bytes_out) *bytes_out = bbase[T];
    return kbase[T];
}

extern "C" {

struct wl_params {
    uint32_t R, S, D, K, M, F;
    double p_plus, p_hash, p_root_plus;
    uint64_t seed;
};

static Params cvt(const wl_params* w) { return Params{w->R, w->S, w->D, w->K, w->M, w->F, w->p_plus, w->p_hash, w->p_root_plus, w->seed}; }

uint64_t wl_space(const wl_params* w) { return space(cvt(w)); }
uint32_t wl_max_len() { return kMaxLen; }

// Subscriptions [first, first+n).  blob must hold n*wl_max_len() bytes, offs n+1 entries.  Returns bytes written.
// Mix: u<p_plus: replace 1 (80%) or 2 (20%) distinct levels of {1..5} by '+' (and with
// p_root_plus also level 0); p_plus<=u<p_plus+p_hash: truncate to depth d (2:5% 3:45% 4:30% 5:20%) + '#'; else exact.
static uint64_t gen_subs_serial(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs) {
    Params p = cvt(w);
    char* o = blob;
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        offs[k] = static_cast<uint32_t>(o - blob);
        Rng g(p.seed, i, 1);
        Tp t = decode(p, g.below(space(p)));
        double u = g.uni();
        uint32_t mask = 0; int depth = 6;
        if (u < p.p_plus) {
            uint32_t a = 1 + g.below(5);
            mask |= 1u << a;
            if (g.uni() < 0.20) { uint32_t b = 1 + g.below(4); if (b >= a) ++b; mask |= 1u << b; }
            if (g.uni() < p.p_root_plus) mask |= 1u;
        } else if (u < p.p_plus + p.p_hash) {
            double v = g.uni();
            depth = v < 0.05 ? 2 : v < 0.50 ? 3 : v < 0.80 ? 4 : 5;
        }
        o = put_filter(o, t, mask, depth);
    }
    offs[n] = static_cast<uint32_t>(o - blob);
    return static_cast<uint64_t>(o - blob);
}

// Root-hash shard of the subscription set: keeps subscription i iff its level 0 is '+' (replicated on every
// shard) or its region is flagged in region_keep[R].  values[k] = original subscription index.
// Returns the number kept; *bytes = blob bytes written.
static uint64_t gen_subs_sharded_serial(const wl_params* w, uint64_t first, uint64_t n, const uint8_t* region_keep,
                             char* blob, uint32_t* offs, uint32_t* values, uint64_t* bytes) {
    Params p = cvt(w);
    char* o = blob;
    uint64_t kept = 0;
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        Rng g(p.seed, i, 1);
        Tp t = decode(p, g.below(space(p)));
        double u = g.uni();
        uint32_t mask = 0; int depth = 6;
        if (u < p.p_plus) {
            uint32_t a = 1 + g.below(5);
            mask |= 1u << a;
            if (g.uni() < 0.20) { uint32_t b = 1 + g.below(4); if (b >= a) ++b; mask |= 1u << b; }
            if (g.uni() < p.p_root_plus) mask |= 1u;
        } else if (u < p.p_plus + p.p_hash) {
            double v = g.uni();
            depth = v < 0.05 ? 2 : v < 0.50 ? 3 : v < 0.80 ? 4 : 5;
        }
        if (!(mask & 1u) && !region_keep[t.r]) continue;
        offs[kept] = static_cast<uint32_t>(o - blob);
        values[kept] = static_cast<uint32_t>(i);
        o = put_filter(o, t, mask, depth);
        ++kept;
    }
    offs[kept] = static_cast<uint32_t>(o - blob);
    if (bytes) *bytes = static_cast<uint64_t>(o - blob);
    return kept;
}

// Publish topics [first, first+n).  frac_from_subs of them re-use the concrete topic of a uniformly drawn
// subscription in [0, n_subs) (C1: 0.5); the rest are uniform over the topic space, optionally restricted to
// the given regions (multi-GPU partitioning by root).  `stream` separates independent batches.
static uint64_t gen_topics_serial(const wl_params* w, uint64_t first, uint64_t n, double frac_from_subs, uint64_t n_subs,
                       const uint32_t* regions, uint32_t nreg, uint64_t stream, char* blob, uint32_t* offs) {
    Params p = cvt(w);
    char* o = blob;
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        offs[k] = static_cast<uint32_t>(o - blob);
        Rng g(p.seed, i, 2 + stream);
        uint64_t x;
        if (n_subs && g.uni() < frac_from_subs) x = sub_topic_index(p, g.below(n_subs));
        else if (nreg) x = pick_region_restricted(p, g, regions, nreg);
        else x = g.below(space(p));
        o = put_filter(o, decode(p, x), 0, 6);
    }
    offs[n] = static_cast<uint32_t>(o - blob);
    return static_cast<uint64_t>(o - blob);
}

// Publish topics with a Zipf(s = 1) popularity over DEVICES (secondary workload of SURVEY §8d): device rank k is
// drawn with P(rank <= k) ~ ln(k+1)/ln(N+1) (continuous approximation of the harmonic CDF), ranks are mapped to
// device ids by a fixed permutation; sensor / metric / channel stay uniform.
uint64_t wl_gen_topics_zipf(const wl_params* w, uint64_t first, uint64_t n, uint64_t stream, char* blob, uint32_t* offs) {
    Params p = cvt(w);
    char* o = blob;
    const uint64_t ndev = 1ull * p.R * p.S * p.D, per_dev = 1ull * p.K * p.M * p.F;
    const double lnN = std::log(static_cast<double>(ndev) + 1.0);
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        offs[k] = static_cast<uint32_t>(o - blob);
        Rng g(p.seed, i, 40 + stream);
        uint64_t rank = static_cast<uint64_t>(std::exp(g.uni() * lnN)) - 1;
        if (rank >= ndev) rank = ndev - 1;
        const uint64_t dev = permute(rank, ndev, p.seed ^ 0x5A17);
        o = put_filter(o, decode(p, dev * per_dev + g.below(per_dev)), 0, 6);
    }
    offs[n] = static_cast<uint32_t>(o - blob);
    return static_cast<uint64_t>(o - blob);
}

// Retained topics: n DISTINCT topics sampled without replacement (item i -> permute(i)).
static uint64_t gen_retained_serial(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs) {
    Params p = cvt(w);
    char* o = blob;
    uint64_t N = space(p);
    for (uint64_t k = 0; k < n; ++k) {
        offs[k] = static_cast<uint32_t>(o - blob);
        o = put_filter(o, decode(p, permute(first + k, N, p.seed)), 0, 6);
    }
    offs[n] = static_cast<uint32_t>(o - blob);
    return static_cast<uint64_t>(o - blob);
}

// SUBSCRIBE filters for the retained lookup (C4): forced wildcard, 85% '+' (1 or 2 levels as above), 15% '#' with d>=3.
uint64_t wl_gen_retain_filters(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs) {
    Params p = cvt(w);
    char* o = blob;
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        offs[k] = static_cast<uint32_t>(o - blob);
        Rng g(p.seed, i, 7);
        Tp t = decode(p, g.below(space(p)));
        uint32_t mask = 0; int depth = 6;
        if (g.uni() < 0.85) {
            uint32_t a = 1 + g.below(5);
            mask |= 1u << a;
            if (g.uni() < 0.20) { uint32_t b = 1 + g.below(4); if (b >= a) ++b; mask |= 1u << b; }
            if (g.uni() < p.p_root_plus) mask |= 1u;
        } else {
            double v = g.uni();
            depth = v < 0.45 ? 3 : v < 0.78 ? 4 : 5;
        }
        o = put_filter(o, t, mask, depth);
    }
    offs[n] = static_cast<uint32_t>(o - blob);
    return static_cast<uint64_t>(o - blob);
}

// ---- the big generators on all host threads: see `stitched` above ----
uint64_t wl_gen_subs(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs) {
    uint64_t bytes = 0;
    stitched(first, n, blob, offs, nullptr, &bytes, [w](uint64_t f, uint64_t c, char* b, uint32_t* o, uint32_t*, uint64_t* by) { *by = gen_subs_serial(w, f, c, b, o); return c; });
    return bytes;
}
uint64_t wl_gen_subs_sharded(const wl_params* w, uint64_t first, uint64_t n, const uint8_t* region_keep, char* blob, uint32_t* offs, uint32_t* values, uint64_t* bytes) {
    return stitched(first, n, blob, offs, values, bytes, [w, region_keep](uint64_t f, uint64_t c, char* b, uint32_t* o, uint32_t* v, uint64_t* by) { return gen_subs_sharded_serial(w, f, c, region_keep, b, o, v, by); });
}
uint64_t wl_gen_topics(const wl_params* w, uint64_t first, uint64_t n, double frac_from_subs, uint64_t n_subs, const uint32_t* regions, uint32_t nreg, uint64_t stream, char* blob, uint32_t* offs) {
    uint64_t bytes = 0;
    stitched(first, n, blob, offs, nullptr, &bytes, [=](uint64_t f, uint64_t c, char* b, uint32_t* o, uint32_t*, uint64_t* by) { *by = gen_topics_serial(w, f, c, frac_from_subs, n_subs, regions, nreg, stream, b, o); return c; });
    return bytes;
}
uint64_t wl_gen_retained(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs) {
    uint64_t bytes = 0;
    stitched(first, n, blob, offs, nullptr, &bytes, [w](uint64_t f, uint64_t c, char* b, uint32_t* o, uint32_t*, uint64_t* by) { *by = gen_retained_serial(w, f, c, b, o); return c; });
    return bytes;
}

// level-0 string of region r ("reg-%02d"); returns length
uint32_t wl_region_name(uint32_t r, char* out) { Tp t{}; t.r = r; return static_cast<uint32_t>(put_level(out, t, 0) - out); }

}  // extern "C"
