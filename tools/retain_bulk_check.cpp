// Developer check of the retained tree's bulk build (retain_tree.cpp set_batch_build): the same topics once through
// set() x n + flatten() and once through the level-wise parallel build; every shipped array must be identical entry for entry
// (the hash table: as a set of entries), and the host trees must be the same tree.
//   g++ -O2 -std=c++17 -pthread tools/retain_bulk_check.cpp rmqtt_b200/csrc/retain_tree.cpp rmqtt_b200/csrc/host_trie.cpp rmqtt_b200/csrc/workload.cpp -o /tmp/retain_bulk_check && /tmp/retain_bulk_check 5000000
#define private public
#include "../rmqtt_b200/csrc/retain_tree.h"
#undef private
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

struct wl_params { uint32_t R, S, D, K, M, F; double p_plus, p_hash, p_root_plus; uint64_t seed; };
extern "C" uint64_t wl_gen_retained(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs);
extern "C" uint32_t wl_max_len();
using namespace gm;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class V> static bool same_bytes(const char* what, const V& a, const V& b) {
    if (a.size() != b.size()) { printf("  %s: sizes differ %zu vs %zu\n", what, a.size(), b.size()); return false; }
    if (a.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(a[0])) != 0) {
        size_t k = 0; while (std::memcmp(&a[k], &b[k], sizeof(a[0])) == 0) ++k;
        printf("  %s: first difference at entry %zu of %zu\n", what, k, a.size()); return false;
    }
    return true;
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
    wl_params w{64, 64, 256, 8, 4, 2, 0.30, 0.05, 0.02, 0xC4};
    std::string blob; std::vector<u32> offs, vals;
    {
        std::vector<char> b(n * wl_max_len()); std::vector<u32> o(n + 1);
        const u64 bytes = wl_gen_retained(&w, 0, n, b.data(), o.data());
        blob.assign(b.data(), bytes); offs.assign(o.begin(), o.end());
    }
    // odd topics: `$`-prefixed roots, literal '+' / '#' levels, blanks, duplicates (the last value stays), invalid ones
    const char* extra[] = {"$SYS/broker/uptime", "$SYS/broker/load", "$share/x", "a/+/b", "a/#", "lit/#", "lit/+/x/#", "", "/", "//", "a//b", "dup/t", "dup/t", "dup/t", "a/$b", "bad/#/x",
                           "reg-00/site-0000", "reg-00", "x/y/z/w/v/u/t/s"};
    for (const char* e : extra) { blob += e; offs.push_back(static_cast<u32>(blob.size())); }
    const u64 N = offs.size() - 1;
    vals.resize(N);
    for (u64 i = 0; i < N; ++i) vals[i] = static_cast<u32>(i * 7 + 1);
    setenv("GM_BULK_PROFILE", "1", 1);
    HostTrie d0(128), d1(128);
    RetainTreeHost a(&d0), b(&d1);
    setenv("GM_BULK_SERIAL", "1", 1);
    double t = now();
    const u64 oka = a.set_batch(blob.data(), offs.data(), vals.data(), N);
    double t1 = now();
    a.prepare_flush();
    printf("one by one: set x n %.2f s, flatten %.2f s\n", t1 - t, now() - t1);
    unsetenv("GM_BULK_SERIAL");
    t = now();
    const u64 okb = b.set_batch(blob.data(), offs.data(), vals.data(), N);
    t1 = now();
    b.prepare_flush();
    printf("bulk build: %.2f s (+ prepare_flush %.2f s)\n", t1 - t, now() - t1);
    bool ok = oka == okb;
    if (!ok) printf("  valid topics differ %llu vs %llu\n", (unsigned long long)oka, (unsigned long long)okb);
    ok &= same_bytes("rnodes", a.rnodes, b.rnodes);
    ok &= same_bytes("rkids", a.rkids, b.rkids);
    ok &= same_bytes("rvals", a.rvals, b.rvals);
    ok &= same_bytes("rparent", a.rparent_, b.rparent_);
    ok &= same_bytes("rtoken", a.rtoken_, b.rtoken_);
    ok &= same_bytes("rcap", a.rcap_, b.rcap_);
    ok &= same_bytes("in_rvals", a.in_rvals_, b.in_rvals_);
    if (a.root_plain_kids != b.root_plain_kids || a.root_plain_val_hi != b.root_plain_val_hi || a.max_depth != b.max_depth || a.n_nodes_ != b.n_nodes_ || a.n_values_ != b.n_values_ ||
        a.live_edges_ != b.live_edges_ || a.flat_valid_ != b.flat_valid_ || a.full != b.full) {
        printf("  scalars differ: plain kids %u/%u, plain val hi %u/%u, depth %u/%u, nodes %llu/%llu, values %llu/%llu\n", a.root_plain_kids, b.root_plain_kids, a.root_plain_val_hi, b.root_plain_val_hi,
               a.max_depth, b.max_depth, (unsigned long long)a.n_nodes_, (unsigned long long)b.n_nodes_, (unsigned long long)a.n_values_, (unsigned long long)b.n_values_);
        ok = false;
    }
    // hash table: same size, same entries
    if (a.redges.size() != b.redges.size()) { printf("  redges sizes differ\n"); ok = false; }
    else {
        u64 cnt = 0;
        for (size_t s = 0; s < a.redges.size(); ++s) {
            const REdge& e = a.redges[s];
            if (!e.child) continue;
            ++cnt;
            const u32 bs = b.edge_slot_of(e.parent, e.token);
            if (bs == 0xFFFFFFFFu || std::memcmp(&b.redges[bs], &e, sizeof e) != 0) { if (ok) printf("  redges: entry (%u, %u) differs\n", e.parent, e.token); ok = false; }
        }
        u64 cntb = 0;
        for (size_t s = 0; s < b.redges.size(); ++s) cntb += b.redges[s].child != 0;
        if (cnt != cntb) { printf("  redges: %llu vs %llu entries\n", (unsigned long long)cnt, (unsigned long long)cntb); ok = false; }
    }
    // host trees, compared through the device numbering
    if (a.nodes_.size() != b.nodes_.size()) { printf("  host nodes %zu vs %zu\n", a.nodes_.size(), b.nodes_.size()); ok = false; }
    else {
        std::vector<u32> a_of_dev(a.nodes_.size());
        for (u32 h = 0; h < a.nodes_.size(); ++h) a_of_dev[a.nodes_[h].dev] = h;
        for (u32 hb = 0; hb < b.nodes_.size() && ok; ++hb) {
            const auto& nb = b.nodes_[hb];
            const auto& na = a.nodes_[a_of_dev[nb.dev]];
            bool same = na.has_val == nb.has_val && na.val == nb.val && na.token == nb.token && a.nodes_[na.parent].dev == b.nodes_[nb.parent].dev && na.kids.size() == nb.kids.size();
            for (size_t k = 0; same && k < na.kids.size(); ++k) same = na.kids[k].first == nb.kids[k].first && a.nodes_[na.kids[k].second].dev == b.nodes_[nb.kids[k].second].dev;
            if (!same) { printf("  host node (device %u) differs\n", nb.dev); ok = false; }
        }
    }
    // and the bulk-built tree keeps working incrementally: the same edits on both, then a re-flatten of both must agree again
    const char* edits[] = {"reg-00/site-0000/dev-0000001/sen-0/met-0/ch-0", "new/branch/x", "dup/t", "$SYS/new"};
    for (const char* e : edits) { bool h; u32 o; a.set(e, static_cast<u32>(strlen(e)), 4242, &h, &o); b.set(e, static_cast<u32>(strlen(e)), 4242, &h, &o); }
    { bool h; u32 o; a.remove("dup/t", 5, &h, &o); b.remove("dup/t", 5, &h, &o); }
    a.prepare_flush(); b.prepare_flush();
    ok &= same_bytes("rnodes after edits", a.rnodes, b.rnodes);
    ok &= same_bytes("rkids after edits", a.rkids, b.rkids);
    ok &= same_bytes("rvals after edits", a.rvals, b.rvals);
    a.give_up(); b.give_up(); a.prepare_flush(); b.prepare_flush();
    ok &= same_bytes("rnodes after re-flatten", a.rnodes, b.rnodes);
    ok &= same_bytes("rkids after re-flatten", a.rkids, b.rkids);
    ok &= same_bytes("rvals after re-flatten", a.rvals, b.rvals);
    printf("%s (%llu topics, %zu nodes)\n", ok ? "IDENTICAL IMAGE" : "MISMATCH", (unsigned long long)N, b.nodes_.size());
    return ok ? 0 : 1;
}
