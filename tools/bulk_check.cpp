// Developer check of the parallel bulk insert (host_trie.cpp insert_batch_parallel): builds the same filter set once
// one-by-one-equivalent (GM_BULK_SERIAL=1) and once with all host threads, and compares an order-independent fingerprint
// of the two tries (every node: path, value set, liveness, fan-out summary, window tag of depth <= 2) plus the invariants
// of the edge table (every node is found from its parent; window counts add up).
//   g++ -O2 -std=c++17 -pthread tools/bulk_check.cpp rmqtt_b200/csrc/host_trie.cpp rmqtt_b200/csrc/workload.cpp -o /tmp/bulk_check && /tmp/bulk_check 10000000
#define private public
#include "../rmqtt_b200/csrc/host_trie.h"
#undef private
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

struct wl_params { uint32_t R, S, D, K, M, F; double p_plus, p_hash, p_root_plus; uint64_t seed; };
extern "C" uint64_t wl_gen_subs(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs);
extern "C" uint32_t wl_max_len();

using namespace gm;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

struct Print { u64 sum = 0, xr = 0, nodes = 0; bool ok = true; };

static Print fingerprint(HostTrie& t) {
    Print p;
    std::vector<u64> path(t.nodes_.size(), 0);
    // parents precede children in both numberings?  not in general: resolve by depth order
    std::vector<u32> order(t.nodes_.size());
    for (u32 i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return t.nodes_[a].depth < t.nodes_[b].depth; });
    u64 win_sum = 0;
    for (u64 c : t.win_count_) win_sum += c;
    if (win_sum != t.edge_count_ || t.edge_count_ + 1 != t.nodes_.size()) { printf("count mismatch: windows %llu edges %llu nodes %zu\n", (unsigned long long)win_sum, (unsigned long long)t.edge_count_, t.nodes_.size()); p.ok = false; }
    for (u32 id : order) {
        const HNode& n = t.nodes_[id];
        if (id) {
            path[id] = mix(path[n.parent] * 0x9E3779B97F4A7C15ull + n.token + 1);
            const u32 slot = t.find_edge(n.parent, n.token, t.nodes_[n.parent].wtag);
            if (slot == 0xFFFFFFFFu || t.edges[slot].child != id || slot != n.edge_slot) { if (p.ok) printf("node %u not found from its parent\n", id); p.ok = false; }
            if (n.depth != t.nodes_[n.parent].depth + 1) p.ok = false;
            if (n.depth > 2 && n.wtag != t.nodes_[n.parent].wtag) { if (p.ok) printf("node %u: window tag not inherited\n", id); p.ok = false; }
            if (n.depth <= 1 && n.wtag != 0) p.ok = false;
        }
        u64 h = path[id];
        h = mix(h + n.nvals * 31 + n.alive * 7 + n.live_children * 131 + n.lit_children * 1031 + (n.plus_child ? 3 : 0) + (n.hash_child ? 5 : 0) + n.wide * 11 + (u64(n.mask) << 20));
        if (n.nvals == 1) h = mix(h + n.v0);
        else if (n.nvals > 1) for (u32 v : t.multi_[id]) h = mix(h + v);
        if (n.plus_child && (t.nodes_[n.plus_child].parent != id || t.nodes_[n.plus_child].token != TOK_PLUS)) p.ok = false;
        if (n.hash_child && (t.nodes_[n.hash_child].parent != id || t.nodes_[n.hash_child].token != TOK_HASH)) p.ok = false;
        p.sum += h; p.xr ^= h; p.nodes++;
    }
    return p;
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
    wl_params w{64, 64, 256, 8, 4, 2, 0.30, 0.05, 0.02, 0xC3};
    std::vector<char> blob(n * wl_max_len());
    std::vector<u32> offs(n + 1), vals(n);
    wl_gen_subs(&w, 0, n, blob.data(), offs.data());
    for (u64 i = 0; i < n; ++i) vals[i] = static_cast<u32>(i);
    // a few invalid / duplicate / odd filters in the middle
    setenv("GM_BULK_PROFILE", "1", 1);
    Print fp[2]; u64 changed[2]; u64 stats[2][6];
    std::vector<std::array<u32, 8>> dict_rows[2]; std::vector<u8> pools[2]; std::vector<u8> dollars[2];
    for (int par = 0; par < 2; ++par) {
        if (par) unsetenv("GM_BULK_SERIAL"); else setenv("GM_BULK_SERIAL", "1", 1);
        HostTrie t(128);
        double a = now();
        t.reserve(n);
        double b = now();
        changed[par] = t.insert_batch(blob.data(), offs.data(), vals.data(), n);
        double c = now();
        const bool ok = t.sync();
        double d = now();
        printf("%s: reserve %.2f s, insert_batch %.2f s, sync %.2f s (ok %d), changed %llu, nodes %zu, dict %llu, plus %llu, values %llu, live %llu\n", par ? "parallel" : "serial", b - a, c - b, d - c, ok,
               (unsigned long long)changed[par], t.nodes_.size(), (unsigned long long)t.dict_count_, (unsigned long long)t.plus_count_, (unsigned long long)t.values_size_, (unsigned long long)t.live_nodes_);
        u64 s[6] = {t.nodes_.size(), t.dict_count_, t.plus_count_, t.values_size_, t.live_nodes_, t.values.size()};
        std::memcpy(stats[par], s, sizeof s);
        {   // balance of the windows / tags (what the match kernel's probe length depends on)
            u64 wmax = 0, wsum = 0, tmax = 0, tsum = 0; size_t nz = 0;
            for (u64 c : t.win_count_) { wmax = std::max(wmax, c); wsum += c; }
            for (size_t k = 1; k < t.tag_count_.size(); ++k) { tmax = std::max(tmax, t.tag_count_[k]); tsum += t.tag_count_[k]; nz += t.tag_count_[k] != 0; }
            printf("  windows %zu: max / mean edges %.3f; tags (without the hot tag 0): max / mean %.3f over %zu used tags; table load %.3f\n", t.win_count_.size(), wmax / (double(wsum) / t.win_count_.size()),
                   tmax / (double(tsum) / std::max<size_t>(1, nz)), nz, double(t.edge_count_) / t.edges.size());
        }
        for (const DictSlot& d : t.dict) if (d.w[0]) { std::array<u32, 8> r; std::memcpy(r.data(), d.w, 32); dict_rows[par].push_back(r); }
        std::sort(dict_rows[par].begin(), dict_rows[par].end());
        pools[par] = t.pool; dollars[par] = t.tok_dollar_;
        for (const auto& r : dict_rows[par]) if (t.lookup_token((r[7] >> 24) == 0xFF ? reinterpret_cast<const char*>(t.pool.data() + r[2]) : reinterpret_cast<const char*>(&r[1]), (r[7] >> 24) == 0xFF ? r[1] : (r[7] >> 24)) != r[0]) { printf("  dictionary: token %u is not found by its string\n", r[0]); fp[par].ok = false; break; }
        { const bool keep_ok = fp[par].ok; fp[par] = fingerprint(t); fp[par].ok &= keep_ok; }
        printf("  fingerprint sum %016llx xor %016llx over %llu nodes, invariants %s\n", (unsigned long long)fp[par].sum, (unsigned long long)fp[par].xr, (unsigned long long)fp[par].nodes, fp[par].ok ? "ok" : "BROKEN");
    }
    const bool dict_same = dict_rows[0] == dict_rows[1] && pools[0] == pools[1] && dollars[0] == dollars[1];
    if (!dict_same) printf("dictionary / pool / '$' flags differ (%zu vs %zu entries)\n", dict_rows[0].size(), dict_rows[1].size());
    const bool same = dict_same && fp[0].sum == fp[1].sum && fp[0].xr == fp[1].xr && fp[0].nodes == fp[1].nodes && changed[0] == changed[1] && std::memcmp(stats[0], stats[1], sizeof stats[0]) == 0 && fp[0].ok && fp[1].ok;
    printf("%s\n", same ? "IDENTICAL CONTENT" : "MISMATCH");
    return same ? 0 : 1;
}
