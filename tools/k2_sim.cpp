// Host-side replay of k_match_fast's per-lane control flow over the REAL device tables (the host mirror is in device
// layout), workload C3: how many loop iterations each lane of a 32-topic tile needs, how uneven that is, and what the
// publish phase has to expand.  No GPU involved — this sizes the instruction-side levers named in DESIGN.md §4.8
// (lane imbalance, publish rows) before spending GPU time on them.
//   build: g++ -O3 -std=c++17 -o k2_sim tools/k2_sim.cpp rmqtt_b200/csrc/host_trie.cpp rmqtt_b200/csrc/workload.cpp
//   run:   ./k2_sim [n_subs=10000000] [n_topics=1000000]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../rmqtt_b200/csrc/host_trie.h"

extern "C" {
struct wl_params { uint32_t R, S, D, K, M, F; double p_plus, p_hash, p_root_plus; uint64_t seed; };
uint32_t wl_max_len();
uint64_t wl_gen_subs(const wl_params* w, uint64_t first, uint64_t n, char* blob, uint32_t* offs);
uint64_t wl_gen_topics(const wl_params* w, uint64_t first, uint64_t n, double frac_from_subs, uint64_t n_subs, const uint32_t* regions, uint32_t nreg,
                       uint64_t stream, char* blob, uint32_t* offs);
}
using namespace gm;

struct Rec { u32 node, plus, hash_ref, own_ref, mask, cnts; };

int main(int argc, char** argv) {
    const uint64_t n_subs = argc > 1 ? strtoull(argv[1], nullptr, 10) : 10000000ull, n_top = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1000000ull;
    wl_params P{64, 64, 256, 8, 4, 2, 0.30, 0.05, 0.02, 0xC3};
    std::vector<char> blob(n_subs * wl_max_len()); std::vector<u32> offs(n_subs + 1), vals(n_subs);
    wl_gen_subs(&P, 0, n_subs, blob.data(), offs.data());
    std::iota(vals.begin(), vals.end(), 0u);
    HostTrie t(128);
    t.reserve(n_subs);
    t.insert_batch(blob.data(), offs.data(), vals.data(), n_subs);
    t.sync();
    std::vector<char> tb(n_top * wl_max_len()); std::vector<u32> to(n_top + 1);
    wl_gen_topics(&P, 0, n_top, 0.0, 0, nullptr, 0, 0, tb.data(), to.data());
    // tokenise + locality order (same grouping as k_bucket_*: by the first two level tokens)
    struct Top { u32 tok[8]; u32 L; u64 key; };
    std::vector<Top> tops(n_top);
    for (u64 i = 0; i < n_top; ++i) {
        const char* s = tb.data() + to[i]; const u32 len = to[i + 1] - to[i];
        Top& q = tops[i]; q.L = 0;
        for (u32 a = 0; a <= len;) { u32 b = a; while (b < len && s[b] != '/') ++b; if (q.L < 8) q.tok[q.L] = t.lookup_token(s + a, b - a); q.L++; a = b + 1; }
        q.key = (static_cast<u64>(fmix32(q.tok[0] * 0x9E3779B1u + q.tok[1]) & 0x3FFFu) << 32) | i;
    }
    std::sort(tops.begin(), tops.end(), [](const Top& a, const Top& b) { return a.key < b.key; });
    const u32 wm = t.win_mask(), ws = t.win_shift(), nm = t.nwin_mask();
    auto rec_of = [&](const EdgeSlot& e) { return Rec{e.child, e.plus, e.hash_ref, e.own_ref, e.mask, e.cnts}; };
    const u32 cf_mask = static_cast<u32>(t.cfilter.size() - 1);
    u64 cf_loads = 0;
    auto cf_maybe = [&](u32 parent, u32 token) { u32 w, bits; cfilter_pos(parent, token, cf_mask, w, bits); ++cf_loads; return (t.cfilter[w] & bits) == bits; };   // the child filter of wide nodes (kernels.cuh cfilter_maybe)
    // per-topic replay
    std::vector<u32> iters(n_top);
    std::vector<std::vector<u32>> descs(n_top);
    u64 loads = 0, probe_steps = 0;
    u64 fail_loads[9] = {0}, fail_probes[9] = {0}, hit_extra[9] = {0}, hit_probes[9] = {0};   // by depth of the probing node
    for (u64 i = 0; i < n_top; ++i) {
        const Top& q = tops[i];
        Rec r{0, t.root_plus, t.root_hash_ref, 0, t.root_mask, t.root_hash_cnt};
        u32 d = 0, pmask = 0, pend[8] = {0}, it = 0;
        auto& ds = descs[i];
        for (;;) {
            ++it;
            const u32 c1 = r.cnts & 0xFFFFu, c2 = d == q.L ? r.cnts >> 16 : 0u;
            if (c1) ds.push_back(c1);
            if (c2) ds.push_back(c2);
            bool probe = false; u32 idx = 0, nd = 0, kt = 0;
            if (d < q.L) {
                if (r.plus) { pend[d] = r.plus; pmask |= 1u << d; }
                const u32 tk = q.tok[d];
                if ((r.mask & MASK_BLOOM) && tk != TOK_UNKNOWN && (r.mask & mask_bit(tk)) && (!(r.mask & MASK_WIDE_FLAG) || cf_maybe(r.node, tk))) { probe = true; kt = tk; nd = d + 1; idx = edge_slot0(r.node, tk, r.mask >> WTAG_SHIFT, wm, ws, nm); }
            }
            bool done = false; const EdgeSlot* e = nullptr;
            u32 steps = 0;
            for (;;) {
                if (!probe) { if (!pmask) { done = true; break; } u32 pd = 31 - __builtin_clz(pmask); pmask &= ~(1u << pd); idx = pend[pd] - 1; nd = pd + 1; }
                e = &t.edges[idx]; ++loads;
                if (!probe) break;
                ++probe_steps; ++steps;
                if (e->child == 0) { probe = false; fail_loads[std::min(d, 8u)] += steps; fail_probes[std::min(d, 8u)]++; continue; }
                if (e->parent == r.node && e->token == kt) { hit_extra[std::min(d, 8u)] += steps - 1; hit_probes[std::min(d, 8u)]++; break; }
                idx = edge_next(idx, wm);
            }
            if (done) break;
            r = rec_of(*e); d = nd;
        }
        iters[i] = it;
    }
    // per-tile statistics
    const u64 ntiles = n_top / 32;
    double sum_it = 0, sum_max = 0, pub_rows = 0, pub_e0 = 0, ids = 0, single_desc = 0, ndesc = 0, single_ids = 0, flat_e0 = 0, lane_ids_max = 0;
    u64 hist[9] = {0};
    u64 sz_sets[8] = {0}, sz_ids[8] = {0};                 // value sets / ids by set size: 1, 2-3, 4-7, 8-15, 16-31, 32-63, 64-127, 128+
    for (u64 tl = 0; tl < ntiles; ++tl) {
        u32 mx = 0, rows = 0; u64 tot_ids = 0, lane_max = 0;
        for (u32 l = 0; l < 32; ++l) {
            const u64 i = tl * 32 + l;
            sum_it += iters[i]; mx = std::max(mx, iters[i]); rows = std::max<u32>(rows, descs[i].size());
            u64 li = 0;
            for (u32 c : descs[i]) {
                li += c; ++ndesc;
                if (c == 1) { ++single_desc; ++single_ids; }
                const u32 b = std::min<u32>(7, 31 - __builtin_clz(c));
                sz_sets[b]++; sz_ids[b] += c;
            }
            tot_ids += li; lane_max = std::max(lane_max, li);
        }
        sum_max += mx; pub_rows += rows; ids += tot_ids; lane_ids_max += lane_max;
        for (u32 k = 0; k < rows; ++k) { u64 tk = 0; for (u32 l = 0; l < 32; ++l) { const auto& ds = descs[tl * 32 + l]; if (k < ds.size()) tk += ds[k]; } pub_e0 += (tk + 31) / 32; }
        flat_e0 += (tot_ids + 31) / 32;
        const double util = sum_it ? 0 : 0; (void)util;
        u32 b = std::min<u32>(8, mx / 8); hist[b]++;
    }
    const double T = static_cast<double>(ntiles);
    printf("C3 replay: %llu subscriptions, %llu topics, %llu tiles; table %zu slots in %u windows\n", (unsigned long long)n_subs, (unsigned long long)n_top, (unsigned long long)ntiles, t.edges.size(), nm + 1);
    printf("walk   : %.2f loop iterations per topic (= node visits), %.2f slot loads per topic (%.3f loads per literal probe)\n", sum_it / (T * 32), double(loads) / n_top, double(loads) / std::max<u64>(1, probe_steps));
    printf("         child-filter words read: %.2f per topic (a few MB, L2-resident)\n", double(cf_loads) / n_top);
    printf("         literal probes by depth of the probing node [hits: probes/topic, extra loads/probe | misses (Bloom false positives): probes/topic, loads/probe]:\n");
    for (int dd = 0; dd < 8; ++dd) if (hit_probes[dd] + fail_probes[dd]) printf("           depth %d: hits %.3f x %.3f extra | misses %.3f x %.3f\n", dd, double(hit_probes[dd]) / n_top, double(hit_extra[dd]) / std::max<u64>(1, hit_probes[dd]), double(fail_probes[dd]) / n_top, double(fail_loads[dd]) / std::max<u64>(1, fail_probes[dd]));
    printf("         per tile: mean of lanes %.1f, max lane %.1f -> lane utilisation of the walk loop %.1f %%\n", sum_it / (T * 32), sum_max / T, 100.0 * sum_it / (32.0 * sum_max));
    printf("         tiles by max-lane iterations [0-7,8-15,...,64+]:");
    for (int b = 0; b < 9; ++b) printf(" %.1f%%", 100.0 * hist[b] / T);
    printf("\npublish: %.1f ids per topic, %.2f value sets per topic (%.1f %% single-value sets holding %.1f %% of the ids)\n", ids / (T * 32), ndesc / (T * 32), 100.0 * single_desc / ndesc, 100.0 * single_ids / ids);
    printf("         per tile: %.1f rows (max sets of a lane), %.1f expansion iterations of 32 ids (row by row) vs %.1f if all %.0f ids were one flat list\n", pub_rows / T, pub_e0 / T, flat_e0 / T, ids / T);
    printf("         value sets by size [1, 2-3, 4-7, 8-15, 16-31, 32-63, 64-127, 128+]: sets");
    for (int b = 0; b < 8; ++b) printf(" %.1f%%", 100.0 * sz_sets[b] / ndesc);
    printf("  ids");
    for (int b = 0; b < 8; ++b) printf(" %.1f%%", 100.0 * sz_ids[b] / ids);
    printf("\n");
    printf("         lane-private copy would loop max-lane ids = %.1f times per tile (mean lane %.1f)\n", lane_ids_max / T, ids / (T * 32));
    return 0;
}
