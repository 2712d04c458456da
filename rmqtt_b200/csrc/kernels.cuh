// sm_100a kernels of the PUBLISH -> matching-subscribers hot path.
//
//   k_tokenize     Topic::from_str (rmqtt/src/topic.rs:326-363) for a batch: split on '/', classify,
//                  validate, and intern every level through the device dictionary -> u32 tokens.
//   k_match_fast   TopicTree::matches (rmqtt/src/trie.rs:299-347, MatchedIter::prepare): one topic per
//                  thread walks the trie depth-first (one 256-bit load per visited node), then the warp
//                  publishes its 32 match lists with a scan + load-balanced (ballot/shuffle) expansion:
//                  single pass, per-topic contiguous output.
//   k_match_slow   the same walk for the topics the fast path defers (more levels than the fast
//                  path stages in shared memory, or more matches than its staging pool): one warp per
//                  topic, count pass + write pass, warp-cooperative value-range copies.
//   k_apply_patches  flush of 32-byte slot patches into the device tables (Router::add/remove).
//
// All arithmetic is u32 integer / pointer chasing: HBM- and L2-latency bound; no tensor cores.
#pragma once
#include <cuda_runtime.h>

#include "layout.h"

namespace gm {

struct MatchParams {
    TrieView tv;
    const u32* tok;      // [tok_levels][n]  (SoA: token of level l of topic t at tok[l*n+t])
    const u32* meta;     // [n]
    u32 n;
    u32 tok_levels;
    uint2* spans;        // [n] (offset, count) into out_ids
    u32* out_ids;
    unsigned long long cap_ids;
    unsigned long long* cursor;   // bump allocator over out_ids (final value = ids needed)
    u32* slow_list;      // [n]
    u32* slow_count;
    u32* tile_counter;
    unsigned long long* stats;    // [4] V,E,F,M + [4..] probe diagnostics (only written by STATS instantiations)
    u32 flags;                    // MP_* tuning switches
};
constexpr u32 MP_L2_HINTS = 1u;   // L2 eviction priorities: hot upper levels / '+' subtrees evict_last, cold deep chain evict_first

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld256(const void* p, u32 (&w)[8]) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}
// L2 eviction-priority policies (createpolicy) and a 256-bit load that carries one.
__device__ __forceinline__ u64 l2_policy_evict_last() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_evict_first() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void ld256_hint(const void* p, u32 (&w)[8], u64 pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p), "l"(pol));
}
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------------------
// K1: tokeniser.  One thread per topic.
constexpr int TOK_THREADS = 256;

__device__ __forceinline__ u32 dict_lookup(const TrieView& tv, u32 h, u32 len, const u32 (&w)[7],
                                           const u8* text) {
    u32 idx = dict_hash_finish(h, len) & tv.dict_mask;
    for (;;) {
        u32 s[8];
        ld256(tv.dict + idx, s);
        if (s[0] == 0) return TOK_UNKNOWN;
        if (len <= DICT_INLINE_MAX) {
            bool eq = true;
#pragma unroll
            for (int k = 0; k < 7; ++k) eq &= (s[k + 1] == w[k]);
            if (eq) return s[0];
        } else if ((s[1] & 0xFF) == 0xFF && s[2] == len && s[4] == h) {
            const u8* q = tv.pool + s[3];
            bool eq = true;
            for (u32 i = 0; i < len && eq; ++i) eq = (q[i] == text[i]);
            if (eq) return s[0];
        }
        idx = (idx + 1) & tv.dict_mask;
    }
}

__global__ void __launch_bounds__(TOK_THREADS)
k_tokenize(const u8* __restrict__ blob, const u32* __restrict__ offs, u32 n, TrieView tv, u32 tok_levels,
           u32* __restrict__ tok, u32* __restrict__ meta, int* __restrict__ status) {
    __shared__ u32 s_w[7][TOK_THREADS];   // packed level bytes of the current level, per thread (conflict-free)
    u32 t = blockIdx.x * TOK_THREADS + threadIdx.x;
    if (t >= n) return;
    u32 pos = offs[t];
    const u32 end = offs[t + 1];
    u32 lev = 0;
    bool invalid = false, dollar = false;
    for (;;) {
        u32 h = FNV_INIT, len = 0, c0 = 0, cur = 0;
        bool wild = false;
        const u32 start = pos;
#pragma unroll
        for (int k = 0; k < 7; ++k) s_w[k][threadIdx.x] = 0;
        // byte 0 of the 28-byte key area is the length; string byte i sits at position i+1
        while (pos < end) {
            u32 c = blob[pos];
            if (c == '/') break;
            h = fnv_step(h, c);
            wild |= (c == '+') | (c == '#');
            if (len == 0) c0 = c;
            if (len < DICT_INLINE_MAX) {
                u32 p = len + 1;
                cur |= c << (8 * (p & 3));
                if ((p & 3) == 3) { s_w[p >> 2][threadIdx.x] = cur; cur = 0; }
            }
            ++len; ++pos;
        }
        const bool last = pos >= end;
        u32 tk = TOK_UNKNOWN;
        if (len == 0) tk = TOK_BLANK;
        else if (len == 1 && c0 == '+') tk = TOK_PLUS;
        else if (len == 1 && c0 == '#') { tk = TOK_HASH; if (!last) invalid = true; }   // topic.rs:209
        else if (wild) invalid = true;                                                   // topic.rs:333-334
        else {
            if (c0 == '$') { if (lev > 0) invalid = true; else dollar = true; }           // topic.rs:210
            if (!invalid && lev < tok_levels) {
                u32 w[7];
                if (len <= DICT_INLINE_MAX) {
                    if ((len & 3) != 3) s_w[len >> 2][threadIdx.x] = cur;   // flush the partial word
#pragma unroll
                    for (int k = 0; k < 7; ++k) w[k] = s_w[k][threadIdx.x];
                    w[0] |= len;
                } else {
#pragma unroll
                    for (int k = 0; k < 7; ++k) w[k] = 0;
                }
                tk = dict_lookup(tv, h, len, w, blob + start);
            }
        }
        if (invalid) break;
        if (lev < tok_levels) tok[static_cast<size_t>(lev) * n + t] = tk;
        ++lev;
        if (last) break;
        ++pos;   // skip '/'
    }
    meta[t] = invalid ? META_INVALID : (lev | (dollar ? META_DOLLAR : 0u));
    status[t] = invalid ? -2 : 0;   // GM_ERR_INVALID_TOPIC: Topic::from_str would return Err
}

// ------------------------------------------------------------------------------------------------
// Work items.  lo = argument (parent node id for a literal probe, index into `plus` for a '+' hop),
// hi = topic slot (5 bits) | depth of the node being loaded << 5 | kind << 21.
constexpr u32 KIND_PROBE = 0, KIND_PLUS = 1;
__device__ __forceinline__ u64 make_item(u32 topic, u32 depth, u32 kind, u32 arg) {
    return (static_cast<u64>(topic | (depth << 5) | (kind << 21)) << 32) | arg;
}

struct NodeRec { u32 node, plus, hash_ref, own_ref, mask, cnts; };

// Loads the record an item points at.  Returns false when the literal child does not exist.
// pol != 0: L2 cache-policy operand for the slot loads.
__device__ __forceinline__ bool load_record(const TrieView& tv, u32 kind, u32 arg, u32 token, NodeRec& r, u64 pol = 0) {
    u32 s[8];
    if (kind == KIND_PLUS) {
        if (pol) ld256_hint(tv.plus + arg, s, pol); else ld256(tv.plus + arg, s);
        r.node = s[0]; r.plus = s[1]; r.hash_ref = s[2]; r.own_ref = s[3]; r.mask = s[4]; r.cnts = s[5];
        return true;
    }
    u32 idx = edge_hash(arg, token) & tv.edge_mask;
    for (;;) {
        if (pol) ld256_hint(tv.edges + idx, s, pol); else ld256(tv.edges + idx, s);
        if (s[2] == 0) return false;                       // empty slot: no such child
        if (s[0] == arg && s[1] == token) break;
        idx = (idx + 1) & tv.edge_mask;
    }
    r.node = s[2]; r.plus = s[3]; r.hash_ref = s[4]; r.own_ref = s[5]; r.mask = s[6]; r.cnts = s[7];
    return true;
}

// A matched value set waiting to be expanded into the output: values[ref .. ref+cnt) (or ref itself).
struct Desc { u32 ref, cnt; };

// K2: TopicTree::matches, one topic per thread (depth-first), one tile of 32 topics per warp.
//
//   walk     every thread follows its own topic down the trie: at a node it records the matched value
//            sets ('#' child always, own values on path exhaustion) as 8-byte descriptors, parks the
//            '+' child of this depth in its shared-memory column, and descends through the literal
//            child with ONE 256-bit load; when the literal path ends it resumes the deepest parked
//            '+' child.  The dependent chain of a thread is one load long per visited node; latency is
//            hidden by the other ~1.5 K resident threads of the SM.
//   publish  warp-cooperative: per-lane totals -> warp scan -> one atomic reservation per tile ->
//            load-balanced expansion of the descriptors (ballot/scan + binary search by shuffle) so
//            that the id copies out of `values` and into out_ids are coalesced runs.
//
// (Round-1 history: a warp-shared ballot-compacted frontier queue ran at 344 warp-instructions per
//  topic and 45 % issue utilisation — collective overhead, not memory, bound it; see profiles/.)
template <int FAST_L, int THREADS, int CTAS_PER_SM, bool STATS>
__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
k_match_fast(MatchParams p, Desc* __restrict__ dpool, u32 pool_rows) {
    __shared__ u32 s_tok[FAST_L][THREADS];    // tokens of this thread's topic (column = thread: conflict-free)
    __shared__ u32 s_pend[FAST_L][THREADS];   // parked '+' child (index into tv.plus) per depth
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 lt = lanemask_lt();
    const u32 nthreads = gridDim.x * THREADS;
    const u32 gtid = blockIdx.x * THREADS + tid;
    const TrieView& tv = p.tv;
    const u32 ntiles = (p.n + 31) >> 5;
    unsigned long long sV = 0, sE = 0, sF = 0, sM = 0;
    // L2 residency: the upper two levels, every '+' record and everything below a '+' edge that replaced one
    // of the first three levels is shared by many topics of a batch (hot: evict_last); the exact chain below
    // a device is touched by ~1 topic per batch (cold: evict_first) and must not flush the hot set out of L2.
    const bool hints = (p.flags & MP_L2_HINTS) != 0;
    const u64 pol_hot = hints ? l2_policy_evict_last() : 0ull;
    const u64 pol_cold = hints ? l2_policy_evict_first() : 0ull;

    for (;;) {
        u32 tile = 0;
        if (lane == 0) tile = atomicAdd(p.tile_counter, 1u);
        tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
        if (tile >= ntiles) break;

        const u32 t = tile * 32 + lane;
        const bool in_range = t < p.n;
        const u32 m = in_range ? __ldcs(p.meta + t) : META_INVALID;
        const bool invalid = (m & META_INVALID) != 0;
        const u32 L = m & META_NLEV_MASK;
        const u32 need = min(L, tv.max_depth);
        const bool slow_pre = in_range && !invalid && need > FAST_L;
        const bool active = in_range && !invalid && !slow_pre;
        bool defer = slow_pre;
        u32 ndesc = 0, total = 0;

        if (active) {
#pragma unroll
            for (int l = 0; l < FAST_L; ++l)
                if (l < need) s_tok[l][tid] = __ldcs(p.tok + static_cast<size_t>(l) * p.n + t);
            NodeRec r{0u, tv.root_plus, tv.root_hash_ref, 0u, tv.root_mask, tv.root_hash_cnt};
            u32 d = 0, pmask = 0, hotmask = 0;
            bool hot = false;                          // current branch lies below a shared '+' edge
            bool droot = (m & META_DOLLAR) != 0;      // `$`-rule: root wildcards skipped (trie.rs:312-318)
            u32 lV = 0, lE = 0, lF = 0;
            for (;;) {
                if (STATS) { lV++; lE += d < L; }
                // '#' child matches the rest of the path and, on exhaustion, the parent (trie.rs:302-308, 321-327)
                const u32 c1 = droot ? 0u : (r.cnts & 0xFFFFu);
                const u32 c2 = (d == L) ? (r.cnts >> 16) : 0u;                       // own values (trie.rs:309-310)
                if (c1) {
                    if (c1 == CNT_BIG || ndesc >= pool_rows) { defer = true; break; }
                    dpool[static_cast<size_t>(ndesc) * nthreads + gtid] = Desc{r.hash_ref, c1};
                    ++ndesc; total += c1;
                    if (STATS) lF++;
                }
                if (c2) {
                    if (c2 == CNT_BIG || ndesc >= pool_rows) { defer = true; break; }
                    dpool[static_cast<size_t>(ndesc) * nthreads + gtid] = Desc{r.own_ref, c2};
                    ++ndesc; total += c2;
                    if (STATS) lF++;
                }
                bool down = false;
                if (d < L) {
                    if (r.plus != 0 && !droot) {                                                 // '+' child (trie.rs:330-334)
                        s_pend[d][tid] = r.plus; pmask |= 1u << d;
                        if (hot || d <= 2) hotmask |= 1u << d; else hotmask &= ~(1u << d);
                    }
                    if (r.mask != 0) {
                        const u32 tk = s_tok[d][tid];
                        if (tk != TOK_UNKNOWN && (r.mask & mask_bit(tk))) {                      // literal child (trie.rs:338-342)
                            NodeRec c;
                            const bool hit = load_record(tv, KIND_PROBE, r.node, tk, c, (hot || d <= 1) ? pol_hot : pol_cold);
                            if (STATS) {   // diagnostics: probes / misses per depth, slot loads per probe
                                atomicAdd(p.stats + 4 + min(d, 7u), 1ull);
                                if (!hit) atomicAdd(p.stats + 12 + min(d, 7u), 1ull);
                                u32 idx = edge_hash(r.node, tk) & tv.edge_mask, steps = 1;
                                while (tv.edges[idx].child != 0 && !(tv.edges[idx].parent == r.node && tv.edges[idx].token == tk)) { idx = (idx + 1) & tv.edge_mask; ++steps; }
                                atomicAdd(p.stats + 20, static_cast<unsigned long long>(steps));
                            }
                            if (hit) { r = c; ++d; down = true; }
                        }
                    }
                }
                droot = false;
                if (!down) {
                    if (pmask == 0) break;
                    const u32 pd = 31u - __clz(pmask);          // resume the deepest parked '+' child
                    pmask &= ~(1u << pd);
                    hot = (hotmask >> pd) & 1u;
                    load_record(tv, KIND_PLUS, s_pend[pd][tid], 0u, r, hot ? pol_hot : pol_cold);
                    d = pd + 1;
                }
            }
            if (STATS && !defer) { sV += lV; sE += lE; sF += lF; sM += total; }
        }

        // ---- publish: one contiguous list per topic inside one chunk per tile ------------------------
        const u32 mine = (active && !defer) ? total : 0u;
        const u32 nd = (active && !defer) ? ndesc : 0u;
        u32 inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += v; }
        const u32 wtotal = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const u32 pre = inc - mine;
        unsigned long long base = 0;
        if (lane == 0 && wtotal) base = atomicAdd(p.cursor, static_cast<unsigned long long>(wtotal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        const bool fits = base + wtotal <= p.cap_ids;
        if (in_range && !defer) p.spans[t] = make_uint2(fits ? static_cast<u32>(base + pre) : 0u, mine);
        const u32 maxd = __reduce_max_sync(0xFFFFFFFFu, nd);
        if (fits && wtotal) {
            u32* __restrict__ out = p.out_ids + base;
            u32 cur = pre;                                    // this lane's write position inside the tile chunk
            for (u32 k = 0; k < maxd; ++k) {                  // row k: the k-th descriptor of every lane
                Desc dsc = k < nd ? dpool[static_cast<size_t>(k) * nthreads + gtid] : Desc{0u, 0u};
                const u32 ni = dsc.cnt;
                const u32 dst = cur;
                cur += ni;
                u32 sc = ni;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { u32 v = __shfl_up_sync(0xFFFFFFFFu, sc, o); if (lane >= o) sc += v; }
                const u32 tot = __shfl_sync(0xFFFFFFFFu, sc, 31);
                const u32 exc = sc - ni;
                // load-balanced expansion: flat id index e -> (owner lane, k) by binary search over `exc`
                for (u32 e0 = 0; e0 < tot; e0 += 32) {
                    const u32 e = e0 + lane;
                    u32 lo = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        u32 v = __shfl_sync(0xFFFFFFFFu, exc, lo + step);
                        if (v <= e) lo += step;
                    }
                    const u32 o_exc = __shfl_sync(0xFFFFFFFFu, exc, lo);
                    const u32 o_ref = __shfl_sync(0xFFFFFFFFu, dsc.ref, lo);
                    const u32 o_n = __shfl_sync(0xFFFFFFFFu, ni, lo);
                    const u32 o_dst = __shfl_sync(0xFFFFFFFFu, dst, lo);
                    if (e < tot) {
                        const u32 kk = e - o_exc;
                        __stcs(out + o_dst + kk, (o_n == 1) ? o_ref : tv.values[o_ref + kk]);   // streaming store: written once
                    }
                }
            }
        }
        const bool deferred = in_range && defer;
        u32 db = __ballot_sync(0xFFFFFFFFu, deferred);
        if (db) {
            u32 sb = 0;
            if (lane == 0) sb = atomicAdd(p.slow_count, static_cast<u32>(__popc(db)));
            sb = __shfl_sync(0xFFFFFFFFu, sb, 0);
            if (deferred) p.slow_list[sb + __popc(db & lt)] = t;
        }
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            sV += __shfl_xor_sync(0xFFFFFFFFu, sV, o); sE += __shfl_xor_sync(0xFFFFFFFFu, sE, o);
            sF += __shfl_xor_sync(0xFFFFFFFFu, sF, o); sM += __shfl_xor_sync(0xFFFFFFFFu, sM, o);
        }
        if (lane == 0) { atomicAdd(p.stats + 0, sV); atomicAdd(p.stats + 1, sE); atomicAdd(p.stats + 2, sF); atomicAdd(p.stats + 3, sM); }
    }
}

// ------------------------------------------------------------------------------------------------
// K3: one warp per deferred topic; pass 0 counts, pass 1 writes.  The frontier stack lives in global
// scratch (gstack, `stack_cap` items per warp: 32*(max_depth+2)+64 bounds the LIFO walk).
// In STATS mode the counters of a deferred topic are taken here (the fast path's partial counts of
// topics it later deferred are subtracted by never being added: see `stats_defer` below).
template <bool STATS>
__global__ void __launch_bounds__(256)
k_match_slow(MatchParams p, u64* __restrict__ gstack, u32 stack_cap) {
    const u32 lane = threadIdx.x & 31;
    const u32 lt = lanemask_lt();
    const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    const TrieView& tv = p.tv;
    u64* stack = gstack + static_cast<size_t>(gwarp) * stack_cap;
    const u32 nslow = *p.slow_count;
    unsigned long long sV = 0, sE = 0, sF = 0, sM = 0;

    for (u32 si = gwarp; si < nslow; si += nwarps) {
        const u32 t = p.slow_list[si];
        const u32 m = p.meta[t];
        const u32 L = m & META_NLEV_MASK;
        const bool dollar = (m & META_DOLLAR) != 0;
        unsigned long long base = 0;
        unsigned long long count = 0;
        bool fits = true;
        for (int pass = 0; pass < 2; ++pass) {
            u32 stack_n = 0;
            unsigned long long written = 0;
            bool bad = false;
            auto consume = [&](bool hit, u32 d, const NodeRec& r, bool dollar_root) {
                const u32 c1 = (hit && !dollar_root) ? (r.cnts & 0xFFFFu) : 0u;
                const u32 c2 = (hit && d == L) ? (r.cnts >> 16) : 0u;
                u64 itA = 0, itB = 0;
                bool pA = false, pB = false;
                if (hit && d < L) {
                    if (r.plus != 0 && !dollar_root) { pA = true; itA = make_item(0, d + 1, KIND_PLUS, r.plus); }
                    if (r.mask != 0 && d < p.tok_levels) {
                        u32 tk = p.tok[static_cast<size_t>(d) * p.n + t];
                        if (tk != TOK_UNKNOWN && (r.mask & mask_bit(tk))) { pB = true; itB = make_item(0, d + 1, KIND_PROBE, r.node); }
                    }
                }
                if (STATS && pass == 0) { sV += hit; sE += (hit && d < L); }
#pragma unroll
                for (int round = 0; round < 2; ++round) {
                    const u32 cn = round == 0 ? c1 : c2;
                    const u32 ref = round == 0 ? r.hash_ref : r.own_ref;
                    if (STATS && pass == 0) sF += (cn != 0);
                    u32 b = __ballot_sync(0xFFFFFFFFu, cn == 1);
                    if (pass == 1 && cn == 1) p.out_ids[base + written + __popc(b & lt)] = ref;
                    written += __popc(b);
                    u32 rb = __ballot_sync(0xFFFFFFFFu, cn > 1);
                    while (rb) {     // warp-cooperative copy of one value set at a time
                        int leader = __ffs(rb) - 1;
                        rb &= rb - 1;
                        u32 rr = __shfl_sync(0xFFFFFFFFu, ref, leader);
                        u32 rc = __shfl_sync(0xFFFFFFFFu, cn, leader);
                        u32 off = rr;
                        if (rc == CNT_BIG) { Range rg = tv.ranges[rr]; off = rg.off; rc = rg.cnt; }
                        if (pass == 1)
                            for (u32 i = lane; i < rc; i += 32) p.out_ids[base + written + i] = tv.values[off + i];
                        written += rc;
                    }
                }
#pragma unroll
                for (int round = 0; round < 2; ++round) {
                    const bool push = round == 0 ? pA : pB;
                    u32 b = __ballot_sync(0xFFFFFFFFu, push);
                    if (b) {
                        u32 tot = __popc(b);
                        if (stack_n + tot <= stack_cap) {
                            if (push) stack[stack_n + __popc(b & lt)] = round == 0 ? itA : itB;
                            stack_n += tot;
                        } else bad = true;
                    }
                }
                __syncwarp();
            };
            {
                NodeRec r{0u, tv.root_plus, tv.root_hash_ref, 0u, tv.root_mask, tv.root_hash_cnt};
                consume(lane == 0, 0u, r, dollar);
            }
            while (stack_n) {
                const u32 take = min(stack_n, 32u);
                const bool have = lane < take;
                __threadfence_block();
                u64 it = have ? stack[stack_n - 1 - lane] : 0ull;
                stack_n -= take;
                __syncwarp();
                const u32 hi = static_cast<u32>(it >> 32), arg = static_cast<u32>(it);
                const u32 d = (hi >> 5) & 0xFFFFu, kind = (hi >> 21) & 1u;
                NodeRec r{};
                bool hit = false;
                if (have) {
                    u32 tk = kind == KIND_PROBE ? p.tok[static_cast<size_t>(d - 1) * p.n + t] : 0u;
                    hit = load_record(tv, kind, arg, tk, r);
                }
                consume(hit, d, r, false);
            }
            if (pass == 0) {
                count = written;
                if (STATS) sM += (lane == 0) ? count : 0;
                if (lane == 0 && count) base = atomicAdd(p.cursor, count);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                fits = !bad && (base + count <= p.cap_ids) && count <= 0xFFFFFFFFull;
                if (!fits) break;
            }
        }
        if (lane == 0) p.spans[t] = make_uint2(fits ? static_cast<u32>(base) : 0u, static_cast<u32>(count));
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            sV += __shfl_xor_sync(0xFFFFFFFFu, sV, o); sE += __shfl_xor_sync(0xFFFFFFFFu, sE, o);
            sF += __shfl_xor_sync(0xFFFFFFFFu, sF, o); sM += __shfl_xor_sync(0xFFFFFFFFu, sM, o);
        }
        if (lane == 0) { atomicAdd(p.stats + 0, sV); atomicAdd(p.stats + 1, sE); atomicAdd(p.stats + 2, sF); atomicAdd(p.stats + 3, sM); }
    }
}

// ------------------------------------------------------------------------------------------------
// Flush: scatter slot patches (EdgeSlot / DictSlot / PlusRec: 32 B, Range: 8 B) into a device table.
template <class T>
__global__ void k_apply_patches(T* __restrict__ table, const u32* __restrict__ idx, const T* __restrict__ data, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    table[idx[i]] = data[i];
}

}  // namespace gm
