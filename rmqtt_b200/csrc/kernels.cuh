// sm_100a kernels of the PUBLISH -> matching-subscribers hot path.
//
//   k_tokenize     Topic::from_str (rmqtt/src/topic.rs:326-363) for a batch: split on '/', classify,
//                  validate, and intern every level through the device dictionary -> u32 tokens.
//   k_match_fast   TopicTree::matches (rmqtt/src/trie.rs:299-347, MatchedIter::prepare) for 32 topics
//                  per warp: a warp-shared LIFO frontier of "load one node record" work items, one
//                  256-bit load per item, warp-ballot compaction of the new frontier items and of the
//                  matched subscriber ids, single pass, per-topic contiguous output.
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
    unsigned long long* stats;    // [4] V,E,F,M  (only written by STATS instantiations)
};

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld256(const void* p, u32 (&w)[8]) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
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
    status[t] = invalid ? 1 : 0;
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
__device__ __forceinline__ bool load_record(const TrieView& tv, u32 kind, u32 arg, u32 token, NodeRec& r) {
    u32 s[8];
    if (kind == KIND_PLUS) {
        ld256(tv.plus + arg, s);
        r.node = s[0]; r.plus = s[1]; r.hash_ref = s[2]; r.own_ref = s[3]; r.mask = s[4]; r.cnts = s[5];
        return true;
    }
    u32 idx = edge_hash(arg, token) & tv.edge_mask;
    for (;;) {
        ld256(tv.edges + idx, s);
        if (s[2] == 0) return false;                       // empty slot: no such child
        if (s[0] == arg && s[1] == token) break;
        idx = (idx + 1) & tv.edge_mask;
    }
    r.node = s[2]; r.plus = s[3]; r.hash_ref = s[4]; r.own_ref = s[5]; r.mask = s[6]; r.cnts = s[7];
    return true;
}

// A matched value set waiting to be expanded into the output: ref + (cnt16 | topic slot << 16).
struct Desc { u32 ref, meta; };

template <int FAST_L, int STACK_CAP>
struct alignas(16) WarpSmem {
    u64 stack[STACK_CAP];
    u32 tok[FAST_L][32];
    u32 cnt[32];
    u32 cur[32];
    u32 nlev[32];
    u32 st[3][32];     // STATS instantiation only: per-topic V, E, F (so deferred topics are not double counted)
};

// K2: 32 topics per warp.  Matched value sets are staged as 8-byte descriptors in a per-warp slice of
// a global scratch pool (written once, read twice, L2-resident), so that shared memory only holds
// the frontier stack and the tokens and many warps fit on an SM (the walk is latency bound).
template <int FAST_L, int STACK_CAP, int WARPS, int CTAS_PER_SM, bool STATS>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM)
k_match_fast(MatchParams p, Desc* __restrict__ gpool, u32 pool_cap) {
    using WS = WarpSmem<FAST_L, STACK_CAP>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WS& W = reinterpret_cast<WS*>(smem_raw)[threadIdx.x >> 5];
    const u32 lane = threadIdx.x & 31;
    const u32 lt = lanemask_lt();
    const TrieView& tv = p.tv;
    const u32 ntiles = (p.n + 31) >> 5;
    Desc* pool = gpool + static_cast<size_t>(blockIdx.x * WARPS + (threadIdx.x >> 5)) * pool_cap;
    unsigned long long sV = 0, sE = 0, sF = 0, sM = 0;

    for (;;) {
        u32 tile = 0;
        if (lane == 0) tile = atomicAdd(p.tile_counter, 1u);
        tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
        if (tile >= ntiles) break;

        const u32 t = tile * 32 + lane;
        const bool in_range = t < p.n;
        const u32 m = in_range ? p.meta[t] : META_INVALID;
        const bool invalid = (m & META_INVALID) != 0;
        const u32 L = m & META_NLEV_MASK;
        const u32 need = min(L, tv.max_depth);
        const bool slow_pre = in_range && !invalid && need > FAST_L;
        const bool active = in_range && !invalid && !slow_pre;
        W.nlev[lane] = L;
        if (STATS) { W.st[0][lane] = 0; W.st[1][lane] = 0; W.st[2][lane] = 0; }
        if (active) {
#pragma unroll
            for (int l = 0; l < FAST_L; ++l)
                if (l < need) W.tok[l][lane] = p.tok[static_cast<size_t>(l) * p.n + t];
        }
        __syncwarp();

        u32 pool_n = 0, stack_n = 0;   // warp-uniform
        u32 ovf = 0;                   // warp-uniform bit mask of topic slots deferred to the slow path

        // Consumes one loaded record per lane (hit == false: lane idle).  Warp-collective.
        auto consume = [&](bool hit, u32 topic, u32 d, u32 Lt, const NodeRec& r, bool dollar_root) {
            // `#` child: matches the rest of the path (trie.rs:321-327) and, on path exhaustion, the
            // parent itself (trie.rs:302-308).  Skipped at the root for `$`-topics (trie.rs:312-318).
            const u32 c1 = (hit && !dollar_root) ? (r.cnts & 0xFFFFu) : 0u;
            const u32 c2 = (hit && d == Lt) ? (r.cnts >> 16) : 0u;              // own values, trie.rs:309-310
            u64 itA = 0, itB = 0;
            bool pA = false, pB = false;
            if (hit && d < Lt) {
                if (r.plus != 0 && !dollar_root) { pA = true; itA = make_item(topic, d + 1, KIND_PLUS, r.plus); }   // trie.rs:330-334
                if (r.mask != 0) {
                    u32 tk = W.tok[d][topic];
                    if (tk != TOK_UNKNOWN && (r.mask & mask_bit(tk))) { pB = true; itB = make_item(topic, d + 1, KIND_PROBE, r.node); }  // trie.rs:338-342
                }
            }
            if (STATS && hit) { atomicAdd(&W.st[0][topic], 1u); if (d < Lt) atomicAdd(&W.st[1][topic], 1u); }
            // ---- matched value sets -> descriptor pool (ballot compaction)
#pragma unroll
            for (int round = 0; round < 2; ++round) {
                const u32 cn = round == 0 ? c1 : c2;
                const u32 ref = round == 0 ? r.hash_ref : r.own_ref;
                const bool some = cn != 0;
                const bool big = cn == CNT_BIG;              // >= 65535 values in one set: deferred path
                u32 b = __ballot_sync(0xFFFFFFFFu, some);
                if (b) {
                    u32 tot = __popc(b);
                    const bool room = pool_n + tot <= pool_cap;
                    if (room) {
                        if (some) pool[pool_n + __popc(b & lt)] = Desc{ref, cn | (topic << 16)};
                        pool_n += tot;
                    }
                    u32 bad = __ballot_sync(0xFFFFFFFFu, some && (big || !room));
                    if (bad) ovf |= __reduce_or_sync(0xFFFFFFFFu, (some && (big || !room)) ? (1u << topic) : 0u);
                }
            }
            // ---- new frontier items -> LIFO stack (ballot compaction)
#pragma unroll
            for (int round = 0; round < 2; ++round) {
                const bool push = round == 0 ? pA : pB;
                u32 b = __ballot_sync(0xFFFFFFFFu, push);
                if (b) {
                    u32 tot = __popc(b);
                    if (stack_n + tot <= STACK_CAP) {
                        if (push) W.stack[stack_n + __popc(b & lt)] = round == 0 ? itA : itB;
                        stack_n += tot;
                    } else {
                        ovf |= __reduce_or_sync(0xFFFFFFFFu, push ? (1u << topic) : 0u);
                    }
                }
            }
            __syncwarp();
        };

        // root of every topic (depth 0): the record comes from the kernel parameters
        {
            NodeRec r{0u, tv.root_plus, tv.root_hash_ref, 0u, tv.root_mask, tv.root_hash_cnt};
            consume(active, lane, 0u, L, r, active && (m & META_DOLLAR) != 0);
        }

        while (stack_n) {
            const u32 take = min(stack_n, 32u);
            bool have = lane < take;
            u64 it = have ? W.stack[stack_n - 1 - lane] : 0ull;
            stack_n -= take;
            __syncwarp();
            const u32 hi = static_cast<u32>(it >> 32), arg = static_cast<u32>(it);
            const u32 topic = hi & 31u, d = (hi >> 5) & 0xFFFFu, kind = (hi >> 21) & 1u;
            have = have && !((ovf >> topic) & 1u);
            NodeRec r{};
            bool hit = false;
            u32 Lt = 0;
            if (have) {
                Lt = W.nlev[topic];
                u32 tk = kind == KIND_PROBE ? W.tok[d - 1][topic] : 0u;
                hit = load_record(tv, kind, arg, tk, r);
            }
            consume(hit, topic, d, Lt, r, false);
        }

        // ---- publish: one contiguous list per topic inside one chunk per tile
        __threadfence_block();   // descriptors written by other lanes of this warp
        W.cnt[lane] = 0;
        __syncwarp();
        for (u32 i = lane; i < pool_n; i += 32) {
            Desc dsc = pool[i];
            u32 tg = (dsc.meta >> 16) & 31u;
            if (!((ovf >> tg) & 1u)) { atomicAdd(&W.cnt[tg], dsc.meta & 0xFFFFu); if (STATS) atomicAdd(&W.st[2][tg], 1u); }
        }
        __syncwarp();
        const u32 c = W.cnt[lane];
        u32 inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 v = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += v; }
        const u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const u32 pre = inc - c;
        unsigned long long base = 0;
        if (lane == 0 && total) base = atomicAdd(p.cursor, static_cast<unsigned long long>(total));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        const bool fits = base + total <= p.cap_ids;
        W.cur[lane] = pre;
        __syncwarp();
        if (fits && total) {
            for (u32 i0 = 0; i0 < pool_n; i0 += 32) {        // 32 descriptors at a time, one per lane
                const u32 i = i0 + lane;
                Desc dsc = i < pool_n ? pool[i] : Desc{0u, 0u};
                const u32 tg = (dsc.meta >> 16) & 31u;
                const u32 ni = (i < pool_n && !((ovf >> tg) & 1u)) ? (dsc.meta & 0xFFFFu) : 0u;
                const u32 dst = ni ? atomicAdd(&W.cur[tg], ni) : 0u;      // position inside the tile chunk
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
                        const u32 k = e - o_exc;
                        p.out_ids[base + o_dst + k] = (o_n == 1) ? o_ref : tv.values[o_ref + k];
                    }
                }
            }
        }
        const bool deferred = slow_pre || (active && ((ovf >> lane) & 1u));
        if (in_range && !deferred) p.spans[t] = make_uint2(fits ? static_cast<u32>(base + pre) : 0u, c);
        if (STATS && active && !deferred) { sV += W.st[0][lane]; sE += W.st[1][lane]; sF += W.st[2][lane]; sM += c; }
        u32 db = __ballot_sync(0xFFFFFFFFu, deferred);
        if (db) {
            u32 sb = 0;
            if (lane == 0) sb = atomicAdd(p.slow_count, static_cast<u32>(__popc(db)));
            sb = __shfl_sync(0xFFFFFFFFu, sb, 0);
            if (deferred) p.slow_list[sb + __popc(db & lt)] = t;
        }
        __syncwarp();
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
