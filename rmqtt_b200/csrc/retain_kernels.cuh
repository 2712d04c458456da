// sm_100a kernels of the SUBSCRIBE -> retained-message lookup (RetainTree::matches,
// rmqtt/src/retain.rs:291-367) for a batch of topic filters.
//
// A query is a FILTER, so its work is data dependent: an exact filter touches one path, `reg/+/+/...`
// fans out to tens of thousands of nodes.  The batch is therefore walked as ONE level-synchronous
// breadth-first frontier over all queries: an item is (query, node, position in the filter); every
// step kernel moves the whole frontier one tree level down.  '+' fan-out is expanded load-balanced
// across the warp (scan + binary search by shuffle) from the node's contiguous child block; '#' is a
// single contiguous range of the pre-order value array (retain_tree.h).  Matched values are recorded as
// descriptors and expanded into per-query contiguous lists by k_retain_expand.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "retain_tree.h"

namespace gm {

struct alignas(32) RItem { u32 q, pos, node, first_kid, nk_flags, val, val_lo, val_hi; };   // the node's record travels with the item
struct alignas(16) RDesc { u32 q, ref, cnt, kind; };   // kind 0: the value itself, 1: rvals[ref .. ref+cnt)

// Work queues (frontier, descriptor list) are split into RQ slices with one counter each: with a single bump
// counter, the ~10^7 warp-aggregated appends of a C4 batch serialise on one L2 atomic unit.
constexpr u32 RQ = 64;

struct RetainParams {
    RetainView v;
    const u32* qtok8;    // [nq][8]  filter levels 0..7
    const u32* qtok;     // [tok_levels][nq]  levels >= 8
    const u32* qmeta;    // [nq]
    u32 nq, tok_levels;
    RDesc* descs;
    u32* n_desc;         // [RQ]
    u32 cap_items, cap_desc;   // per SLICE
    u32* qtotal;         // [nq] matched values per query
    u32* err;            // bit 0: frontier overflow, bit 1: descriptor overflow
};

__device__ __forceinline__ u32 retain_tok(const RetainParams& p, u32 q, u32 pos) {
    return pos < TOK8 ? p.qtok8[static_cast<size_t>(q) * TOK8 + pos] : p.qtok[static_cast<size_t>(pos) * p.nq + q];
}

// exact child lookup: the slot carries the child's whole record
__device__ __forceinline__ bool retain_child(const RetainView& v, u32 node, u32 token, u32 (&s)[8]) {
    u32 idx = redge_hash(node, token) & v.edge_mask;
    for (;;) {
        ld256(v.edges + idx, s);
        if (s[2] == 0) return false;
        if (s[0] == node && s[1] == token) return true;
        idx = (idx + 1) & v.edge_mask;
    }
}

// warp-aggregated append of one element per flagged lane into slice `sq` of a sliced queue
template <class T>
__device__ __forceinline__ void warp_append(bool want, const T& item, T* arr, u32* counters, u32 slice_cap, u32 sq, u32* err, u32 errbit, u32 lane, u32 lt) {
    const u32 b = __ballot_sync(0xFFFFFFFFu, want);
    if (!b) return;
    u32 base = 0;
    if (lane == static_cast<u32>(__ffs(b) - 1)) base = atomicAdd(counters + sq, static_cast<u32>(__popc(b)));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(b) - 1);
    if (want) {
        const u32 at = base + __popc(b & lt);
        if (at < slice_cap) arr[static_cast<size_t>(sq) * slice_cap + at] = item; else atomicOr(err, errbit);
    }
}

// block-wide: exclusive prefix of the (clamped) slice counts into s_pre[RQ + 1]; returns the total
__device__ __forceinline__ u32 queue_prefix(const u32* __restrict__ counters, u32 slice_cap, u32* s_pre) {
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (u32 k = 0; k < RQ; ++k) { s_pre[k] = run; run += min(counters[k], slice_cap); }
        s_pre[RQ] = run;
    }
    __syncthreads();
    return s_pre[RQ];
}
// logical index g -> physical index inside the sliced array
__device__ __forceinline__ size_t queue_locate(const u32* s_pre, u32 g, u32 slice_cap) {
    u32 lo = 0;
#pragma unroll
    for (u32 step = RQ / 2; step; step >>= 1) if (s_pre[lo + step] <= g) lo += step;
    return static_cast<size_t>(lo) * slice_cap + (g - s_pre[lo]);
}

__global__ void k_retain_init(RetainParams p, RItem* out, u32* n_out) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 lane = threadIdx.x & 31, lt = lanemask_lt();
    const bool ok = q < p.nq && !(p.qmeta[q] & META_INVALID);
    warp_append(ok, RItem{q, 0u, 0u, p.v.root_first_kid, p.v.root_nk_flags, 0u, 0u, p.v.root_plain_val_hi}, out, n_out, p.cap_items, (q >> 5) % RQ, p.err, 1u, lane, lt);
}

__global__ void __launch_bounds__(256)
k_retain_step(RetainParams p, const RItem* __restrict__ in, const u32* __restrict__ n_in_p, RItem* __restrict__ out, u32* n_out) {
    const u32 lane = threadIdx.x & 31, lt = lanemask_lt();
    const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const RetainView& v = p.v;
    __shared__ u32 s_pre[RQ + 1];
    const u32 n_in = queue_prefix(n_in_p, p.cap_items, s_pre);
    const u32 sq = gwarp % RQ;                        // this warp appends to its own slice
    for (u32 base = gwarp * 32; base < n_in; base += nwarps * 32) {
        const u32 i = base + lane;
        const bool have = i < n_in;
        RItem it{};
        if (have) { u32 w[8]; ld256(in + queue_locate(s_pre, i, p.cap_items), w); it = RItem{w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]}; }
        u32 exp_kb = 0, exp_n = 0, exp_mode = 0;      // child-block expansion request: 1 = '+', 2 = '#' level by level
        bool push1 = false, emit1 = false, next_hash = false;
        RItem pitem{};
        RDesc d1{};
        u32 L = 0;
        if (have) {
            L = p.qmeta[it.q] & META_NLEV_MASK;
            const u32 nkids = it.nk_flags & RNK_MASK, flags = it.nk_flags >> 28;
            if (nkids == 0 || it.pos == L) {                                     // retain.rs:305-311
                if (it.pos == L && (flags & 8u)) { emit1 = true; d1 = RDesc{it.q, it.val, 1u, 0u}; }
            } else {
                const u32 tok = retain_tok(p, it.q, it.pos);
                next_hash = (it.pos + 1 < L) && retain_tok(p, it.q, it.pos + 1) == TOK_HASH;
                // precise matching first — Level equality, so a stored literal "+" / "#" child shadows the
                // wildcard expansion (retain.rs:313)
                const bool exact_try = tok >= TOK_BLANK || (tok == TOK_PLUS && (flags & RF_LIT_PLUS)) || (tok == TOK_HASH && (flags & RF_LIT_HASH));
                u32 c[8];
                const bool found = exact_try && retain_child(v, it.node, tok, c);
                const bool root = it.node == 0;
                if (found) {
                    if (next_hash && ((c[4] >> 28) & 8u)) { emit1 = true; d1 = RDesc{it.q, c[5], 1u, 0u}; }   // '#' matches the parent, retain.rs:317-322
                    push1 = true; pitem = RItem{it.q, it.pos + 1, c[2], c[3], c[4], c[5], c[6], c[7]};
                } else if (tok == TOK_PLUS) {                                    // retain.rs:324-342
                    exp_kb = it.first_kid; exp_n = root ? v.root_plain_kids : nkids; exp_mode = 1;
                } else if (tok == TOK_HASH) {                                    // retain.rs:343-365
                    if (!(flags & RF_SUB_LIT_HASH)) {       // every strict descendant (minus `$` subtrees at the root): one range
                        const u32 lo = it.val_lo + ((flags & 8u) ? 1u : 0u), hi = root ? v.root_plain_val_hi : it.val_hi;
                        if (hi > lo) { emit1 = true; d1 = RDesc{it.q, lo, hi - lo, 1u}; }
                    } else {
                        exp_kb = it.first_kid; exp_n = root ? v.root_plain_kids : nkids; exp_mode = 2;
                    }
                }
            }
        }
        if (emit1) atomicAdd(p.qtotal + d1.q, d1.cnt);
        warp_append(emit1, d1, p.descs, p.n_desc, p.cap_desc, sq, p.err, 2u, lane, lt);
        warp_append(push1, pitem, out, n_out, p.cap_items, sq, p.err, 1u, lane, lt);

        // ---- load-balanced expansion of the requested child blocks
        u32 sc = exp_n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(0xFFFFFFFFu, sc, o); if (lane >= o) sc += x; }
        const u32 tot = __shfl_sync(0xFFFFFFFFu, sc, 31);
        const u32 exc = sc - exp_n;
        for (u32 e0 = 0; e0 < tot; e0 += 32) {
            const u32 e = e0 + lane;
            u32 lo = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                u32 x = __shfl_sync(0xFFFFFFFFu, exc, lo + step);
                if (x <= e) lo += step;
            }
            const u32 o_exc = __shfl_sync(0xFFFFFFFFu, exc, lo);
            const u32 o_kb = __shfl_sync(0xFFFFFFFFu, exp_kb, lo);
            const u32 o_mode = __shfl_sync(0xFFFFFFFFu, exp_mode, lo);
            const u32 o_q = __shfl_sync(0xFFFFFFFFu, it.q, lo);
            const u32 o_pos = __shfl_sync(0xFFFFFFFFu, it.pos, lo);
            const u32 o_L = __shfl_sync(0xFFFFFFFFu, L, lo);
            const bool o_nh = __shfl_sync(0xFFFFFFFFu, next_hash ? 1u : 0u, lo) != 0;
            bool em = false, pu = false;
            RDesc d{};
            RItem ni{};
            if (e < tot) {
                u32 kd[8];                                              // {token, child, first_kid, nk_flags, val, val_lo, val_hi, pad}
                ld256(v.kids + o_kb + (e - o_exc), kd);
                const bool has_val = ((kd[3] >> 28) & 8u) != 0;
                const u32 kn = kd[3] & RNK_MASK;
                if (o_mode == 1) {
                    if (o_pos + 1 == o_L) em = has_val;                 // filter ends here: the child's own value
                    else { em = o_nh && has_val; pu = kn > 0; }         // `.../+/#` parent match; descend only if the child has branches
                    ni = RItem{o_q, o_pos + 1, kd[1], kd[2], kd[3], kd[4], kd[5], kd[6]};
                } else {                                                // '#' one level at a time (a literal "#" child hides below)
                    em = has_val; pu = kn > 0;
                    ni = RItem{o_q, o_pos, kd[1], kd[2], kd[3], kd[4], kd[5], kd[6]};
                }
                d = RDesc{o_q, kd[4], 1u, 0u};
            }
            if (em) atomicAdd(p.qtotal + d.q, 1u);
            warp_append(em, d, p.descs, p.n_desc, p.cap_desc, sq, p.err, 2u, lane, lt);
            warp_append(pu, ni, out, n_out, p.cap_items, sq, p.err, 1u, lane, lt);
        }
    }
}

// Exclusive scan of qtotal -> qbase, spans; single block (nq is a batch of SUBSCRIBE filters).
__global__ void __launch_bounds__(1024)
k_retain_scan(const u32* __restrict__ qtotal, u32 nq, u32* __restrict__ qbase, uint2* __restrict__ spans, unsigned long long* grand) {
    __shared__ unsigned long long s_part[1024];
    const u32 tid = threadIdx.x;
    const u32 per = (nq + 1023) / 1024;
    const u32 b = min(nq, tid * per), e = min(nq, b + per);
    unsigned long long sum = 0;
    for (u32 i = b; i < e; ++i) sum += qtotal[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) { unsigned long long x = s_part[i]; s_part[i] = run; run += x; }
        *grand = run;
    }
    __syncthreads();
    unsigned long long run = s_part[tid];
    for (u32 i = b; i < e; ++i) {
        const u32 c = qtotal[i];
        const u32 off = run > 0xFFFFFFFFull ? 0xFFFFFFFFu : static_cast<u32>(run);
        qbase[i] = off;
        spans[i] = make_uint2(off, c);
        run += c;
    }
}

// Expands the descriptors into per-query contiguous lists (load-balanced like the publish step of k_match_fast).
__global__ void __launch_bounds__(256)
k_retain_expand(const RDesc* __restrict__ descs, const u32* __restrict__ n_desc_p, u32 cap_desc, const u32* __restrict__ rvals,
                const u32* __restrict__ qbase, u32* __restrict__ qcur, u32* __restrict__ out, unsigned long long cap_ids) {
    const u32 lane = threadIdx.x & 31;
    const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    __shared__ u32 s_pre[RQ + 1];
    const u32 nd = queue_prefix(n_desc_p, cap_desc, s_pre);
    for (u32 base = gwarp * 32; base < nd; base += nwarps * 32) {
        const u32 i = base + lane;
        RDesc d = i < nd ? descs[queue_locate(s_pre, i, cap_desc)] : RDesc{0, 0, 0, 0};
        const u32 ni = d.cnt;
        const u32 dst = ni ? qbase[d.q] + atomicAdd(qcur + d.q, ni) : 0u;
        u32 sc = ni;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 x = __shfl_up_sync(0xFFFFFFFFu, sc, o); if (lane >= o) sc += x; }
        const u32 tot = __shfl_sync(0xFFFFFFFFu, sc, 31);
        const u32 exc = sc - ni;
        for (u32 e0 = 0; e0 < tot; e0 += 32) {
            const u32 e = e0 + lane;
            u32 lo = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                u32 x = __shfl_sync(0xFFFFFFFFu, exc, lo + step);
                if (x <= e) lo += step;
            }
            const u32 o_exc = __shfl_sync(0xFFFFFFFFu, exc, lo);
            const u32 o_ref = __shfl_sync(0xFFFFFFFFu, d.ref, lo);
            const u32 o_kind = __shfl_sync(0xFFFFFFFFu, d.kind, lo);
            const u32 o_dst = __shfl_sync(0xFFFFFFFFu, dst, lo);
            if (e < tot) {
                const u32 k = e - o_exc;
                const unsigned long long at = static_cast<unsigned long long>(o_dst) + k;
                if (at < cap_ids) __stcs(out + at, o_kind ? rvals[o_ref + k] : o_ref);
            }
        }
    }
}

}  // namespace gm
