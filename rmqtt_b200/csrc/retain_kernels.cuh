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

struct alignas(16) RItem { u32 q, node, pos, pad; };
struct alignas(16) RDesc { u32 q, ref, cnt, kind; };   // kind 0: the value itself, 1: rvals[ref .. ref+cnt)

struct RetainParams {
    RetainView v;
    const u32* qtok8;    // [nq][8]  filter levels 0..7
    const u32* qtok;     // [tok_levels][nq]  levels >= 8
    const u32* qmeta;    // [nq]
    u32 nq, tok_levels;
    RDesc* descs;
    u32* n_desc;
    u32 cap_items, cap_desc;
    u32* qtotal;         // [nq] matched values per query
    u32* err;            // bit 0: frontier overflow, bit 1: descriptor overflow
};

__device__ __forceinline__ u32 retain_tok(const RetainParams& p, u32 q, u32 pos) {
    return pos < TOK8 ? p.qtok8[static_cast<size_t>(q) * TOK8 + pos] : p.qtok[static_cast<size_t>(pos) * p.nq + q];
}

__device__ __forceinline__ u32 retain_child(const RetainView& v, u32 node, u32 token) {
    u32 idx = redge_hash(node, token) & v.edge_mask;
    for (;;) {
        const uint4 s = __ldg(reinterpret_cast<const uint4*>(v.edges + idx));
        if (s.z == 0) return 0u;
        if (s.x == node && s.y == token) return s.z;
        idx = (idx + 1) & v.edge_mask;
    }
}

// warp-aggregated append of one element per flagged lane
template <class T>
__device__ __forceinline__ void warp_append(bool want, const T& item, T* arr, u32* counter, u32 cap, u32* err, u32 errbit, u32 lane, u32 lt) {
    const u32 b = __ballot_sync(0xFFFFFFFFu, want);
    if (!b) return;
    u32 base = 0;
    if (lane == static_cast<u32>(__ffs(b) - 1)) base = atomicAdd(counter, static_cast<u32>(__popc(b)));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(b) - 1);
    if (want) {
        const u32 at = base + __popc(b & lt);
        if (at < cap) arr[at] = item; else atomicOr(err, errbit);
    }
}

__global__ void k_retain_init(RetainParams p, RItem* out, u32* n_out) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 lane = threadIdx.x & 31, lt = lanemask_lt();
    const bool ok = q < p.nq && !(p.qmeta[q] & META_INVALID);
    warp_append(ok, RItem{q, 0u, 0u, 0u}, out, n_out, p.cap_items, p.err, 1u, lane, lt);
}

__global__ void __launch_bounds__(256)
k_retain_step(RetainParams p, const RItem* __restrict__ in, const u32* __restrict__ n_in_p, RItem* __restrict__ out, u32* n_out) {
    const u32 lane = threadIdx.x & 31, lt = lanemask_lt();
    const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const RetainView& v = p.v;
    const u32 n_in = min(*n_in_p, p.cap_items);
    for (u32 base = gwarp * 32; base < n_in; base += nwarps * 32) {
        const u32 i = base + lane;
        const bool have = i < n_in;
        RItem it = have ? in[i] : RItem{0, 0, 0, 0};
        u32 exp_kb = 0, exp_n = 0, exp_mode = 0;      // child-block expansion request: 1 = '+', 2 = '#' level by level
        bool push1 = false, emit1 = false, next_hash = false;
        RItem pitem{};
        RDesc d1{};
        u32 L = 0;
        if (have) {
            L = p.qmeta[it.q] & META_NLEV_MASK;
            u32 w[8];
            ld256(v.nodes + it.node, w);
            const u32 first_kid = w[0], nkids = w[1], val = w[2], val_lo = w[3], val_hi = w[4], flags = w[5];
            if (nkids == 0 || it.pos == L) {                                     // retain.rs:305-311
                if (it.pos == L && (flags & 8u)) { emit1 = true; d1 = RDesc{it.q, val, 1u, 0u}; }
            } else {
                const u32 tok = retain_tok(p, it.q, it.pos);
                next_hash = (it.pos + 1 < L) && retain_tok(p, it.q, it.pos + 1) == TOK_HASH;
                // precise matching first — Level equality, so a stored literal "+" / "#" child shadows the
                // wildcard expansion (retain.rs:313)
                const bool exact_try = tok >= TOK_BLANK || (tok == TOK_PLUS && (flags & RF_LIT_PLUS)) || (tok == TOK_HASH && (flags & RF_LIT_HASH));
                const u32 child = exact_try ? retain_child(v, it.node, tok) : 0u;
                const bool root = it.node == 0;
                if (child) {
                    if (next_hash) {                                             // '#' matches the parent, retain.rs:317-322
                        u32 c[8];
                        ld256(v.nodes + child, c);
                        if (c[5] & 8u) { emit1 = true; d1 = RDesc{it.q, c[2], 1u, 0u}; }
                    }
                    push1 = true; pitem = RItem{it.q, child, it.pos + 1, 0u};
                } else if (tok == TOK_PLUS) {                                    // retain.rs:324-342
                    exp_kb = first_kid; exp_n = root ? v.root_plain_kids : nkids; exp_mode = 1;
                } else if (tok == TOK_HASH) {                                    // retain.rs:343-365
                    if (!(flags & RF_SUB_LIT_HASH)) {       // every strict descendant (minus `$` subtrees at the root): one range
                        const u32 lo = val_lo + ((flags & 8u) ? 1u : 0u), hi = root ? v.root_plain_val_hi : val_hi;
                        if (hi > lo) { emit1 = true; d1 = RDesc{it.q, lo, hi - lo, 1u}; }
                    } else {
                        exp_kb = first_kid; exp_n = root ? v.root_plain_kids : nkids; exp_mode = 2;
                    }
                }
            }
        }
        if (emit1) atomicAdd(p.qtotal + d1.q, d1.cnt);
        warp_append(emit1, d1, p.descs, p.n_desc, p.cap_desc, p.err, 2u, lane, lt);
        warp_append(push1, pitem, out, n_out, p.cap_items, p.err, 1u, lane, lt);

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
                const uint4 kd = __ldg(reinterpret_cast<const uint4*>(v.kids + o_kb + (e - o_exc)));   // {token, child, val, nkids|has<<31}
                const bool has_val = (kd.w >> 31) != 0;
                const u32 kn = kd.w & 0x7FFFFFFFu;
                if (o_mode == 1) {
                    if (o_pos + 1 == o_L) em = has_val;                 // filter ends here: the child's own value
                    else { em = o_nh && has_val; pu = kn > 0; }         // `.../+/#` parent match; descend only if the child has branches
                    ni = RItem{o_q, kd.y, o_pos + 1, 0u};
                } else {                                                // '#' one level at a time (a literal "#" child hides below)
                    em = has_val; pu = kn > 0;
                    ni = RItem{o_q, kd.y, o_pos, 0u};
                }
                d = RDesc{o_q, kd.z, 1u, 0u};
            }
            if (em) atomicAdd(p.qtotal + d.q, 1u);
            warp_append(em, d, p.descs, p.n_desc, p.cap_desc, p.err, 2u, lane, lt);
            warp_append(pu, ni, out, n_out, p.cap_items, p.err, 1u, lane, lt);
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
    const u32 nd = min(*n_desc_p, cap_desc);
    for (u32 base = gwarp * 32; base < nd; base += nwarps * 32) {
        const u32 i = base + lane;
        RDesc d = i < nd ? descs[i] : RDesc{0, 0, 0, 0};
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
