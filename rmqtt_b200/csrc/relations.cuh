// Device-side relation expansion: what DefaultRouter::_matches does with every matched relation
// AFTER the trie walk (rmqtt/src/router.rs:182-239, rmqtt/src/types.rs:478-508), for a whole batch:
//
//   * `no_local` (router.rs:184-189): a v5 subscriber that set no_local does not receive its own PUBLISH — the relation is
//     dropped when the publisher's Id equals the subscriber's Id;
//   * shared-subscription members (router.rs:192-200) and v3 relations pass through (the random choice of ONE member per
//     group, router.rs:224-238, stays on the host: it is rand::random in the reference);
//   * v5 relations are de-duplicated PER CLIENT (types.rs:488-506): a client that matches through several filters gets ONE
//     relation, and the subscription identifiers of all its matching subscriptions accumulate.
//
// Input: the match kernels' own output (spans + relation handles, device buffers) and a 16-byte record per handle.
// Output: per topic the finished gm_sub_relation records {node id, handle, group, sub-id range} (the reference's SubRelation
// minus the strings the host resolves from the handle) and the accumulated subscription ids.  One warp per topic; the v5 relations of a topic are staged
// in shared memory and de-duplicated by an all-pairs pass (a topic rarely has more than a few dozen); a topic with more
// than REL_STAGE v5 relations is flagged (status 1) and handed to the host un-deduplicated.
#pragma once
#include <cuda_runtime.h>

#include "../../include/gpumqtt.h"
#include "kernels.cuh"

namespace gm {

constexpr u32 REL_STAGE = 256;
constexpr u32 REL_NONE = 0xFFFFFFFFu;

struct RelParams {
    const uint2* spans; const u32* ids; u32 n;
    const u32* pubs;                 // [n] id_idx of the publisher or REL_NONE; may be null
    const gm_rel* rels; u32 n_rels;
    uint2* out_spans; gm_sub_relation* out_rels; unsigned long long cap_rels;
    u32* out_sub_ids; unsigned long long cap_sub_ids;
    unsigned long long* needed;      // [2] relations, sub ids  ([2] unused)
    int* status;                     // [n]
};

__device__ __forceinline__ u32 warp_incl_scan(u32 v, u32 lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xFFFFFFFFu, v, o); if (lane >= o) v += x; }
    return v;
}

__device__ __forceinline__ gm_rel load_rel(const gm_rel* p) {      // 24-byte record, 8-byte aligned
    const uint2* q = reinterpret_cast<const uint2*>(p);
    const uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    gm_rel r;
    r.node_id = (static_cast<unsigned long long>(a.y) << 32) | a.x; r.client_key = b.x; r.id_idx = b.y; r.sub_id = c.x; r.flags = c.y;
    return r;
}

__global__ void __launch_bounds__(256)
k_relations(RelParams p) {
    __shared__ u32 s_key[8][REL_STAGE], s_h[8][REL_STAGE], s_sub[8][REL_STAGE];
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lt = lanemask_lt();
    const u32 gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    u32* sk = s_key[wid]; u32* sh = s_h[wid]; u32* ss = s_sub[wid];
    for (u32 t = gwarp; t < p.n; t += nwarps) {
        const uint2 sp = p.spans[t];
        const u32 off = sp.x, cnt = sp.y;
        const u32 pub = p.pubs ? p.pubs[t] : REL_NONE;
        // classify one handle: 0 dropped, 1 direct (v3 / shared-group member), 2 v5 relation subject to per-client de-dup
        auto classify = [&](u32 i, u32& h, gm_rel& r) -> u32 {
            if (i >= cnt) return 0u;
            h = p.ids[off + i];
            if (h >= p.n_rels) return 0u;
            r = load_rel(p.rels + h);
            if (!(r.flags & GM_REL_LIVE)) return 0u;
            if ((r.flags & GM_REL_V5) && (r.flags & GM_REL_NO_LOCAL) && pub != REL_NONE && r.id_idx == pub) return 0u;   // router.rs:184-189
            return ((r.flags & GM_REL_V5) && (r.flags >> 8) == 0u) ? 2u : 1u;
        };
        // ---- pass 1: count, stage the v5 relations ----
        u32 ndirect = 0, nv = 0;
        for (u32 i0 = 0; i0 < cnt; i0 += 32) {
            u32 h = 0; gm_rel r{};
            const u32 c = classify(i0 + lane, h, r);
            ndirect += __popc(__ballot_sync(0xFFFFFFFFu, c == 1u));
            const u32 bv = __ballot_sync(0xFFFFFFFFu, c == 2u);
            if (c == 2u) { const u32 at = nv + __popc(bv & lt); if (at < REL_STAGE) { sk[at] = r.client_key; sh[at] = h; ss[at] = r.sub_id; } }
            nv += __popc(bv);
        }
        __syncwarp();
        const bool overflow = nv > REL_STAGE;
        // ---- de-dup: entry e represents its client iff it holds the smallest handle of the client's relations (the same handle can
        //      occur twice: a topic with a literal '+' / '#' level visits a wildcard child twice — ties go to the earlier entry) ----
        u32 nreps = 0, nsubs = 0;
        if (!overflow) {
            for (u32 e0 = 0; e0 < nv; e0 += 32) {
                const u32 e = e0 + lane;
                bool rep = false; u32 nsub = 0;
                if (e < nv) {
                    const u32 key = sk[e], h = sh[e];
                    rep = true;
                    for (u32 j = 0; j < nv; ++j) if (sk[j] == key) { rep &= !(sh[j] < h || (sh[j] == h && j < e)); nsub += ss[j] != 0u; }
                }
                nreps += __popc(__ballot_sync(0xFFFFFFFFu, rep));
                nsubs += __reduce_add_sync(0xFFFFFFFFu, rep ? nsub : 0u);
            }
        } else nreps = nv;                       // handed over un-deduplicated
        const u32 total = ndirect + nreps;
        unsigned long long hb = 0, sb = 0;
        if (lane == 0) {
            if (total) hb = atomicAdd(p.needed + 0, static_cast<unsigned long long>(total));
            if (nsubs) sb = atomicAdd(p.needed + 1, static_cast<unsigned long long>(nsubs));
        }
        hb = __shfl_sync(0xFFFFFFFFu, hb, 0); sb = __shfl_sync(0xFFFFFFFFu, sb, 0);
        const bool fits = hb + total <= p.cap_rels && sb + nsubs <= p.cap_sub_ids && hb + total <= 0xFFFFFFFFull && sb + nsubs <= 0xFFFFFFFFull;
        if (lane == 0) { p.out_spans[t] = make_uint2(fits ? static_cast<u32>(hb) : 0u, total); if (overflow) p.status[t] = 1; }
        if (!fits || total == 0) { __syncwarp(); continue; }
        // ---- pass 2: write the finished gm_sub_relation records.  Direct relations first (after an overflow, the v5 ones with them) ----
        u32 w = 0;
        for (u32 i0 = 0; i0 < cnt; i0 += 32) {
            u32 h = 0; gm_rel r{};
            const u32 c = classify(i0 + lane, h, r);
            const bool take = c == 1u || (overflow && c == 2u);
            const u32 b = __ballot_sync(0xFFFFFFFFu, take);
            if (take) p.out_rels[hb + w + __popc(b & lt)] = gm_sub_relation{r.node_id, h, r.flags >> 8, 0u, 0u};
            w += __popc(b);
        }
        if (!overflow) {
            u32 ws = 0;
            for (u32 e0 = 0; e0 < nv; e0 += 32) {
                const u32 e = e0 + lane;
                bool rep = false; u32 nsub = 0, key = 0, h = 0;
                if (e < nv) {
                    key = sk[e]; h = sh[e];
                    rep = true;
                    for (u32 j = 0; j < nv; ++j) if (sk[j] == key) { rep &= !(sh[j] < h || (sh[j] == h && j < e)); nsub += ss[j] != 0u; }
                }
                const u32 br = __ballot_sync(0xFFFFFFFFu, rep);
                const u32 mys = rep ? nsub : 0u;
                const u32 inc = warp_incl_scan(mys, lane);
                if (rep) {
                    const u32 so = static_cast<u32>(sb) + ws + (inc - mys);
                    p.out_rels[hb + w + __popc(br & lt)] = gm_sub_relation{load_rel(p.rels + h).node_id, h, 0u, nsub ? so : 0u, nsub};
                    u32 k = 0;
                    for (u32 j = 0; j < nv && k < nsub; ++j) if (sk[j] == key && ss[j] != 0u) p.out_sub_ids[so + k++] = ss[j];   // types.rs:497-503
                }
                w += __popc(br); ws += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
        }
        __syncwarp();
    }
}

}  // namespace gm
