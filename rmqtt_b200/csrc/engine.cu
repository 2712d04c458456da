// libgpumqtt: C ABI (include/gpumqtt.h) over the host mirror (host_trie.cpp) and the sm_100a kernels
// (kernels.cuh).  There is no CPU fallback: without a CUDA device every entry point that would match
// returns GM_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/gpumqtt.h"
#include "host_trie.h"
#include "comm.cuh"
#include "kernels.cuh"
#include "retain_kernels.cuh"
#include "retain_tree.h"
#include "router_host.h"

using namespace gm;

namespace {

thread_local std::string g_err;

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            g_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                             \
            return (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? GM_ERR_NO_DEVICE : GM_ERR_CUDA; \
        }                                                                                           \
    } while (0)

struct DevBuf {   // owning device allocation (freed on destruction, also on the early-return error paths)
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t ensure(size_t bytes, bool keep = false, cudaStream_t s = nullptr) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = std::max(bytes, cap + cap / 2);
        ncap = (ncap + 255) & ~size_t(255);
        void* np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess) return e;
        if (keep && p && cap) { e = cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, s); if (e != cudaSuccess) return e; cudaStreamSynchronize(s); }
        if (p) cudaFree(p);
        p = np; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct Ctrl {                       // device control block, zeroed before every match
    unsigned long long cursor;
    unsigned long long stats[24];
    u32 slow_count;
    u32 tile_counter;
};

static_assert(GM_ERR_INVALID_TOPIC == -2, "k_tokenize writes the per-topic status code directly");

// fast-path geometry: one topic per thread, 512 threads per CTA, 3 CTAs per SM (64 KB of shared memory each)
constexpr int K2_FAST_L = 8;       // levels staged in shared memory; deeper topics take the deferred kernel
constexpr int K2_THREADS = 512;
constexpr int K2_CTAS_PER_SM = 3;
constexpr u32 K2_POOL_ROWS = 24;   // matched value sets per topic beyond the 8 held in shared memory, before the topic is deferred

}  // namespace

struct gm_engine {
    std::mutex mu;
    int device = 0;
    u32 flags = 0;
    int num_sms = 0;
    HostTrie trie;
    RetainTreeHost rtree{&trie};     // retained-message tree (shares the level dictionary)
    cudaStream_t stream = nullptr;   // host-buffer matches
    cudaStream_t side = nullptr;     // flush
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of the pipelined host-buffer match
    static constexpr int MAXC = 32;  // chunks per pipelined call
    cudaEvent_t ev_h2d[MAXC] = {}, ev_comp[MAXC] = {};
    unsigned long long* h_cur = nullptr;             // pinned: cursor snapshot after every chunk
    static constexpr int RING = 64;   // per-kernel timing events of the last RING match calls
    cudaEvent_t ev_flush = nullptr, ev_match = nullptr;
    cudaEvent_t ev_ring[RING][4] = {};
    u64 ring_n = 0;
    bool match_recorded = false;
    // device tables
    DevBuf d_edges, d_ranges, d_values, d_dict, d_pool, d_cfilter;
    size_t up_ranges = 0, up_values = 0, up_pool = 0, up_edges_slots = 0, up_dict_slots = 0;
    u64 up_values_epoch = 0;
    // scratch
    DevBuf d_tok, d_tok8, d_meta, d_slow, d_ctrl, d_gstack, d_gpool, d_patch_idx, d_patch_data, d_sort, d_hist;
    DevBuf d_blob, d_offs, d_spans, d_ids, d_status;
    // retained tree (device copy of the flattened arrays) + scratch of the retained lookup
    DevBuf d_rnodes, d_rkids, d_redges, d_rvals;
    DevBuf d_rfront[2], d_rdescs, d_rctl, d_rq;
    u32 r_cap_items = 1u << 22, r_cap_desc = 1u << 22;   // totals over the RQ slices of each queue
    u64 launches = 0;
    bool k2_attr_set = false;
    // multi-GPU (comm.cuh): NCCL communicator of the root-hash shards, scratch of the size exchange
    ncclComm_t comm = nullptr;
    u32 comm_rank = 0, comm_world = 1;
    DevBuf d_comm, d_part;
    unsigned long long* h_comm = nullptr;   // pinned [2 * world + 64]
    // tuning / diagnostics knobs, read from the environment once at creation
    struct Knobs { u32 site_bits = 14, sub_bits = 0; bool sorted_rows = true; int k2_ctas = 0; u32 diag_flags = 0; u32 tile_chunk = 1; } knobs;
    void read_knobs() {
        if (const char* ev = getenv("GM_BUCKET_BITS")) { int a = 14, b = 0; if (sscanf(ev, "%d,%d", &a, &b) >= 1 && a >= 10 && b >= 0 && a + b <= int(MAX_BUCKET_BITS)) { knobs.site_bits = a; knobs.sub_bits = b; } }
        if (const char* ev = getenv("GM_SORTED_ROWS")) knobs.sorted_rows = atoi(ev) != 0;
        if (const char* ev = getenv("GM_K2_CTAS")) knobs.k2_ctas = atoi(ev);
        if (const char* ev = getenv("GM_TILE_CHUNK")) { int v = atoi(ev); if (v >= 1 && v <= 1024) knobs.tile_chunk = static_cast<u32>(v); }
        if (getenv("GM_DIAG_NO_PUBLISH")) knobs.diag_flags |= MP_DIAG_NO_PUBLISH;
    }

    explicit gm_engine(u32 max_levels) : trie(max_levels) {}

    // The view the kernels get is a SNAPSHOT taken when the tables were last shipped (flush): with
    // GM_FLAG_MANUAL_FLUSH the host mirror may already have grown / re-hashed a table that the device has not seen yet.
    TrieView dev_view{};
    RetainView dev_rview{};
    size_t up_rkids = 0, up_rvals = 0, up_redges_slots = 0;
    TrieView view() const {
        TrieView v{};
        v.edges = d_edges.as<EdgeSlot>(); v.ranges = d_ranges.as<Range>();
        v.values = d_values.as<u32>(); v.dict = d_dict.as<DictSlot>(); v.pool = d_pool.as<u8>();
        v.cfilter = d_cfilter.as<u32>(); v.cfilter_mask = static_cast<u32>(trie.cfilter.size() - 1);
        v.edge_mask = static_cast<u32>(trie.edges.size() - 1);
        v.win_mask = trie.win_mask(); v.win_shift = trie.win_shift(); v.nwin_mask = trie.nwin_mask();
        v.dict_mask = static_cast<u32>(trie.dict.size() - 1);
        v.root_plus = trie.root_plus; v.root_hash_ref = trie.root_hash_ref; v.root_hash_cnt = trie.root_hash_cnt; v.root_mask = trie.root_mask;
        v.max_depth = trie.max_depth;
        return v;
    }

    // ---- flush: ship the staged mutations to HBM on the side stream -------------------------------
    // Scatter the listed (already final) host slots into the device copy of the table.
    template <class V>
    int patch_table(DevBuf& buf, const V& host, std::vector<u32>& dirty) {
        using T = std::remove_cv_t<std::remove_reference_t<decltype(host[0])>>;
        std::sort(dirty.begin(), dirty.end());
        dirty.erase(std::unique(dirty.begin(), dirty.end()), dirty.end());
        const u32 nd = static_cast<u32>(dirty.size());
        if (nd == 0) return GM_OK;
        std::vector<T> data(nd);
        for (u32 i = 0; i < nd; ++i) data[i] = host[dirty[i]];
        CUDA_TRY(d_patch_idx.ensure(nd * sizeof(u32)));
        CUDA_TRY(d_patch_data.ensure(nd * sizeof(T)));
        CUDA_TRY(cudaMemcpyAsync(d_patch_idx.p, dirty.data(), nd * sizeof(u32), cudaMemcpyHostToDevice, side));
        CUDA_TRY(cudaMemcpyAsync(d_patch_data.p, data.data(), nd * sizeof(T), cudaMemcpyHostToDevice, side));
        k_apply_patches<T><<<(nd + 255) / 256, 256, 0, side>>>(buf.as<T>(), d_patch_idx.as<u32>(), d_patch_data.as<T>(), nd);
        launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(side));   // the staging vectors go out of scope
        dirty.clear();
        return GM_OK;
    }

    // Hash tables (edges, dict): whole-table copy after a re-hash or when most of it changed, else patches.
    template <class Slot, class A>
    int upload_table(DevBuf& buf, const std::vector<Slot, A>& host, bool& full, std::vector<u32>& dirty, size_t& up_slots) {
        const size_t bytes = host.size() * sizeof(Slot);
        if (full || up_slots != host.size() || dirty.size() * 8 > host.size()) {
            CUDA_TRY(buf.ensure(bytes));
            CUDA_TRY(cudaMemcpyAsync(buf.p, host.data(), bytes, cudaMemcpyHostToDevice, side));
            up_slots = host.size();
            dirty.clear();
        } else {
            int st = patch_table(buf, host, dirty);
            if (st != GM_OK) return st;
        }
        full = false;
        return GM_OK;
    }

    // Append-only arrays (plus, ranges, values, pool): copy the new tail; patch older entries that changed.
    template <class V>
    int upload_appendable(DevBuf& buf, const V& host, size_t& up, std::vector<u32>* dirty) {
        using T = std::remove_cv_t<std::remove_reference_t<decltype(host[0])>>;
        const size_t bytes = host.size() * sizeof(T);
        if (bytes > buf.cap) {   // grow: re-ship the whole array from the mirror
            CUDA_TRY(buf.ensure(std::max(bytes * 2, size_t(4096))));
            up = 0;
        }
        const size_t before = up;
        if (host.size() > up) {
            CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(buf.p) + up * sizeof(T), host.data() + up, (host.size() - up) * sizeof(T), cudaMemcpyHostToDevice, side));
            up = host.size();
        }
        if (dirty) {
            dirty->erase(std::remove_if(dirty->begin(), dirty->end(), [&](u32 i) { return i >= before; }), dirty->end());
            int st = patch_table(buf, host, *dirty);
            if (st != GM_OK) return st;
            dirty->clear();
        }
        return GM_OK;
    }

    RetainView rview() const {
        RetainView v{};
        v.kids = d_rkids.as<RKid>(); v.edges = d_redges.as<REdge>(); v.vals = d_rvals.as<u32>();
        v.edge_mask = static_cast<u32>(rtree.redges.size() - 1);
        if (!rtree.rnodes.empty()) { v.root_first_kid = rtree.rnodes[0].first_kid; v.root_nk_flags = rtree.rnodes[0].nkids | (rtree.rnodes[0].flags << 28); }
        v.root_plain_kids = rtree.root_plain_kids; v.root_plain_val_hi = rtree.root_plain_val_hi; v.max_depth = rtree.max_depth;
        return v;
    }

    template <class T, class A>
    int upload_whole(DevBuf& buf, const std::vector<T, A>& host, size_t slack_elems = 0) {
        CUDA_TRY(buf.ensure(std::max<size_t>((host.size() + slack_elems) * sizeof(T), 256)));
        if (!host.empty()) CUDA_TRY(cudaMemcpyAsync(buf.p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice, side));
        return GM_OK;
    }

    int flush_locked() {
        if (flags & GM_FLAG_HOST_ONLY) {
            if (!trie.sync()) { g_err = "more than 2^32 live value words"; return GM_ERR_TOO_LARGE; }
            if (rtree.dirty) { rtree.prepare_flush(); rtree.shipped(); }
            return GM_OK;
        }
        if (!trie.any_dirty() && !rtree.dirty) return GM_OK;
        CUDA_TRY(cudaSetDevice(device));
        if (!trie.sync()) { g_err = "more than 2^32 live value words"; return GM_ERR_TOO_LARGE; }
        if (trie.values_epoch != up_values_epoch) { up_values = up_ranges = 0; up_values_epoch = trie.values_epoch; }   // value sets were compacted: re-ship whole
        if (match_recorded) CUDA_TRY(cudaStreamWaitEvent(side, ev_match, 0));   // never patch under a running match
        if (rtree.dirty) {   // retained tree: whole arrays after a (re)flatten, else only the entries set / remove edited in place
            rtree.prepare_flush();
            int rs;
            if (rtree.full) {
                if ((rs = upload_whole(d_rnodes, std::vector<u32>{0u})) != GM_OK) return rs;   // marker: "retained tree shipped" (records travel in rkids / redges)
                if ((rs = upload_whole(d_rkids, rtree.rkids, rtree.rkids.size() / 4 + 1024)) != GM_OK) return rs;   // room for in-place appends
                if ((rs = upload_whole(d_redges, rtree.redges)) != GM_OK) return rs;
                if ((rs = upload_whole(d_rvals, rtree.rvals)) != GM_OK) return rs;
                up_rkids = rtree.rkids.size(); up_rvals = rtree.rvals.size(); up_redges_slots = rtree.redges.size();
            } else {
                bool full_edges = false;
                if ((rs = upload_appendable(d_rkids, rtree.rkids, up_rkids, &rtree.dirty_kids)) != GM_OK) return rs;
                if ((rs = upload_table(d_redges, rtree.redges, full_edges, rtree.dirty_edges, up_redges_slots)) != GM_OK) return rs;
                if ((rs = upload_appendable(d_rvals, rtree.rvals, up_rvals, &rtree.dirty_vals)) != GM_OK) return rs;
            }
            rtree.shipped();
        }
        int st;
        if ((st = upload_table(d_edges, trie.edges, trie.full_edges, trie.dirty_edges, up_edges_slots)) != GM_OK) return st;
        if ((st = upload_table(d_dict, trie.dict, trie.full_dict, trie.dirty_dict, up_dict_slots)) != GM_OK) return st;
        if ((st = upload_appendable(d_ranges, trie.ranges, up_ranges, nullptr)) != GM_OK) return st;
        if ((st = upload_appendable(d_values, trie.values, up_values, nullptr)) != GM_OK) return st;
        if ((st = upload_appendable(d_pool, trie.pool, up_pool, nullptr)) != GM_OK) return st;
        if (trie.cfilter_dirty) {   // child filter of wide nodes: a few MB, shipped whole
            if ((st = upload_whole(d_cfilter, trie.cfilter)) != GM_OK) return st;
            trie.cfilter_dirty = false;
        }
        trie.root_dirty = false;
        dev_view = view();
        dev_rview = rview();
        CUDA_TRY(cudaEventRecord(ev_flush, side));
        CUDA_TRY(cudaStreamSynchronize(side));
        return GM_OK;
    }

    // ---- the match pipeline, all on `s`, all buffers on the device ---------------------------------
    // `desc`: descriptor mode (d_ids_ is then a uint2 array of value-set references, cap_ids counts descriptors).
    // `d_sel`: optional selection — row t matches entry d_sel[t] of the packed batch (n = number of selected rows).
    int enqueue_match(const void* d_blob_, u64 blob_bytes, const u32* d_offs_, u64 n, gm_span* d_spans_, void* d_ids_, u64 cap_ids,
                      u64* d_needed, int32_t* d_status_, cudaStream_t s, bool stats, bool keep_cursor = false, bool desc = false,
                      const u32* d_sel = nullptr) {
        if (n == 0) { if (d_needed) CUDA_TRY(cudaMemsetAsync(d_needed, 0, sizeof(u64), s)); return GM_OK; }
        if (n > 0xFFFFFFF0ull) { g_err = "batch too large"; return GM_ERR_TOO_LARGE; }
        if (cap_ids > 0xFFFFFFFFull) cap_ids = 0xFFFFFFFFull;   // spans carry 32-bit offsets
        const u32 n32 = static_cast<u32>(n);
        const u32 S = std::max<u32>(1u, dev_view.max_depth);
        CUDA_TRY(d_tok.ensure(S > TOK8 ? static_cast<size_t>(S) * n32 * sizeof(u32) : 256));
        CUDA_TRY(d_tok8.ensure(static_cast<size_t>(n32) * TOK8 * sizeof(u32)));
        CUDA_TRY(d_meta.ensure(n32 * sizeof(u32)));
        CUDA_TRY(d_slow.ensure(n32 * sizeof(u32)));
        CUDA_TRY(d_ctrl.ensure(sizeof(Ctrl)));
        // locality pass scratch: bkey[n], perm[n]; hist + cursor [NBUCKETS] each
        CUDA_TRY(d_sort.ensure(static_cast<size_t>(n32) * 11 * sizeof(u32) + 64));
        const u32 site_bits = knobs.site_bits, sub_bits = knobs.sub_bits;
        const u32 NBUCKETS = 1u << (site_bits + sub_bits);
        CUDA_TRY(d_hist.ensure(2 * static_cast<size_t>(NBUCKETS) * sizeof(u32)));
        u32* bkey = d_sort.as<u32>();
        u32* perm = bkey + n32;
        u32* meta_sorted = perm + n32;
        u32* tok8_sorted = meta_sorted + n32 + ((8 - (3 * static_cast<size_t>(n32)) % 8) % 8);   // 32-byte aligned rows
        const bool sorted_rows = knobs.sorted_rows;
        u32* hist = d_hist.as<u32>();
        u32* bcursor = hist + NBUCKETS;
        const int k3_blocks = num_sms * 4;
        const u32 stack_cap = 32u * (dev_view.max_depth + 2u) + 64u;
        CUDA_TRY(d_gstack.ensure(static_cast<size_t>(k3_blocks) * 8 * stack_cap * sizeof(u64)));
        const int k2_ctas = (knobs.k2_ctas >= 1 && knobs.k2_ctas <= K2_CTAS_PER_SM) ? knobs.k2_ctas : K2_CTAS_PER_SM;
        const int k2_grid = num_sms * k2_ctas;
        CUDA_TRY(d_gpool.ensure(static_cast<size_t>(k2_grid) * K2_THREADS * K2_POOL_ROWS * sizeof(Desc)));
        CUDA_TRY(cudaStreamWaitEvent(s, ev_flush, 0));
        if (match_recorded) CUDA_TRY(cudaStreamWaitEvent(s, ev_match, 0));   // scratch is shared: one match in flight
        // the bump cursor over out_ids survives between the chunks of one pipelined host call
        if (keep_cursor) CUDA_TRY(cudaMemsetAsync(static_cast<char*>(d_ctrl.p) + sizeof(unsigned long long), 0, sizeof(Ctrl) - sizeof(unsigned long long), s));
        else CUDA_TRY(cudaMemsetAsync(d_ctrl.p, 0, sizeof(Ctrl), s));
        CUDA_TRY(cudaMemsetAsync(hist, 0, NBUCKETS * sizeof(u32), s));
        Ctrl* ctrl = d_ctrl.as<Ctrl>();
        const TrieView tv = dev_view;

        cudaEvent_t* ev_t = ev_ring[ring_n % RING];
        CUDA_TRY(cudaEventRecord(ev_t[0], s));
        k_tokenize<<<(n32 + TOK_THREADS - 1) / TOK_THREADS, TOK_THREADS, 0, s>>>(
            static_cast<const u8*>(d_blob_), static_cast<u32>(blob_bytes), d_offs_, d_sel, n32, tv, S, d_tok8.as<u32>(), d_tok.as<u32>(), d_meta.as<u32>(), d_status_, bkey, hist, site_bits, sub_bits);
        k_bucket_scan<<<1, 1024, 0, s>>>(hist, bcursor, NBUCKETS);
        k_bucket_scatter<<<(n32 + 255) / 256, 256, 0, s>>>(bkey, bcursor, n32, perm, d_tok8.as<u32>(), d_meta.as<u32>(),
                                                            sorted_rows ? tok8_sorted : nullptr, meta_sorted);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_t[1], s));

        MatchParams mp{};
        mp.tv = tv; mp.tok8 = d_tok8.as<u32>(); mp.tok = d_tok.as<u32>(); mp.meta = d_meta.as<u32>(); mp.n = n32; mp.tok_levels = S;
        mp.spans = reinterpret_cast<uint2*>(d_spans_); mp.out_ids = static_cast<u32*>(d_ids_); mp.out_desc = static_cast<uint2*>(d_ids_); mp.cap_ids = cap_ids;
        mp.status = d_status_;
        mp.cursor = &ctrl->cursor; mp.slow_list = d_slow.as<u32>(); mp.slow_count = &ctrl->slow_count;
        mp.tile_counter = &ctrl->tile_counter; mp.stats = ctrl->stats;
        mp.perm = perm; mp.tok8_sorted = tok8_sorted; mp.meta_sorted = meta_sorted;
        mp.flags = (sorted_rows ? MP_SORTED_ROWS : 0u) | knobs.diag_flags;
        mp.tile_chunk = knobs.tile_chunk;
        constexpr size_t k2_smem = k2_smem_bytes<K2_FAST_L, K2_THREADS>();
        auto k2_ss = k_match_fast<K2_FAST_L, K2_THREADS, K2_CTAS_PER_SM, true, false>;
        auto k2_sd = k_match_fast<K2_FAST_L, K2_THREADS, K2_CTAS_PER_SM, true, true>;
        auto k2_ns = k_match_fast<K2_FAST_L, K2_THREADS, K2_CTAS_PER_SM, false, false>;
        auto k2_nd = k_match_fast<K2_FAST_L, K2_THREADS, K2_CTAS_PER_SM, false, true>;
        if (!k2_attr_set) {
            for (auto k : {k2_ss, k2_sd, k2_ns, k2_nd}) CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(k2_smem)));
            k2_attr_set = true;
        }
        (stats ? (desc ? k2_sd : k2_ss) : (desc ? k2_nd : k2_ns))<<<k2_grid, K2_THREADS, k2_smem, s>>>(mp, d_gpool.as<Desc>(), K2_POOL_ROWS);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_t[2], s));
        auto k3 = stats ? (desc ? k_match_slow<true, true> : k_match_slow<true, false>) : (desc ? k_match_slow<false, true> : k_match_slow<false, false>);
        k3<<<k3_blocks, 256, 0, s>>>(mp, d_gstack.as<u64>(), stack_cap);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_t[3], s));
        ring_n++;
        launches += 5;
        if (d_needed) CUDA_TRY(cudaMemcpyAsync(d_needed, &ctrl->cursor, sizeof(u64), cudaMemcpyDeviceToDevice, s));
        CUDA_TRY(cudaEventRecord(ev_match, s));
        match_recorded = true;
        return GM_OK;
    }

    // ---- retained lookup: tokenise filters -> frontier BFS (one step kernel per tree level) -> publish ----
    struct RCtl { unsigned long long grand; u32 err; u32 pad; u32 n_desc[RQ]; u32 counts[RQ]; };   // counts[(max_depth + 3) * RQ] follow

    int enqueue_retain(const void* d_blob_, u64 blob_bytes, const u32* d_offs_, u64 n, gm_span* d_spans_, u32* d_ids_, u64 cap_ids, int32_t* d_status_, cudaStream_t s) {
        const u32 nq = static_cast<u32>(n);
        const u32 depth = dev_rview.max_depth;
        const u32 S = depth + 2;                       // the walk reads filter levels pos and pos+1 with pos <= tree depth
        const size_t ctl_bytes = sizeof(RCtl) + static_cast<size_t>(depth + 3) * RQ * sizeof(u32);
        CUDA_TRY(d_tok.ensure(S > TOK8 ? static_cast<size_t>(S) * nq * sizeof(u32) : 256));
        CUDA_TRY(d_tok8.ensure(static_cast<size_t>(nq) * TOK8 * sizeof(u32)));
        CUDA_TRY(d_meta.ensure(nq * sizeof(u32)));
        CUDA_TRY(d_rq.ensure(static_cast<size_t>(nq) * 3 * sizeof(u32)));
        CUDA_TRY(d_rctl.ensure(ctl_bytes));
        const u32 slice_items = r_cap_items / RQ, slice_desc = r_cap_desc / RQ;   // queues are split into RQ slices
        CUDA_TRY(d_rfront[0].ensure(static_cast<size_t>(slice_items) * RQ * sizeof(RItem)));
        CUDA_TRY(d_rfront[1].ensure(static_cast<size_t>(slice_items) * RQ * sizeof(RItem)));
        CUDA_TRY(d_rdescs.ensure(static_cast<size_t>(slice_desc) * RQ * sizeof(RDesc)));
        CUDA_TRY(cudaStreamWaitEvent(s, ev_flush, 0));
        if (match_recorded) CUDA_TRY(cudaStreamWaitEvent(s, ev_match, 0));
        CUDA_TRY(cudaMemsetAsync(d_rctl.p, 0, ctl_bytes, s));
        CUDA_TRY(cudaMemsetAsync(d_rq.p, 0, static_cast<size_t>(nq) * 3 * sizeof(u32), s));
        RCtl* ctl = d_rctl.as<RCtl>();
        u32* qtotal = d_rq.as<u32>();
        u32* qbase = qtotal + nq;
        u32* qcur = qbase + nq;
        cudaEvent_t* ev_t = ev_ring[ring_n % RING];
        CUDA_TRY(cudaEventRecord(ev_t[0], s));
        k_tokenize<<<(nq + TOK_THREADS - 1) / TOK_THREADS, TOK_THREADS, 0, s>>>(static_cast<const u8*>(d_blob_), static_cast<u32>(blob_bytes), d_offs_, nullptr, nq, dev_view, S, d_tok8.as<u32>(), d_tok.as<u32>(), d_meta.as<u32>(), d_status_, nullptr, nullptr, 0u, 0u);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_t[1], s));
        RetainParams rp{};
        rp.v = dev_rview; rp.qtok8 = d_tok8.as<u32>(); rp.qtok = d_tok.as<u32>(); rp.qmeta = d_meta.as<u32>(); rp.nq = nq; rp.tok_levels = S;
        rp.descs = d_rdescs.as<RDesc>(); rp.n_desc = ctl->n_desc; rp.cap_items = slice_items; rp.cap_desc = slice_desc;
        rp.qtotal = qtotal; rp.err = &ctl->err;
        k_retain_init<<<(nq + 255) / 256, 256, 0, s>>>(rp, d_rfront[0].as<RItem>(), &ctl->counts[0]);
        const int grid = num_sms * 8;
        for (u32 lvl = 0; lvl <= depth; ++lvl)
            k_retain_step<<<grid, 256, 0, s>>>(rp, d_rfront[lvl & 1].as<RItem>(), &ctl->counts[static_cast<size_t>(lvl) * RQ], d_rfront[(lvl + 1) & 1].as<RItem>(),
                                               &ctl->counts[static_cast<size_t>(lvl + 1) * RQ]);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_t[2], s));
        k_retain_scan<<<1, 1024, 0, s>>>(qtotal, nq, qbase, reinterpret_cast<uint2*>(d_spans_), &ctl->grand);
        k_retain_expand<<<grid, 256, 0, s>>>(d_rdescs.as<RDesc>(), ctl->n_desc, slice_desc, d_rvals.as<u32>(), qbase, qcur, d_ids_, cap_ids);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_t[3], s));
        ring_n++;
        launches += 4 + depth + 1;
        CUDA_TRY(cudaEventRecord(ev_match, s));
        match_recorded = true;
        return GM_OK;
    }

    // runs the pipeline, growing the frontier / descriptor scratch until nothing overflowed; leaves *total
    int run_retain(const void* d_blob_, u64 blob_bytes, const u32* d_offs_, u64 n, gm_span* d_spans_, u32* d_ids_, u64 cap_ids, int32_t* d_status_, cudaStream_t s, u64* total) {
        for (int attempt = 0; attempt < 12; ++attempt) {
            int st = enqueue_retain(d_blob_, blob_bytes, d_offs_, n, d_spans_, d_ids_, cap_ids, d_status_, s);
            if (st != GM_OK) return st;
            RCtl h{};
            CUDA_TRY(cudaMemcpyAsync(&h, d_rctl.p, sizeof(RCtl), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            if (h.err == 0) { *total = h.grand; return GM_OK; }
            if (h.err & 1u) { if (r_cap_items > (1u << 30)) break; r_cap_items *= 4; }
            if (h.err & 2u) { if (r_cap_desc > (1u << 30)) break; r_cap_desc *= 4; }
        }
        g_err = "retained lookup: frontier does not fit the scratch limits, split the batch";
        return GM_ERR_TOO_LARGE;
    }
};

// =====================================================================================================
extern "C" {

const char* gm_version(void) { return "libgpumqtt 0.1 (sm_100a)"; }
const char* gm_last_error(gm_engine*) { return g_err.c_str(); }

int32_t gm_create(const gm_config* cfg, gm_engine** out) {
    if (!out) return GM_ERR_INVALID_ARG;
    *out = nullptr;
    gm_config c{};
    c.struct_size = sizeof(gm_config); c.device = -1;
    if (cfg) std::memcpy(&c, cfg, std::min<size_t>(cfg->struct_size ? cfg->struct_size : sizeof(gm_config), sizeof(gm_config)));
    if (c.flags & GM_FLAG_HOST_ONLY) {   // staging mirror only (tests, off-device shard preparation): nothing can match
        gm_engine* eng = new gm_engine(c.max_levels ? c.max_levels : 128u);
        eng->flags = c.flags;
        eng->device = -1;
        if (c.filters_hint) eng->trie.reserve(c.filters_hint);
        *out = eng;
        return GM_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_err = "no CUDA device: libgpumqtt has no CPU fallback"; return GM_ERR_NO_DEVICE; }
    int dev = c.device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) { g_err = "device ordinal out of range"; return GM_ERR_INVALID_ARG; }
    CUDA_TRY(cudaSetDevice(dev));
    gm_engine* eng = new gm_engine(c.max_levels ? c.max_levels : 128u);
    eng->device = dev;
    eng->flags = c.flags;
    eng->read_knobs();
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    eng->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->side, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < gm_engine::MAXC; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&eng->ev_h2d[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&eng->ev_comp[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaMallocHost(&eng->h_cur, gm_engine::MAXC * sizeof(unsigned long long)));
    CUDA_TRY(cudaEventCreateWithFlags(&eng->ev_flush, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&eng->ev_match, cudaEventDisableTiming));
    for (auto& q : eng->ev_ring) for (auto& ev : q) CUDA_TRY(cudaEventCreate(&ev));
    CUDA_TRY(cudaEventRecord(eng->ev_flush, eng->side));
    if (c.filters_hint) eng->trie.reserve(c.filters_hint);
    *out = eng;
    return GM_OK;
}

void gm_destroy(gm_engine* e) {
    if (!e) return;
    if (e->flags & GM_FLAG_HOST_ONLY) { delete e; return; }
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&e->d_tok8, &e->d_cfilter, &e->d_edges, &e->d_ranges, &e->d_values, &e->d_dict, &e->d_pool, &e->d_tok, &e->d_meta,
                      &e->d_slow, &e->d_ctrl, &e->d_gstack, &e->d_gpool, &e->d_sort, &e->d_hist, &e->d_patch_idx, &e->d_patch_data, &e->d_blob, &e->d_offs,
                      &e->d_spans, &e->d_ids, &e->d_status, &e->d_rnodes, &e->d_rkids, &e->d_redges, &e->d_rvals, &e->d_rfront[0],
                      &e->d_rfront[1], &e->d_rdescs, &e->d_rctl, &e->d_rq})
        b->release();
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
    if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
    for (int i = 0; i < gm_engine::MAXC; ++i) { if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]); if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]); }
    if (e->h_cur) cudaFreeHost(e->h_cur);
    if (e->comm) { NcclApi::get().CommDestroy(e->comm); e->comm = nullptr; }
    if (e->h_comm) cudaFreeHost(e->h_comm);
    e->d_comm.release(); e->d_part.release();
    if (e->ev_flush) cudaEventDestroy(e->ev_flush);
    if (e->ev_match) cudaEventDestroy(e->ev_match);
    for (auto& q : e->ev_ring) for (auto& ev : q) if (ev) cudaEventDestroy(ev);
    delete e;
}

static int32_t map_parse(int st, const char* what) {
    if (st == PARSE_OK) return GM_OK;
    if (st == PARSE_TOO_DEEP) { g_err = std::string(what) + ": filter deeper than max_levels"; return GM_ERR_TOO_DEEP; }
    g_err = std::string(what) + ": invalid topic filter";
    return GM_ERR_INVALID_TOPIC;
}

int32_t gm_sub_add(gm_engine* e, const char* filter, uint32_t len, uint32_t value, int32_t* changed) {
    if (!e || (!filter && len)) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    bool ch = false;
    int st = e->trie.insert(filter, len, value, &ch);
    if (changed) *changed = ch ? 1 : 0;
    return map_parse(st, "gm_sub_add");
}

int32_t gm_sub_remove(gm_engine* e, const char* filter, uint32_t len, uint32_t value, int32_t* changed) {
    if (!e || (!filter && len)) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    bool ch = false;
    int st = e->trie.remove(filter, len, value, &ch);
    if (changed) *changed = ch ? 1 : 0;
    return map_parse(st, "gm_sub_remove");
}

int32_t gm_bulk_load(gm_engine* e, const char* blob, const uint32_t* offsets, const uint32_t* values, uint64_t n, uint64_t* n_changed) {
    if (!e || (n && (!blob || !offsets || !values))) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    e->trie.reserve(n);
    uint64_t ok = 0;
    if (getenv("GM_BULK_ONE_BY_ONE")) {       // A/B switch for the host-side bulk-load measurement
        for (uint64_t i = 0; i < n; ++i) { bool ch = false; e->trie.insert(blob + offsets[i], offsets[i + 1] - offsets[i], values[i], &ch); ok += ch; }
    } else ok = e->trie.insert_batch(blob, offsets, values, n);
    if (n_changed) *n_changed = ok;
    return GM_OK;
}

int32_t gm_compact(gm_engine* e) {
    if (!e) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    // the retained tree interns its levels in the same dictionary: carry its tokens over and re-label its nodes
    const std::vector<u32> keep = e->rtree.used_tokens();
    std::vector<u32> remap;
    e->trie.compact(&keep, &remap);
    e->rtree.remap_tokens(remap);
    e->up_ranges = e->up_values = e->up_pool = 0;
    e->up_values_epoch = e->trie.values_epoch;
    e->up_edges_slots = e->up_dict_slots = 0;
    return (e->flags & GM_FLAG_HOST_ONLY) ? GM_OK : e->flush_locked();
}

int32_t gm_flush(gm_engine* e) {
    if (!e) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    return e->flush_locked();
}

// shared implementation of the device-buffer entry points
static int32_t match_device_impl(gm_engine* e, const void* d_blob, uint64_t blob_bytes, const uint32_t* d_offs, uint64_t n_entries, const uint32_t* d_sel,
                                 uint64_t n, gm_span* d_spans, void* d_out, uint64_t cap, uint64_t* d_needed, int32_t* d_status, void* stream,
                                 bool desc, gm_work* work) {
    if (!e || (n && (!d_offs || !d_spans || !d_status))) return GM_ERR_INVALID_ARG;
    if (blob_bytes > 0xFFFFFFFFull) { g_err = "topic blob >= 4 GiB"; return GM_ERR_TOO_LARGE; }
    std::lock_guard<std::mutex> g(e->mu);
    if (e->flags & GM_FLAG_HOST_ONLY) { g_err = "host-only engine cannot match: there is no CPU fallback"; return GM_ERR_NO_DEVICE; }
    CUDA_TRY(cudaSetDevice(e->device));
    if (!(e->flags & GM_FLAG_MANUAL_FLUSH) || !e->d_edges.p) { int st = e->flush_locked(); if (st != GM_OK) return st; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int st = e->enqueue_match(d_blob, blob_bytes, d_offs, n, d_spans, d_out, cap, d_needed, d_status, s, work != nullptr, false, desc, d_sel);
    if (st != GM_OK || !work) return st;
    std::memset(work, 0, sizeof(*work));
    if (n == 0) return GM_OK;
    Ctrl h{};
    std::vector<u32> meta(n), offs(n_entries + 1), sel(d_sel ? n : 0);
    CUDA_TRY(cudaMemcpyAsync(&h, e->d_ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(meta.data(), e->d_meta.p, n * sizeof(u32), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(offs.data(), d_offs, (n_entries + 1) * sizeof(u32), cudaMemcpyDeviceToHost, s));
    if (d_sel) CUDA_TRY(cudaMemcpyAsync(sel.data(), d_sel, n * sizeof(u32), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    work->visited = h.stats[0]; work->probed = h.stats[1]; work->filters = h.stats[2]; work->ids = h.stats[3];
    work->deferred = h.slow_count;
    for (int k = 0; k < 8; ++k) { work->probes_by_depth[k] = h.stats[4 + k]; work->misses_by_depth[k] = h.stats[12 + k]; }
    work->slot_loads = h.stats[20];
    for (uint64_t i = 0; i < n; ++i)
        if (!(meta[i] & META_INVALID)) { const u64 j = d_sel ? sel[i] : i; work->levels += meta[i] & META_NLEV_MASK; work->bytes += offs[j + 1] - offs[j]; }
    return GM_OK;
}

int32_t gm_match_batch_device(gm_engine* e, const void* d_blob, uint64_t blob_bytes, const uint32_t* d_offs, uint64_t n,
                              gm_span* d_spans, uint32_t* d_ids, uint64_t cap_ids, uint64_t* d_needed, int32_t* d_status, void* stream) {
    return match_device_impl(e, d_blob, blob_bytes, d_offs, n, nullptr, n, d_spans, d_ids, cap_ids, d_needed, d_status, stream, false, nullptr);
}

int32_t gm_match_batch_device_stats(gm_engine* e, const void* d_blob, uint64_t blob_bytes, const uint32_t* d_offs, uint64_t n,
                                    gm_span* d_spans, uint32_t* d_ids, uint64_t cap_ids, uint64_t* d_needed, int32_t* d_status,
                                    void* stream, gm_work* work) {
    if (!work) return GM_ERR_INVALID_ARG;
    return match_device_impl(e, d_blob, blob_bytes, d_offs, n, nullptr, n, d_spans, d_ids, cap_ids, d_needed, d_status, stream, false, work);
}

int32_t gm_match_batch_device_ex(gm_engine* e, const gm_match_args* a) {
    if (!a || a->struct_size < sizeof(gm_match_args)) return GM_ERR_INVALID_ARG;
    if (a->d_sel && a->n > a->n_entries) return GM_ERR_INVALID_ARG;
    return match_device_impl(e, a->d_blob, a->blob_bytes, a->d_offsets, a->d_sel ? a->n_entries : a->n, a->d_sel, a->n, a->d_spans, a->d_out, a->cap,
                             a->d_needed, a->d_status, a->stream, (a->flags & GM_MATCH_DESCRIPTORS) != 0, a->work);
}

// shared implementation of the host-buffer entry points; `elem` = bytes per output element (4: ids, 8: descriptors)
static int32_t match_host_impl(gm_engine* e, const char* blob, const uint32_t* offsets, uint64_t n, gm_span* out_spans, void* out, size_t elem,
                               uint64_t cap_ids, uint64_t* needed, int32_t* status) {
    if (!e || (n && (!offsets || !out_spans || !status)) || (cap_ids && !out)) return GM_ERR_INVALID_ARG;
    if (needed) *needed = 0;
    if (n == 0) return GM_OK;
    if (n > 0xFFFFFFF0ull) { g_err = "batch too large"; return GM_ERR_TOO_LARGE; }
    const bool desc = elem == 8;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->flags & GM_FLAG_HOST_ONLY) { g_err = "host-only engine cannot match: there is no CPU fallback"; return GM_ERR_NO_DEVICE; }
    CUDA_TRY(cudaSetDevice(e->device));
    if (!(e->flags & GM_FLAG_MANUAL_FLUSH) || !e->d_edges.p) { int st = e->flush_locked(); if (st != GM_OK) return st; }
    const u64 blob_bytes = offsets[n];
    if (cap_ids > 0xFFFFFFFFull) cap_ids = 0xFFFFFFFFull;
    CUDA_TRY(e->d_blob.ensure(blob_bytes + 16));
    CUDA_TRY(e->d_offs.ensure((n + 1) * sizeof(u32)));
    CUDA_TRY(e->d_spans.ensure(n * sizeof(gm_span)));
    CUDA_TRY(e->d_status.ensure(n * sizeof(int32_t)));
    CUDA_TRY(e->d_ids.ensure(std::max<u64>(cap_ids, 1) * elem));
    // Pipelined in chunks over three streams: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernels of
    // chunk c.  All chunks share one bump cursor, so the output of chunk c is the contiguous range
    // [cursor after c-1, cursor after c) and can be copied out as soon as that chunk's kernels finished.
    const u64 chunk = std::max<u64>(131072, (n + gm_engine::MAXC - 1) / gm_engine::MAXC);
    const int nchunks = static_cast<int>((n + chunk - 1) / chunk);
    cudaStream_t sc = e->stream;
    if (e->match_recorded) CUDA_TRY(cudaStreamWaitEvent(e->s_h2d, e->ev_match, 0));   // previous call still reads d_blob / scratch
    for (int c = 0; c < nchunks; ++c) {
        const u64 c0 = c * chunk, c1 = std::min<u64>(n, c0 + chunk);
        const u64 b0 = offsets[c0], b1 = offsets[c1];
        if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(e->d_blob.p) + b0, blob + b0, b1 - b0, cudaMemcpyHostToDevice, e->s_h2d));
        CUDA_TRY(cudaMemcpyAsync(e->d_offs.as<u32>() + c0, offsets + c0, (c1 - c0 + 1) * sizeof(u32), cudaMemcpyHostToDevice, e->s_h2d));
        CUDA_TRY(cudaEventRecord(e->ev_h2d[c], e->s_h2d));
    }
    for (int c = 0; c < nchunks; ++c) {
        const u64 c0 = c * chunk, c1 = std::min<u64>(n, c0 + chunk);
        CUDA_TRY(cudaStreamWaitEvent(sc, e->ev_h2d[c], 0));
        int st = e->enqueue_match(e->d_blob.p, offsets[c1], e->d_offs.as<u32>() + c0, c1 - c0, e->d_spans.as<gm_span>() + c0, e->d_ids.p, cap_ids, nullptr,
                                  e->d_status.as<int32_t>() + c0, sc, false, c != 0, desc);
        if (st != GM_OK) return st;
        CUDA_TRY(cudaMemcpyAsync(&e->h_cur[c], &e->d_ctrl.as<Ctrl>()->cursor, sizeof(u64), cudaMemcpyDeviceToHost, sc));
        CUDA_TRY(cudaEventRecord(e->ev_comp[c], sc));
    }
    u64 done = 0;
    for (int c = 0; c < nchunks; ++c) {
        const u64 c0 = c * chunk, c1 = std::min<u64>(n, c0 + chunk);
        CUDA_TRY(cudaEventSynchronize(e->ev_comp[c]));
        const u64 cur = e->h_cur[c];
        CUDA_TRY(cudaMemcpyAsync(out_spans + c0, e->d_spans.as<gm_span>() + c0, (c1 - c0) * sizeof(gm_span), cudaMemcpyDeviceToHost, e->s_d2h));
        CUDA_TRY(cudaMemcpyAsync(status + c0, e->d_status.as<int32_t>() + c0, (c1 - c0) * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_d2h));
        const u64 hi = std::min<u64>(cur, cap_ids);
        if (hi > done) {
            CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(out) + done * elem, static_cast<char*>(e->d_ids.p) + done * elem, (hi - done) * elem, cudaMemcpyDeviceToHost, e->s_d2h));
            done = hi;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(e->s_d2h));
    const u64 total = e->h_cur[nchunks - 1];
    if (needed) *needed = total;
    if (total > 0xFFFFFFFFull) { g_err = "batch produces >= 2^32 output elements: split it"; return GM_ERR_TOO_LARGE; }
    if (total > cap_ids) { g_err = "output buffer too small"; return GM_ERR_CAPACITY; }
    return GM_OK;
}

int32_t gm_match_batch(gm_engine* e, const char* blob, const uint32_t* offsets, uint64_t n, gm_span* out_spans, uint32_t* out_ids,
                       uint64_t cap_ids, uint64_t* needed, int32_t* status) {
    return match_host_impl(e, blob, offsets, n, out_spans, out_ids, sizeof(uint32_t), cap_ids, needed, status);
}

int32_t gm_match_batch_desc(gm_engine* e, const char* blob, const uint32_t* offsets, uint64_t n, gm_span* out_spans, gm_desc* out_descs,
                            uint64_t cap_descs, uint64_t* needed, int32_t* status) {
    static_assert(sizeof(gm_desc) == 8, "descriptor = one 64-bit value-set reference");
    return match_host_impl(e, blob, offsets, n, out_spans, out_descs, sizeof(gm_desc), cap_descs, needed, status);
}

int32_t gm_values_view(gm_engine* e, gm_values* out) {
    if (!e || !out) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    out->values = e->trie.values.data(); out->n_values = e->trie.values.size();
    out->ranges = reinterpret_cast<const gm_span*>(e->trie.ranges.data()); out->n_ranges = e->trie.ranges.size();
    out->epoch = e->trie.values_epoch;
    return GM_OK;
}

int32_t gm_desc_expand(gm_engine* e, const gm_desc* descs, uint64_t n, uint32_t* out_ids, uint64_t cap_ids, uint64_t* needed) {
    if (!e || (n && !descs) || (cap_ids && !out_ids)) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    const u32* vals = e->trie.values.data();
    const u64 nvals = e->trie.values.size();
    u64 w = 0;
    for (uint64_t i = 0; i < n; ++i) {
        u32 ref = descs[i].ref, cnt = descs[i].cnt;
        if (cnt == 1) { if (w < cap_ids) out_ids[w] = ref; ++w; continue; }
        u64 off = ref;
        if (cnt == CNT_BIG) { if (ref >= e->trie.ranges.size()) { g_err = "gm_desc_expand: stale descriptor"; return GM_ERR_INVALID_ARG; } off = e->trie.ranges[ref].off; cnt = e->trie.ranges[ref].cnt; }
        if (off + cnt > nvals) { g_err = "gm_desc_expand: stale descriptor (values were compacted since the match)"; return GM_ERR_INVALID_ARG; }
        if (w + cnt <= cap_ids) std::memcpy(out_ids + w, vals + off, cnt * sizeof(u32));
        w += cnt;
    }
    if (needed) *needed = w;
    if (w > cap_ids) { g_err = "out_ids too small"; return GM_ERR_CAPACITY; }
    return GM_OK;
}

// ---- retained-message tree ----------------------------------------------------------------------------
int32_t gm_retain_set(gm_engine* e, const char* topic, uint32_t len, uint32_t value, int32_t* had_old, uint32_t* old_value) {
    if (!e || (!topic && len)) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    bool had = false; u32 old = 0;
    int st = e->rtree.set(topic, len, value, &had, &old);
    if (had_old) *had_old = had ? 1 : 0;
    if (old_value && had) *old_value = old;
    return map_parse(st, "gm_retain_set");
}

int32_t gm_retain_remove(gm_engine* e, const char* topic, uint32_t len, int32_t* had_old, uint32_t* old_value) {
    if (!e || (!topic && len)) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    bool had = false; u32 old = 0;
    int st = e->rtree.remove(topic, len, &had, &old);
    if (had_old) *had_old = had ? 1 : 0;
    if (old_value && had) *old_value = old;
    return map_parse(st, "gm_retain_remove");
}

int32_t gm_retain_bulk_load(gm_engine* e, const char* blob, const uint32_t* offsets, const uint32_t* values, uint64_t n, uint64_t* n_set) {
    if (!e || (n && (!blob || !offsets || !values))) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    uint64_t ok = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (e->rtree.set(blob + offsets[i], offsets[i + 1] - offsets[i], values[i], nullptr, nullptr) == PARSE_OK) ok++;
    if (n_set) *n_set = ok;
    return GM_OK;
}

int32_t gm_retain_match_batch_device(gm_engine* e, const void* d_blob, uint64_t blob_bytes, const uint32_t* d_offs, uint64_t n,
                                     gm_span* d_spans, uint32_t* d_ids, uint64_t cap_ids, uint64_t* needed, int32_t* d_status, void* stream) {
    if (!e || (n && (!d_offs || !d_spans || !d_status))) return GM_ERR_INVALID_ARG;
    if (needed) *needed = 0;
    if (n == 0) return GM_OK;
    if (blob_bytes > 0xFFFFFFFFull || n > 0xFFFFFFF0ull) { g_err = "filter batch too large"; return GM_ERR_TOO_LARGE; }
    std::lock_guard<std::mutex> g(e->mu);
    if (e->flags & GM_FLAG_HOST_ONLY) { g_err = "host-only engine cannot match: there is no CPU fallback"; return GM_ERR_NO_DEVICE; }
    CUDA_TRY(cudaSetDevice(e->device));
    if (!(e->flags & GM_FLAG_MANUAL_FLUSH) || !e->d_rnodes.p || !e->d_edges.p) { int st = e->flush_locked(); if (st != GM_OK) return st; }
    u64 total = 0;
    int st = e->run_retain(d_blob, blob_bytes, d_offs, n, d_spans, d_ids, std::min<u64>(cap_ids, 0xFFFFFFFFull), d_status, static_cast<cudaStream_t>(stream), &total);
    if (st != GM_OK) return st;
    if (needed) *needed = total;
    if (total > 0xFFFFFFFFull) { g_err = "batch produces >= 2^32 ids: split it"; return GM_ERR_TOO_LARGE; }
    if (total > cap_ids) { g_err = "out_ids too small"; return GM_ERR_CAPACITY; }
    return GM_OK;
}

int32_t gm_retain_match_batch(gm_engine* e, const char* blob, const uint32_t* offsets, uint64_t n, gm_span* out_spans, uint32_t* out_ids,
                              uint64_t cap_ids, uint64_t* needed, int32_t* status) {
    if (!e || (n && (!offsets || !out_spans || !status)) || (cap_ids && !out_ids)) return GM_ERR_INVALID_ARG;
    if (needed) *needed = 0;
    if (n == 0) return GM_OK;
    if (n > 0xFFFFFFF0ull) { g_err = "filter batch too large"; return GM_ERR_TOO_LARGE; }
    std::lock_guard<std::mutex> g(e->mu);
    if (e->flags & GM_FLAG_HOST_ONLY) { g_err = "host-only engine cannot match: there is no CPU fallback"; return GM_ERR_NO_DEVICE; }
    CUDA_TRY(cudaSetDevice(e->device));
    if (!(e->flags & GM_FLAG_MANUAL_FLUSH) || !e->d_rnodes.p || !e->d_edges.p) { int st = e->flush_locked(); if (st != GM_OK) return st; }
    const u64 blob_bytes = offsets[n];
    cudaStream_t s = e->stream;
    CUDA_TRY(e->d_blob.ensure(blob_bytes + 16));
    CUDA_TRY(e->d_offs.ensure((n + 1) * sizeof(u32)));
    CUDA_TRY(e->d_spans.ensure(n * sizeof(gm_span)));
    CUDA_TRY(e->d_status.ensure(n * sizeof(int32_t)));
    CUDA_TRY(e->d_ids.ensure(std::max<u64>(cap_ids, 1) * sizeof(u32)));
    if (blob_bytes) CUDA_TRY(cudaMemcpyAsync(e->d_blob.p, blob, blob_bytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->d_offs.p, offsets, (n + 1) * sizeof(u32), cudaMemcpyHostToDevice, s));
    u64 total = 0;
    int st = e->run_retain(e->d_blob.p, blob_bytes, e->d_offs.as<u32>(), n, e->d_spans.as<gm_span>(), e->d_ids.as<u32>(), std::min<u64>(cap_ids, 0xFFFFFFFFull),
                           e->d_status.as<int32_t>(), s, &total);
    if (st != GM_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(status, e->d_status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_spans, e->d_spans.p, n * sizeof(gm_span), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (needed) *needed = total;
    if (total > 0xFFFFFFFFull) { g_err = "batch produces >= 2^32 ids: split it"; return GM_ERR_TOO_LARGE; }
    if (total > cap_ids) { g_err = "out_ids too small"; return GM_ERR_CAPACITY; }
    if (total) {
        CUDA_TRY(cudaMemcpyAsync(out_ids, e->d_ids.p, total * sizeof(u32), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    return GM_OK;
}

// ---- GpuRouter (router_host.cpp) ------------------------------------------------------------------------
struct gm_router { GpuRouter impl; explicit gm_router(gm_engine* e) : impl(e) {} };

static GpuRouter::Id to_id(const gm_id* id) {
    GpuRouter::Id r;
    r.node_id = id->node_id; r.tag = id->tag;
    if (id->client_id) r.client_id.assign(id->client_id, id->client_len);
    return r;
}

int32_t gmr_create(gm_engine* e, gm_router** out) {
    if (!e || !out) return GM_ERR_INVALID_ARG;
    *out = new gm_router(e);
    return GM_OK;
}
void gmr_destroy(gm_router* r) { delete r; }

int32_t gmr_add(gm_router* r, const char* filter, uint32_t len, const gm_id* id, const gm_sub_opts* opts) {
    if (!r || !id || (!filter && len)) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(r->impl.mu);
    GpuRouter::Opts o;
    if (opts) { o.qos = opts->qos; o.is_v5 = opts->is_v5; o.no_local = opts->no_local; o.sub_id = opts->sub_id; if (opts->shared_group) o.group.assign(opts->shared_group, opts->shared_group_len); }
    return r->impl.add(filter, len, to_id(id), o);
}
int32_t gmr_remove(gm_router* r, const char* filter, uint32_t len, const gm_id* id, int32_t* removed) {
    if (!r || !id || (!filter && len)) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(r->impl.mu);
    bool rm = false;
    int32_t rc = r->impl.remove(filter, len, to_id(id), &rm);
    if (removed) *removed = rm ? 1 : 0;
    return rc;
}
int64_t gmr_topics(gm_router* r) { return r ? r->impl.topics() : 0; }
int64_t gmr_routes(gm_router* r) { return r ? r->impl.routes() : 0; }

int32_t gmr_matches_batch(gm_router* r, const gm_id* publishers, const char* blob, const uint32_t* offs, uint64_t n, gm_span* out_spans,
                          gm_sub_relation* out_rels, uint64_t cap_rels, uint32_t* out_sub_ids, uint64_t cap_sub_ids, uint64_t* needed_rels,
                          uint64_t* needed_sub_ids, int32_t* status) {
    if (!r || (n && (!offs || !out_spans || !status))) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(r->impl.mu);
    std::vector<gm_span> spans; std::vector<gm_sub_relation> rels; std::vector<uint32_t> ids; std::vector<int32_t> st;
    int32_t rc = r->impl.matches_batch(publishers, blob, offs, n, spans, rels, ids, st);
    if (rc != GM_OK) return rc;
    if (needed_rels) *needed_rels = rels.size();
    if (needed_sub_ids) *needed_sub_ids = ids.size();
    std::copy(st.begin(), st.end(), status);
    if (rels.size() > cap_rels || ids.size() > cap_sub_ids) { g_err = "gmr_matches_batch: output too small"; return GM_ERR_CAPACITY; }
    std::copy(spans.begin(), spans.end(), out_spans);
    std::copy(rels.begin(), rels.end(), out_rels);
    std::copy(ids.begin(), ids.end(), out_sub_ids);
    return GM_OK;
}

int32_t gmr_relation(gm_router* r, uint32_t handle, const char** filter, uint32_t* filter_len, const char** client, uint32_t* client_len) {
    if (!r || !filter || !filter_len || !client || !client_len) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(r->impl.mu);
    const std::string *f, *c;
    if (!r->impl.relation(handle, &f, &c)) return GM_ERR_INVALID_ARG;
    *filter = f->data(); *filter_len = static_cast<uint32_t>(f->size()); *client = c->data(); *client_len = static_cast<uint32_t>(c->size());
    return GM_OK;
}

// ---- multi-GPU: NCCL communicator, device partition of a mixed batch, all-gatherv of match lists (comm.cuh) ----------
#define NCCL_TRY(expr)                                                                          \
    do {                                                                                        \
        ncclResult_t _r = (expr);                                                               \
        if (_r != ncclSuccess) { g_err = std::string(#expr) + ": " + nc.GetErrorString(_r); return GM_ERR_COMM; } \
    } while (0)

int32_t gm_comm_unique_id(uint8_t* out_id) {
    if (!out_id) return GM_ERR_INVALID_ARG;
    NcclApi& nc = NcclApi::get();
    if (!nc.ok()) { g_err = nc.error; return GM_ERR_COMM; }
    static_assert(GM_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "gm_comm_unique_id hands out an ncclUniqueId");
    ncclUniqueId id;
    NCCL_TRY(nc.GetUniqueId(&id));
    std::memcpy(out_id, id.internal, GM_COMM_ID_BYTES);
    return GM_OK;
}

int32_t gm_comm_init(gm_engine* e, const uint8_t* id128, uint32_t rank, uint32_t world) {
    if (!e || !id128 || world == 0 || rank >= world) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->flags & GM_FLAG_HOST_ONLY) { g_err = "host-only engine has no device to communicate from"; return GM_ERR_NO_DEVICE; }
    NcclApi& nc = NcclApi::get();
    if (!nc.ok()) { g_err = nc.error; return GM_ERR_COMM; }
    CUDA_TRY(cudaSetDevice(e->device));
    if (e->comm) { nc.CommDestroy(e->comm); e->comm = nullptr; }
    ncclUniqueId id;
    std::memcpy(id.internal, id128, GM_COMM_ID_BYTES);
    NCCL_TRY(nc.CommInitRank(&e->comm, static_cast<int>(world), id, static_cast<int>(rank)));
    e->comm_rank = rank; e->comm_world = world;
    CUDA_TRY(e->d_comm.ensure((2 + 2 * static_cast<size_t>(world)) * sizeof(unsigned long long) + (static_cast<size_t>(world) + 2) * sizeof(u32)));
    if (e->h_comm) { cudaFreeHost(e->h_comm); e->h_comm = nullptr; }
    CUDA_TRY(cudaMallocHost(&e->h_comm, (2 * static_cast<size_t>(world) + 64) * sizeof(unsigned long long)));
    return GM_OK;
}

int32_t gm_comm_destroy(gm_engine* e) {
    if (!e) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->comm) { cudaSetDevice(e->device); NcclApi::get().CommDestroy(e->comm); e->comm = nullptr; }
    e->comm_world = 1; e->comm_rank = 0;
    return GM_OK;
}

int32_t gm_partition_batch_device(gm_engine* e, const void* d_blob, uint64_t blob_bytes, const uint32_t* d_offsets, uint64_t n, uint32_t n_shards,
                                  uint32_t rank, uint32_t* d_sel, uint32_t* d_shard, uint64_t* n_local, uint64_t* shard_counts, void* stream) {
    if (!e || !n_local || n_shards == 0 || n_shards > 4096 || rank >= n_shards || (n && (!d_blob || !d_offsets || !d_sel))) return GM_ERR_INVALID_ARG;
    if (blob_bytes > 0xFFFFFFFFull || n > 0xFFFFFFF0ull) { g_err = "batch too large"; return GM_ERR_TOO_LARGE; }
    *n_local = 0;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->flags & GM_FLAG_HOST_ONLY) { g_err = "host-only engine has no device"; return GM_ERR_NO_DEVICE; }
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DevBuf& cb = e->d_part;
    CUDA_TRY(cb.ensure((static_cast<size_t>(n_shards) + 1) * sizeof(u32)));
    CUDA_TRY(cudaMemsetAsync(cb.p, 0, (static_cast<size_t>(n_shards) + 1) * sizeof(u32), s));
    if (n) {
        k_partition<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(static_cast<const u8*>(d_blob), static_cast<u32>(blob_bytes), d_offsets, static_cast<u32>(n),
                                                                            n_shards, rank, d_sel, d_shard, cb.as<u32>());
        e->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    std::vector<u32> h(n_shards + 1);
    CUDA_TRY(cudaMemcpyAsync(h.data(), cb.p, h.size() * sizeof(u32), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *n_local = h[n_shards];
    if (shard_counts) for (u32 r = 0; r < n_shards; ++r) shard_counts[r] = h[r];
    return GM_OK;
}

int32_t gm_allgatherv_device(gm_engine* e, const uint32_t* d_index, const gm_span* d_spans, uint64_t k, const uint32_t* d_ids, const uint64_t* d_m,
                             uint32_t* d_all_index, gm_span* d_all_spans, uint64_t cap_topics, uint32_t* d_all_ids, uint64_t cap_ids, uint64_t* sizes,
                             void* stream) {
    if (!e || !d_m || !sizes || (k && (!d_index || !d_spans)) || !d_all_index || !d_all_spans || !d_all_ids) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    if (!e->comm) { g_err = "gm_allgatherv_device: call gm_comm_init first"; return GM_ERR_INVALID_ARG; }
    NcclApi& nc = NcclApi::get();
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const u32 W = e->comm_world, R = e->comm_rank;
    unsigned long long* d_send = e->d_comm.as<unsigned long long>();
    unsigned long long* d_all = d_send + 2;
    // 1. sizes: (k, m) of every rank.  m lives on the device (the match kernels' cursor): no host hop before the exchange.
    k_comm_sizes<<<1, 1, 0, s>>>(d_send, k, reinterpret_cast<const unsigned long long*>(d_m));
    NCCL_TRY(nc.AllGather(d_send, d_all, 2, ncclUint64, e->comm, s));
    CUDA_TRY(cudaMemcpyAsync(e->h_comm, d_all, 2 * static_cast<size_t>(W) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));      // NCCL needs the counts on the host: the one host synchronisation of the collective
    u64 K = 0, M = 0;
    for (u32 r = 0; r < W; ++r) { sizes[2 * r] = e->h_comm[2 * r]; sizes[2 * r + 1] = e->h_comm[2 * r + 1]; K += sizes[2 * r]; M += sizes[2 * r + 1]; }
    if (M > 0xFFFFFFFFull) { g_err = "gathered match lists exceed 2^32 ids: split the batch"; return GM_ERR_TOO_LARGE; }
    if (K > cap_topics || M > cap_ids) { g_err = "gm_allgatherv_device: output too small (sizes[] holds what every rank contributes)"; return GM_ERR_CAPACITY; }
    // 2. one grouped launch: every rank broadcasts its three arrays straight out of the buffers the match kernels wrote
    NCCL_TRY(nc.GroupStart());
    u64 ko = 0, mo = 0;
    for (u32 r = 0; r < W; ++r) {
        const u64 kr = sizes[2 * r], mr = sizes[2 * r + 1];
        if (kr) {
            NCCL_TRY(nc.Broadcast(r == R ? static_cast<const void*>(d_index) : d_all_index + ko, d_all_index + ko, kr, ncclUint32, static_cast<int>(r), e->comm, s));
            NCCL_TRY(nc.Broadcast(r == R ? static_cast<const void*>(d_spans) : d_all_spans + ko, d_all_spans + ko, kr, ncclUint64, static_cast<int>(r), e->comm, s));
        }
        if (mr) NCCL_TRY(nc.Broadcast(r == R ? static_cast<const void*>(d_ids) : d_all_ids + mo, d_all_ids + mo, mr, ncclUint32, static_cast<int>(r), e->comm, s));
        ko += kr; mo += mr;
    }
    NCCL_TRY(nc.GroupEnd());
    // 3. spans of rank r index rank r's ids: re-base them onto the gathered id array
    if (K) { k_rebase_spans<<<static_cast<unsigned>((K + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint2*>(d_all_spans), d_all, W, static_cast<u32>(K)); e->launches++; }
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return GM_OK;
}

int32_t gm_tokenize_batch(gm_engine* e, const char* blob, const uint32_t* offsets, uint64_t n, uint32_t max_tok, uint32_t* out_tokens, uint32_t* out_meta) {
    if (!e || !max_tok || (n && (!offsets || !out_tokens || !out_meta))) return GM_ERR_INVALID_ARG;
    if (n == 0) return GM_OK;
    std::lock_guard<std::mutex> g(e->mu);
    if (e->flags & GM_FLAG_HOST_ONLY) { g_err = "host-only engine cannot match: there is no CPU fallback"; return GM_ERR_NO_DEVICE; }
    CUDA_TRY(cudaSetDevice(e->device));
    { int st = e->flush_locked(); if (st != GM_OK) return st; }
    cudaStream_t s = e->stream;
    const u64 blob_bytes = offsets[n];
    DevBuf tok, tok8, meta, stat;
    CUDA_TRY(e->d_blob.ensure(blob_bytes + 16));
    CUDA_TRY(e->d_offs.ensure((n + 1) * sizeof(u32)));
    CUDA_TRY(tok.ensure(static_cast<size_t>(max_tok) * n * sizeof(u32)));
    CUDA_TRY(tok8.ensure(static_cast<size_t>(n) * TOK8 * sizeof(u32)));
    CUDA_TRY(meta.ensure(n * sizeof(u32)));
    CUDA_TRY(stat.ensure(n * sizeof(int)));
    if (blob_bytes) CUDA_TRY(cudaMemcpyAsync(e->d_blob.p, blob, blob_bytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->d_offs.p, offsets, (n + 1) * sizeof(u32), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(tok.p, 0, static_cast<size_t>(max_tok) * n * sizeof(u32), s));
    CUDA_TRY(cudaStreamWaitEvent(s, e->ev_flush, 0));
    k_tokenize<<<(static_cast<u32>(n) + TOK_THREADS - 1) / TOK_THREADS, TOK_THREADS, 0, s>>>(
        e->d_blob.as<u8>(), static_cast<u32>(blob_bytes), e->d_offs.as<u32>(), nullptr, static_cast<u32>(n), e->dev_view, max_tok, tok8.as<u32>(), tok.as<u32>(), meta.as<u32>(), stat.as<int>(), nullptr, nullptr, 0u, 0u);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_tokens, tok.p, static_cast<size_t>(max_tok) * n * sizeof(u32), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_meta, meta.p, n * sizeof(u32), cudaMemcpyDeviceToHost, s));
    std::vector<u32> rows(static_cast<size_t>(n) * TOK8);
    CUDA_TRY(cudaMemcpyAsync(rows.data(), tok8.p, rows.size() * sizeof(u32), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (uint64_t i = 0; i < n; ++i)          // levels 0..7 live in the per-topic row; present them level-major too
        for (u32 l = 0; l < TOK8 && l < max_tok; ++l) out_tokens[static_cast<size_t>(l) * n + i] = rows[i * TOK8 + l];
    return GM_OK;
}

int32_t gm_get_stats(gm_engine* e, gm_stats* out) {
    if (!e || !out) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    std::memset(out, 0, sizeof(*out));
    const HostTrie& t = e->trie;
    out->values = t.values_size(); out->nodes = t.nodes_size(); out->device_nodes = t.node_count() - 1;
    out->edges = t.edge_count(); out->edge_slots = t.edges.size();
    out->dict_entries = t.dict_count(); out->dict_slots = t.dict.size();
    out->plus_nodes = t.plus_count();
    out->value_words = t.values.size(); out->garbage_value_words = t.garbage_values;
    out->device_bytes = e->d_edges.cap + e->d_ranges.cap + e->d_values.cap + e->d_dict.cap + e->d_pool.cap;
    out->max_depth = t.max_depth;
    out->pending = (t.any_dirty() || e->rtree.dirty) ? 1 : 0;
    out->retained_values = e->rtree.values_size(); out->retained_nodes = e->rtree.nodes_size();
    out->device_bytes += e->d_rnodes.cap + e->d_rkids.cap + e->d_redges.cap + e->d_rvals.cap;
    return GM_OK;
}

int32_t gm_kernel_ms_ring(gm_engine* e, float* out_ms, uint32_t max_calls, uint32_t* n_calls) {
    if (!e || !out_ms || !n_calls) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    *n_calls = 0;
    if (e->flags & GM_FLAG_HOST_ONLY) return GM_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    const u64 have = std::min<u64>(e->ring_n, gm_engine::RING);
    const u64 take = std::min<u64>(have, max_calls);
    for (u64 k = 0; k < take; ++k) {                       // oldest first
        cudaEvent_t* ev = e->ev_ring[(e->ring_n - take + k) % gm_engine::RING];
        CUDA_TRY(cudaEventSynchronize(ev[3]));
        for (int j = 0; j < 3; ++j) CUDA_TRY(cudaEventElapsedTime(&out_ms[3 * k + j], ev[j], ev[j + 1]));
    }
    *n_calls = static_cast<uint32_t>(take);
    return GM_OK;
}

uint64_t gm_kernel_launches(gm_engine* e) { return e ? e->launches : 0; }

int32_t gm_debug_table(gm_engine* e, uint32_t which, const void** ptr, uint64_t* count) {
    if (!e || !ptr || !count) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    e->trie.sync();
    const HostTrie& t = e->trie;
    static thread_local uint32_t root[8];
    static thread_local uint64_t rstat[6];
    switch (which) {
        case 0: *ptr = t.edges.data(); *count = t.edges.size(); break;
        case 2: *ptr = t.ranges.data(); *count = t.ranges.size(); break;
        case 3: *ptr = t.values.data(); *count = t.values.size(); break;
        case 4: *ptr = t.dict.data(); *count = t.dict.size(); break;
        case 5: *ptr = t.pool.data(); *count = t.pool.size(); break;
        case 6: root[0] = t.root_plus; root[1] = t.root_hash_ref; root[2] = t.root_mask; root[3] = t.max_depth; root[4] = t.root_hash_cnt;
                root[5] = t.win_mask(); root[6] = t.win_shift(); root[7] = t.nwin_mask(); *ptr = root; *count = 8; break;
        case 12: *ptr = t.cfilter.data(); *count = t.cfilter.size(); break;
        case 10: e->rtree.prepare_flush(); *ptr = e->rtree.redges.data(); *count = e->rtree.redges.size(); break;
        case 11: e->rtree.debug_stats(rstat); *ptr = rstat; *count = 6; break;
        case 7: e->rtree.prepare_flush(); *ptr = e->rtree.rnodes.data(); *count = e->rtree.rnodes.size(); break;
        case 8: e->rtree.prepare_flush(); *ptr = e->rtree.rkids.data(); *count = e->rtree.rkids.size(); break;
        case 9: e->rtree.prepare_flush(); *ptr = e->rtree.rvals.data(); *count = e->rtree.rvals.size(); break;
        default: return GM_ERR_INVALID_ARG;
    }
    return GM_OK;
}

int32_t gm_debug_knob(gm_engine* e, const char* name, int64_t value) {
    if (!e || !name) return GM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(e->mu);
    const std::string k(name);
    if (k == "tile_chunk" && value >= 1 && value <= 1024) e->knobs.tile_chunk = static_cast<u32>(value);
    else if (k == "k2_ctas" && value >= 0 && value <= 8) e->knobs.k2_ctas = static_cast<int>(value);
    else if (k == "sorted_rows") e->knobs.sorted_rows = value != 0;
    else if (k == "bucket_bits" && value / 100 >= 10 && value % 100 + value / 100 <= int64_t(MAX_BUCKET_BITS)) { e->knobs.site_bits = static_cast<u32>(value / 100); e->knobs.sub_bits = static_cast<u32>(value % 100); }
    else if (k == "diag_flags") e->knobs.diag_flags = static_cast<u32>(value);
    else return GM_ERR_INVALID_ARG;
    return GM_OK;
}

uint32_t gm_shard_of(const char* s, uint32_t len, uint32_t n_shards) {
    if (!s || n_shards == 0) return 0xFFFFFFFFu;
    uint32_t l0 = 0;
    while (l0 < len && s[l0] != '/') ++l0;
    if (l0 == 1 && (s[0] == '+' || s[0] == '#')) return 0xFFFFFFFFu;
    return shard_of_hash(HostTrie::level0_hash(s, len), n_shards);
}

int32_t gm_shard_of_batch(const char* blob, const uint32_t* offsets, uint64_t n, uint32_t n_shards, uint32_t* out_shard) {
    if ((n && (!blob || !offsets || !out_shard)) || n_shards == 0) return GM_ERR_INVALID_ARG;
    for (uint64_t i = 0; i < n; ++i) out_shard[i] = gm_shard_of(blob + offsets[i], offsets[i + 1] - offsets[i], n_shards);
    return GM_OK;
}

// ---- NUMA placement of the host side (2-socket GPU servers: a pinned buffer on the far socket halves the PCIe rate and
//      makes 8 ranks contend for the inter-socket link) ------------------------------------------------------------------
static int device_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c) *c = static_cast<char>(tolower(*c));
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

int32_t gm_device_numa_node(int32_t device) {
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return -1;
    return device_numa_node(device);
}

int32_t gm_bind_thread_near_device(int32_t device) {
    const int node = gm_device_numa_node(device);
    if (node < 0) return GM_OK;                       // no NUMA information: nothing to do
    const std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return GM_OK;
    char buf[4096] = {0};
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return GM_OK;
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed); CPU_ZERO(&want);
    sched_getaffinity(0, sizeof(allowed), &allowed);
    int n_want = 0;
    for (char* p = buf; *p;) {
        char* e;
        long a = strtol(p, &e, 10), b = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long k = a; k <= b && k < CPU_SETSIZE; ++k) if (CPU_ISSET(k, &allowed)) { CPU_SET(k, &want); ++n_want; }
        if (*e != ',') break;
        p = e + 1;
    }
    if (n_want) sched_setaffinity(0, sizeof(want), &want);
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(long))] |= 1ul << (node % (8 * sizeof(long)));
    syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, static_cast<unsigned long>(node + 2));
    return GM_OK;
}

void* gm_host_alloc_near(gm_engine* e, uint64_t bytes) {
    const int node = (e && !(e->flags & GM_FLAG_HOST_ONLY)) ? device_numa_node(e->device) : -1;
    unsigned long mask[16] = {0};
    if (node >= 0) {
        mask[node / (8 * sizeof(long))] |= 1ul << (node % (8 * sizeof(long)));
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, static_cast<unsigned long>(node + 2));
    }
    void* p = nullptr;
    if (e && !(e->flags & GM_FLAG_HOST_ONLY)) cudaSetDevice(e->device);
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) p = nullptr;
    if (node >= 0) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
    return p;
}

void* gm_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void gm_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
