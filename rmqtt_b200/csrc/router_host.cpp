// GpuRouter — see router_host.h.
#include "router_host.h"

#include <algorithm>

namespace gm {

int32_t GpuRouter::add(const char* filter, uint32_t len, const Id& id, const Opts& opts) {
    std::string f(filter, len);
    uint32_t fi;
    auto fit = filter_index_.find(f);
    const bool new_filter_name = fit == filter_index_.end();
    fi = new_filter_name ? static_cast<uint32_t>(filter_names_.size()) : fit->second;
    auto rit = relations_.find(fi);
    uint32_t handle = 0;
    bool existing = false;
    if (rit != relations_.end()) {
        auto cit = rit->second.find(id.client_id);
        if (cit != rit->second.end()) { handle = cit->second; existing = true; }
    }
    if (!existing) {
        if (!free_handles_.empty()) { handle = free_handles_.back(); }
        else handle = static_cast<uint32_t>(by_handle_.size());
    }
    // Topic::from_str + topics.insert (router.rs:419-421): an invalid filter is an Err before any state changes
    int32_t changed = 0;
    int32_t rc = gm_sub_add(e_, filter, len, handle, &changed);
    if (rc != GM_OK) return rc;
    if (new_filter_name) { filter_index_.emplace(f, fi); filter_names_.push_back(f); }
    if (!existing) {
        if (!free_handles_.empty()) free_handles_.pop_back();
        else by_handle_.emplace_back();
    }
    auto& clients = relations_[fi];
    if (clients.empty()) topics_++;                                  // or_insert_with(|| topics_count.inc()) (router.rs:426-429)
    Rel& r = by_handle_[handle];
    r.filter_idx = fi; r.client = id.client_id; r.id = id; r.opts = opts; r.live = true;
    if (clients.emplace(id.client_id, handle).second) routes_++;     // HashMap::insert: replace keeps the count (router.rs:430-433)
    return GM_OK;
}

int32_t GpuRouter::remove(const char* filter, uint32_t len, const Id& id, bool* removed) {
    if (removed) *removed = false;
    auto fit = filter_index_.find(std::string(filter, len));
    if (fit == filter_index_.end()) return GM_OK;
    auto rit = relations_.find(fit->second);
    if (rit == relations_.end()) return GM_OK;
    auto cit = rit->second.find(id.client_id);
    if (cit == rit->second.end()) return GM_OK;
    const uint32_t handle = cit->second;
    if (!(by_handle_[handle].id == id)) return GM_OK;                 // "input id not the same" (router.rs:444-451)
    rit->second.erase(cit);
    routes_--;
    if (rit->second.empty()) { relations_.erase(rit); topics_--; }    // router.rs:466-473
    by_handle_[handle].live = false;
    free_handles_.push_back(handle);
    int32_t changed = 0;
    int32_t rc = gm_sub_remove(e_, filter, len, handle, &changed);
    if (rc != GM_OK) return rc;
    if (removed) *removed = true;
    return GM_OK;
}

bool GpuRouter::relation(uint32_t handle, const std::string** filter, const std::string** client) const {
    if (handle >= by_handle_.size() || !by_handle_[handle].live) return false;
    *filter = &filter_names_[by_handle_[handle].filter_idx];
    *client = &by_handle_[handle].client;
    return true;
}

int32_t GpuRouter::matches_batch(const gm_id* publishers, const char* blob, const uint32_t* offs, uint64_t n, std::vector<gm_span>& spans,
                                 std::vector<gm_sub_relation>& rels, std::vector<uint32_t>& sub_ids, std::vector<int32_t>& status) {
    spans.assign(n, gm_span{0, 0});
    status.assign(n, 0);
    rels.clear(); sub_ids.clear();
    if (n == 0) return GM_OK;
    tmp_spans_.assign(n, gm_span{0, 0});
    uint64_t needed = 0;
    if (tmp_ids_.size() < 1024) tmp_ids_.resize(std::max<size_t>(1024, 32 * n));
    for (;;) {
        int32_t rc = gm_match_batch(e_, blob, offs, n, tmp_spans_.data(), tmp_ids_.data(), tmp_ids_.size(), &needed, status.data());
        if (rc == GM_ERR_CAPACITY) { tmp_ids_.resize(needed + 1024); continue; }
        if (rc != GM_OK) return rc;
        break;
    }
    struct V5Entry { size_t rel_pos; std::vector<uint32_t> ids; };
    for (uint64_t i = 0; i < n; ++i) {
        const size_t begin = rels.size();
        if (status[i] == 0) {
            const Id pub{publishers ? publishers[i].node_id : 0, publishers && publishers[i].client_id ? std::string(publishers[i].client_id, publishers[i].client_len) : std::string(),
                         publishers ? publishers[i].tag : 0};
            std::unordered_map<std::string, V5Entry> v5;     // key: node id + client id (one collector per node, types.rs:466)
            const uint32_t* hs = tmp_ids_.data() + tmp_spans_[i].off;
            for (uint32_t k = 0; k < tmp_spans_[i].cnt; ++k) {
                const uint32_t h = hs[k];
                if (h >= by_handle_.size() || !by_handle_[h].live) continue;
                const Rel& r = by_handle_[h];
                if (r.opts.is_v5 && r.opts.no_local && publishers && pub == r.id) continue;                      // router.rs:184-189
                if (!r.opts.group.empty()) {                                                                      // router.rs:192-200
                    std::string key = filter_names_[r.filter_idx]; key.push_back('\0'); key += r.opts.group;
                    auto g = group_index_.emplace(key, static_cast<uint32_t>(group_index_.size() + 1)).first->second;
                    rels.push_back(gm_sub_relation{r.id.node_id, h, g, 0, 0});
                } else if (!r.opts.is_v5) {                                                                       // types.rs:486-487
                    rels.push_back(gm_sub_relation{r.id.node_id, h, 0, 0, 0});
                } else {                                                                                          // types.rs:488-506
                    std::string key = std::to_string(r.id.node_id); key.push_back('\0'); key += r.client;
                    auto it = v5.find(key);
                    if (it == v5.end()) {
                        V5Entry e{rels.size(), {}};
                        if (r.opts.sub_id) e.ids.push_back(r.opts.sub_id);
                        rels.push_back(gm_sub_relation{r.id.node_id, h, 0, 0, 0});
                        v5.emplace(std::move(key), std::move(e));
                    } else if (r.opts.sub_id) it->second.ids.push_back(r.opts.sub_id);
                }
            }
            for (auto& kv : v5) {
                gm_sub_relation& sr = rels[kv.second.rel_pos];
                sr.sub_ids_off = static_cast<uint32_t>(sub_ids.size());
                sr.sub_ids_cnt = static_cast<uint32_t>(kv.second.ids.size());
                sub_ids.insert(sub_ids.end(), kv.second.ids.begin(), kv.second.ids.end());
            }
        }
        spans[i] = gm_span{static_cast<uint32_t>(begin), static_cast<uint32_t>(rels.size() - begin)};
    }
    return GM_OK;
}

}  // namespace gm
