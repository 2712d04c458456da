// GpuRouter — host-side mirror of rmqtt's DefaultRouter (rmqtt/src/router.rs:109-115) above the matching
// engine: the reference toolchain (Rust) is absent, so the Router-level semantics a Rust `GpuRouter` plugin
// would keep are implemented here in C++ and driven through the gmr_* C entry points.
//
//   state     relations: filter -> client -> (Id, SubscriptionOptions)         router.rs:113  (AllRelationsMap, types.rs:443)
//             topics / routes counters                                          router.rs:112,114
//             the trie of the reference (`topics: TopicTree<()>`) is the device-resident trie of the engine;
//             its values are relation handles (one u32 per (filter, client))
//   add       router.rs:417-436        remove    router.rs:439-479 (Id-equality rule)
//   matches   router.rs:162-248 for a batch of PUBLISHes: engine match -> relation lookup -> no_local (:184-189)
//             -> shared-group bucketing (:192-200) -> SubscriptioRelationsCollector::add (types.rs:478-508:
//             v3 one relation per (filter, client); v5 per-client de-dup, subscription identifiers accumulate)
//             The shared-subscription *choice* is rand::random in the reference (subscribe.rs:88): members are
//             returned, not chosen.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gpumqtt.h"

namespace gm {

class GpuRouter {
  public:
    struct Id { uint64_t node_id = 0; std::string client_id; uint64_t tag = 0;
                bool operator==(const Id& o) const { return node_id == o.node_id && tag == o.tag && client_id == o.client_id; } };   // types.rs:1746-1757
    struct Opts { uint8_t qos = 0, is_v5 = 0, no_local = 0; uint32_t sub_id = 0; std::string group; };                          // types.rs:565-718

    explicit GpuRouter(gm_engine* e) : e_(e) {}

    int32_t add(const char* filter, uint32_t len, const Id& id, const Opts& opts);
    int32_t remove(const char* filter, uint32_t len, const Id& id, bool* removed);
    int64_t topics() const { return topics_; }                // Router::topics  (router.rs:554-556)
    int64_t routes() const { return routes_; }                // Router::routes
    int32_t matches_batch(const gm_id* publishers, const char* blob, const uint32_t* offs, uint64_t n, std::vector<gm_span>& spans,
                          std::vector<gm_sub_relation>& rels, std::vector<uint32_t>& sub_ids, std::vector<int32_t>& status);
    bool relation(uint32_t handle, const std::string** filter, const std::string** client) const;
    std::mutex mu;

  private:
    struct Rel { uint32_t filter_idx = 0; std::string client; Id id; Opts opts; bool live = false; };
    gm_engine* e_;
    std::unordered_map<std::string, uint32_t> filter_index_;
    std::vector<std::string> filter_names_;
    std::unordered_map<uint32_t, std::unordered_map<std::string, uint32_t>> relations_;   // filter idx -> client -> handle
    std::unordered_map<std::string, uint32_t> group_index_;                                // "filter\0group" -> 1-based id
    std::vector<Rel> by_handle_;
    std::vector<uint32_t> free_handles_;
    int64_t topics_ = 0, routes_ = 0;
    std::vector<gm_span> tmp_spans_;
    std::vector<uint32_t> tmp_ids_;
};

}  // namespace gm
