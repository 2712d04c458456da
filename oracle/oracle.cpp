// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// C++17 restatement of the reference algorithm for the north-star hot path of rmqtt
// (reference commit 4f9f2185).  The reference is Rust and cannot be compiled in this image
// (no cargo/rustc), so this file re-states its algorithm, structure-faithfully, on the CPU:
//
//   * Level / Topic parsing + validation ......... rmqtt/src/topic.rs:89-96, 204-216, 326-363
//   * string matcher Topic::matches_str .......... rmqtt/src/topic.rs:167-185, 315-324
//   * TopicTree<V> (Node{values,branches}) ....... rmqtt/src/trie.rs:69-73
//       insert / remove (with pruning) ........... rmqtt/src/trie.rs:99-135
//       matches (MatchedIter::prepare) ........... rmqtt/src/trie.rs:299-347
//   * DefaultRouter add / remove / _matches ...... rmqtt/src/router.rs:162-248, 417-479
//   * RetainTree<V> insert/remove/_matches ....... rmqtt/src/retain.rs:221-257, 298-367
//
// Containers follow the reference: children live in a hash map keyed by Level (enum + string),
// values in an ordered set (BTreeSet -> std::set).  Iteration order of a hash map is unspecified in
// both, so every comparison made with this oracle is on SORTED MULTISETS.
//
// Parity status: PINNED.  tests/test_oracle_golden.py replays every assertion of the reference's
// own tests for this path (trie.rs:415-513, retain.rs:449-482, topic.rs:429-586) against this file.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
// this library.  It is "a C++ restatement of the reference algorithm", never "the reference binary".
#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <set>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace orc {

// ---------------------------------------------------------------- topic.rs:89-96
enum class Kind : uint8_t { Normal = 0, Metadata = 1, Blank = 2, Single = 3, Multi = 4 };

struct Level {
    Kind kind;
    std::string s;  // payload for Normal / Metadata, empty otherwise
    bool operator==(const Level& o) const { return kind == o.kind && s == o.s; }
    bool is_metadata() const { return kind == Kind::Metadata; }  // topic.rs:153-155
};

struct LevelHash {
    size_t operator()(const Level& l) const {
        return std::hash<std::string>()(l.s) * 1315423911u + static_cast<size_t>(l.kind);
    }
};

static inline bool has_wild(std::string_view s) {
    return s.find('+') != std::string_view::npos || s.find('#') != std::string_view::npos;
}

// topic.rs:326-346  Level::from_str
static bool parse_level(std::string_view s, Level& out) {
    if (s == "+") { out = {Kind::Single, {}}; return true; }
    if (s == "#") { out = {Kind::Multi, {}}; return true; }
    if (s.empty()) { out = {Kind::Blank, {}}; return true; }
    if (has_wild(s)) return false;                                   // InvalidLevel
    if (s[0] == '$') { out = {Kind::Metadata, std::string(s)}; return true; }
    out = {Kind::Normal, std::string(s)};
    return true;
}

// topic.rs:140-146 Level::is_valid
static bool level_valid(const Level& l) {
    switch (l.kind) {
        case Kind::Normal: return !(l.s.size() && l.s[0] == '$') && !has_wild(l.s);
        case Kind::Metadata: return l.s.size() && l.s[0] == '$' && !has_wild(l.s);
        default: return true;
    }
}

using Topic = std::vector<Level>;

// topic.rs:204-216 Topic::is_valid
static bool topic_valid(const Topic& t) {
    for (auto& l : t) if (!level_valid(l)) return false;
    for (size_t pos = 0; pos < t.size(); ++pos) {
        if (t[pos].kind == Kind::Multi && pos != t.size() - 1) return false;
        if (t[pos].kind == Kind::Metadata && pos != 0) return false;
    }
    return true;
}

// topic.rs:348-363 Topic::from_str   (str::split('/') yields [""] for the empty string)
static bool parse_topic(std::string_view s, Topic& out) {
    out.clear();
    size_t start = 0;
    for (;;) {
        size_t p = s.find('/', start);
        std::string_view lv = s.substr(start, p == std::string_view::npos ? std::string_view::npos : p - start);
        Level l;
        if (!parse_level(lv, l)) return false;
        out.push_back(std::move(l));
        if (p == std::string_view::npos) break;
        start = p + 1;
    }
    return topic_valid(out);
}

// topic.rs:365-392 Display
static void level_to_string(const Level& l, std::string& o) {
    switch (l.kind) {
        case Kind::Normal: case Kind::Metadata: o += l.s; break;
        case Kind::Blank: break;
        case Kind::Single: o += '+'; break;
        case Kind::Multi: o += '#'; break;
    }
}

// topic.rs:315-324  impl<T: AsRef<str>> MatchLevel for T
static bool str_match_level(std::string_view self, const Level& level) {
    bool meta = !self.empty() && self[0] == '$';
    switch (level.kind) {
        case Kind::Normal: return !meta && level.s == self;
        case Kind::Metadata: return meta && level.s == self;
        case Kind::Blank: return self.empty();
        default: return !meta;
    }
}

// topic.rs:167-185 (macro matches!) + :223-224 matches_str — note the `break` on a failed `+`.
static bool matches_str(const Topic& filter, std::string_view topic) {
    size_t li = 0;  // lhs iterator position
    size_t start = 0;
    bool done = false;
    while (!done) {
        size_t p = topic.find('/', start);
        std::string_view rhs = topic.substr(start, p == std::string_view::npos ? std::string_view::npos : p - start);
        if (p == std::string_view::npos) done = true; else start = p + 1;
        if (li >= filter.size()) return false;          // lhs.next() == None -> `_ => return false`
        const Level& lhs = filter[li++];
        if (lhs.kind == Kind::Single) {
            if (!str_match_level(rhs, lhs)) break;
        } else if (lhs.kind == Kind::Multi) {
            return str_match_level(rhs, lhs);
        } else if (str_match_level(rhs, lhs)) {
            continue;
        } else {
            return false;
        }
    }
    if (li < filter.size()) return filter[li].kind == Kind::Multi;
    return true;
}

// ---------------------------------------------------------------- counters
struct Counters {
    uint64_t V = 0;  // visited trie nodes  (= MatchedIter::prepare calls)
    uint64_t E = 0;  // visited nodes with a non-empty remaining path (literal probe issued)
    uint64_t F = 0;  // matched filter nodes (items yielded)
    uint64_t M = 0;  // matched values
    uint64_t L = 0;  // levels
    uint64_t B = 0;  // topic bytes
    void add(const Counters& o) { V += o.V; E += o.E; F += o.F; M += o.M; L += o.L; B += o.B; }
};

// ---------------------------------------------------------------- trie.rs:69-73
template <class V>
struct TrieNode {
    std::set<V> values;                                                         // BTreeSet<V>
    std::unordered_map<Level, std::unique_ptr<TrieNode>, LevelHash> branches;   // HashMap<Level,Node>

    // trie.rs:99-112
    bool insert(const Topic& path, size_t i, const V& v) {
        if (i < path.size()) {
            auto& slot = branches[path[i]];
            if (!slot) slot = std::make_unique<TrieNode>();
            return slot->insert(path, i + 1, v);
        }
        return values.insert(v).second;
    }
    // trie.rs:115-135
    bool remove(const Topic& path, size_t i, const V& v) {
        if (i == path.size()) return values.erase(v) > 0;
        auto it = branches.find(path[i]);
        if (it == branches.end()) return false;
        bool res = it->second->remove(path, i + 1, v);
        if (it->second->values.empty() && it->second->branches.empty()) branches.erase(it);
        return res;
    }
    size_t values_size() const {  // trie.rs:148-151
        size_t n = values.size();
        for (auto& kv : branches) n += kv.second->values_size();
        return n;
    }
    size_t nodes_size() const {   // trie.rs:154-157
        size_t n = branches.size();
        for (auto& kv : branches) n += kv.second->nodes_size();
        return n;
    }
};

static const Level kMulti{Kind::Multi, {}};
static const Level kSingle{Kind::Single, {}};

// trie.rs:299-347  MatchedIter::prepare, restated as a recursion (the reference is a lazy DFS whose
// yield order is unspecified anyway: LIFO curr_items + hash-map iteration).
// `emit(filter_path, values)` is called once per yielded item.
template <class V, class Emit>
static void trie_walk(const TrieNode<V>* node, const Topic& path, size_t i,
                      std::vector<const Level*>& sub_path, Counters& c, Emit&& emit) {
    c.V++;
    auto find = [&](const Level& l) -> const TrieNode<V>* {
        auto it = node->branches.find(l);
        return it == node->branches.end() ? nullptr : it->second.get();
    };
    if (i == path.size()) {
        // Match parent #   (trie.rs:301-308)
        if (const TrieNode<V>* b = find(kMulti)) {
            if (!b->values.empty()) {
                sub_path.push_back(&kMulti);
                c.F++; c.M += b->values.size();
                emit(sub_path, b->values);
                sub_path.pop_back();
            }
        }
        if (!node->values.empty()) {     // add_to_items skips empty sets (trie.rs:278-282)
            c.F++; c.M += node->values.size();
            emit(sub_path, node->values);
        }
        return;
    }
    c.E++;
    const TrieNode<V>* multi = find(kMulti);
    const TrieNode<V>* single = find(kSingle);
    // `$`-rule  (trie.rs:312-318)
    bool skip = sub_path.empty() && path[i].kind != Kind::Blank && path[i].is_metadata() && (multi || single);
    if (!skip) {
        if (multi && !multi->values.empty()) {          // trie.rs:321-327
            sub_path.push_back(&kMulti);
            c.F++; c.M += multi->values.size();
            emit(sub_path, multi->values);
            sub_path.pop_back();
        }
        if (single) {                                    // trie.rs:330-334
            sub_path.push_back(&kSingle);
            trie_walk(single, path, i + 1, sub_path, c, emit);
            sub_path.pop_back();
        }
    }
    if (const TrieNode<V>* b = find(path[i])) {          // trie.rs:338-342 precise matching
        sub_path.push_back(&path[i]);
        trie_walk(b, path, i + 1, sub_path, c, emit);
        sub_path.pop_back();
    }
}

// ---------------------------------------------------------------- router.rs (DefaultRouter)
struct Relation {                 // stands for (Id, SubscriptionOptions)
    uint32_t rel_id; uint64_t id_tag;
    uint64_t node_id = 0; bool is_v5 = false, no_local = false; uint32_t sub_id = 0; std::string group;
};

struct Router {
    TrieNode<char> topics;                                                       // TopicTree<()>
    std::unordered_map<std::string, std::unordered_map<std::string, Relation>> relations;  // DashMap<TopicFilter, HashMap<ClientId,(Id,opts)>>
    int64_t topics_count = 0, relations_count = 0;

    // router.rs:417-436
    bool add(std::string_view filter, std::string_view client, uint32_t rel_id, uint64_t id_tag) {
        Topic t;
        if (!parse_topic(filter, t)) return false;
        topics.insert(t, 0, 0);
        auto it = relations.find(std::string(filter));
        if (it == relations.end()) { topics_count++; it = relations.emplace(std::string(filter), std::unordered_map<std::string, Relation>()).first; }
        auto ins = it->second.insert_or_assign(std::string(client), Relation{rel_id, id_tag, 0, false, false, 0, {}});
        if (ins.second) relations_count++;
        return true;
    }
    bool add_full(std::string_view filter, std::string_view client, const Relation& rel) {
        if (!add(filter, client, rel.rel_id, rel.id_tag)) return false;
        relations[std::string(filter)][std::string(client)] = rel;
        return true;
    }
    // router.rs:162-248 + types.rs:470-508, canonicalised: one text line per result element, sorted.
    //   "3|node|filter|client"            v3 relation
    //   "5|node|client|id,id,..."         v5 relation after per-client de-dup, subscription identifiers sorted
    //   "g|filter|group|node:client;..."  members of one shared group of one matched filter (the reference picks one at random)
    bool matches_full(std::string_view topic, uint64_t pub_node, std::string_view pub_client, uint64_t pub_tag, std::vector<std::string>& lines) const {
        Topic t;
        if (!parse_topic(topic, t)) return false;
        Counters c;
        std::vector<const Level*> sp;
        std::string fs;
        std::unordered_map<std::string, std::vector<uint32_t>> v5;   // "node|client" -> sub ids
        trie_walk(&topics, t, 0, sp, c, [&](const std::vector<const Level*>& fp, const std::set<char>&) {
            fs.clear();
            for (size_t k = 0; k < fp.size(); ++k) { if (k) fs += '/'; level_to_string(*fp[k], fs); }
            auto it = relations.find(fs);
            if (it == relations.end()) return;
            std::unordered_map<std::string, std::vector<std::string>> groups;
            for (auto& kv : it->second) {
                const Relation& r = kv.second;
                if (r.is_v5 && r.no_local && r.node_id == pub_node && r.id_tag == pub_tag && kv.first == pub_client) continue;   // router.rs:184-189
                if (!r.group.empty()) { groups[r.group].push_back(std::to_string(r.node_id) + ":" + kv.first); continue; }
                if (!r.is_v5) lines.push_back("3|" + std::to_string(r.node_id) + "|" + fs + "|" + kv.first);
                else {
                    auto& ids = v5[std::to_string(r.node_id) + "|" + kv.first];
                    if (r.sub_id) ids.push_back(r.sub_id);
                }
            }
            for (auto& g : groups) {
                std::sort(g.second.begin(), g.second.end());
                std::string l = "g|" + fs + "|" + g.first + "|";
                for (size_t k = 0; k < g.second.size(); ++k) { if (k) l += ';'; l += g.second[k]; }
                lines.push_back(l);
            }
        });
        for (auto& kv : v5) {
            std::sort(kv.second.begin(), kv.second.end());
            std::string l = "5|" + kv.first + "|";
            for (size_t k = 0; k < kv.second.size(); ++k) { if (k) l += ','; l += std::to_string(kv.second[k]); }
            lines.push_back(l);
        }
        std::sort(lines.begin(), lines.end());
        return true;
    }
    // Secondary readers of the same tree, canonicalised as sorted text lines:
    //   kind 0  _has_matches (router.rs:139-142 -> trie.rs:138-140): one line "1" if any filter matches
    //   kind 1  _get_routes  (router.rs:145-158): "filter" per UNIQUE matched filter (the caller adds its own node id)
    //   kind 2  get          (router.rs:522-546): "node|filter" per unique matched filter and unique node id among its relations
    bool readers(std::string_view topic, int kind, std::vector<std::string>& lines) const {
        Topic t;
        if (!parse_topic(topic, t)) return false;
        Counters c;
        std::vector<const Level*> sp;
        std::set<std::string> uniq;                                  // itertools .unique() over the matched filter paths
        std::string fs;
        trie_walk(&topics, t, 0, sp, c, [&](const std::vector<const Level*>& fp, const std::set<char>&) {
            fs.clear();
            for (size_t k = 0; k < fp.size(); ++k) { if (k) fs += '/'; level_to_string(*fp[k], fs); }
            uniq.insert(fs);
        });
        if (kind == 0) { if (!uniq.empty()) lines.push_back("1"); return true; }
        for (const std::string& f : uniq) {
            if (kind == 1) { lines.push_back(f); continue; }
            auto it = relations.find(f);
            if (it == relations.end()) continue;
            std::set<uint64_t> nodes;
            for (auto& kv : it->second) nodes.insert(kv.second.node_id);
            for (uint64_t nid : nodes) lines.push_back(std::to_string(nid) + "|" + f);
        }
        std::sort(lines.begin(), lines.end());
        return true;
    }
    // router.rs:439-479  (returns 1 removed, 0 not removed, -1 invalid filter on the prune path)
    int remove(std::string_view filter, std::string_view client, uint64_t id_tag) {
        auto it = relations.find(std::string(filter));
        if (it == relations.end()) return 0;
        auto cit = it->second.find(std::string(client));
        if (cit == it->second.end() || cit->second.id_tag != id_tag) return 0;
        it->second.erase(cit);
        relations_count--;
        if (it->second.empty()) {
            relations.erase(it);
            topics_count--;
            Topic t;
            if (!parse_topic(filter, t)) return -1;
            topics.remove(t, 0, 0);
        }
        return 1;
    }
    // router.rs:162-248 up to (and excluding) no_local / shared-group choice / v5 de-dup:
    // parse -> trie walk -> per matched filter rebuild the string (trie.rs:248-257) -> relations.get
    // -> one output per (filter, client) relation.
    bool matches(std::string_view topic, std::vector<uint32_t>& out, Counters& c) const {
        Topic t;
        if (!parse_topic(topic, t)) return false;
        c.L += t.size(); c.B += topic.size();
        std::vector<const Level*> sp;
        std::string fs;
        trie_walk(&topics, t, 0, sp, c, [&](const std::vector<const Level*>& fp, const std::set<char>&) {
            fs.clear();
            for (size_t k = 0; k < fp.size(); ++k) { if (k) fs += '/'; level_to_string(*fp[k], fs); }
            auto it = relations.find(fs);
            if (it != relations.end())
                for (auto& kv : it->second) out.push_back(kv.second.rel_id);
        });
        return true;
    }
};

// ---------------------------------------------------------------- retain.rs:204-207
template <class V>
struct RetainNode {
    bool has_value = false;
    V value{};
    std::unordered_map<Level, std::unique_ptr<RetainNode>, LevelHash> branches;

    void insert(const Topic& path, size_t i, const V& v) {       // retain.rs:221-234
        if (i < path.size()) {
            auto& slot = branches[path[i]];
            if (!slot) slot = std::make_unique<RetainNode>();
            slot->insert(path, i + 1, v);
        } else { has_value = true; value = v; }
    }
    bool remove(const Topic& path, size_t i, V* old) {           // retain.rs:237-257
        if (i == path.size()) { bool h = has_value; if (h && old) *old = value; has_value = false; return h; }
        auto it = branches.find(path[i]);
        if (it == branches.end()) return false;
        bool res = it->second->remove(path, i + 1, old);
        if (!it->second->has_value && it->second->branches.empty()) branches.erase(it);
        return res;
    }
    size_t values_size() const { size_t n = has_value; for (auto& kv : branches) n += kv.second->values_size(); return n; }
    size_t nodes_size() const { size_t n = branches.size(); for (auto& kv : branches) n += kv.second->nodes_size(); return n; }

    // retain.rs:298-367  _matches.  `depth0` stands for sub_path.is_empty().
    void matches(const Topic& path, size_t i, bool depth0, Counters& c, std::vector<V>& out) const {
        c.V++;
        if (branches.empty() || i == path.size()) {
            if (i == path.size() && has_value) { out.push_back(value); c.M++; }
            return;
        }
        c.E++;
        auto it = branches.find(path[i]);
        bool next_multi = (i + 1 < path.size()) && path[i + 1].kind == Kind::Multi;
        if (it != branches.end()) {                              // precise matching :313-323
            const RetainNode* r = it->second.get();
            if (next_multi && r->has_value) { out.push_back(r->value); c.M++; }
            r->matches(path, i + 1, false, c, out);
        } else if (path[i].kind == Kind::Single) {               // :324-342
            for (auto& kv : branches) {
                if (depth0 && kv.first.kind != Kind::Blank && kv.first.is_metadata()) continue;
                const RetainNode* v = kv.second.get();
                if (next_multi && v->has_value) { out.push_back(v->value); c.M++; }
                v->matches(path, i + 1, false, c, out);
            }
        } else if (path[i].kind == Kind::Multi) {                // :343-365
            for (auto& kv : branches) {
                if (depth0 && kv.first.kind != Kind::Blank && kv.first.is_metadata()) continue;
                const RetainNode* v = kv.second.get();
                if (v->branches.empty()) {
                    if (v->has_value) { out.push_back(v->value); c.M++; }
                } else {
                    if (v->has_value) { out.push_back(v->value); c.M++; }
                    v->matches(path, i, false, c, out);
                }
            }
        }
    }
};

using Tree = TrieNode<uint64_t>;
using RTree = RetainNode<int64_t>;

// CPUs this process may run on (cgroup cpuset / affinity mask), in order
static const std::vector<int>& allowed_cpus() {
    static const std::vector<int> cpus = [] {
        std::vector<int> v;
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0)
            for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &set)) v.push_back(c);
        return v;
    }();
    return cpus;
}
static void pin_self(int k) {
    const auto& cpus = allowed_cpus();
    if (cpus.empty()) return;
    cpu_set_t s;
    CPU_ZERO(&s);
    CPU_SET(cpus[static_cast<size_t>(k) % cpus.size()], &s);
    pthread_setaffinity_np(pthread_self(), sizeof(s), &s);
}

// Reader threads over a batch: one thread per allowed CPU (pinned, so the measurement does not depend on where the
// scheduler happens to put 128 threads), work handed out in small chunks from a shared counter (no straggler slice).
// f(tid, b, e) may be called many times per thread.
template <class F>
static void parallel_for(uint64_t n, int nthreads, F&& f) {
    if (nthreads <= 1 || n < 64) { f(0, 0, n); return; }
    std::vector<std::thread> th;
    std::atomic<uint64_t> next{0};
    const uint64_t grain = std::max<uint64_t>(16, std::min<uint64_t>(1024, n / (static_cast<uint64_t>(nthreads) * 16)));
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([=, &f, &next] {
            pin_self(t);
            for (;;) {
                const uint64_t b = next.fetch_add(grain, std::memory_order_relaxed);
                if (b >= n) break;
                f(t, b, std::min<uint64_t>(n, b + grain));
            }
        });
    }
    for (auto& t : th) t.join();
}

}  // namespace orc

using namespace orc;

extern "C" {

struct orc_counters { uint64_t V, E, F, M, L, B; };
static void put(orc_counters* o, const Counters& c) { if (o) { o->V = c.V; o->E = c.E; o->F = c.F; o->M = c.M; o->L = c.L; o->B = c.B; } }

// ---- parsing -----------------------------------------------------------------
// returns number of levels (kinds written, up to cap) or -1 if Topic::from_str would return Err.
int64_t orc_topic_parse(const char* s, uint32_t len, uint8_t* kinds, uint32_t cap) {
    Topic t;
    if (!parse_topic(std::string_view(s, len), t)) return -1;
    for (size_t i = 0; i < t.size() && i < cap; ++i) kinds[i] = static_cast<uint8_t>(t[i].kind);
    return static_cast<int64_t>(t.size());
}
// 1 / 0, or -1 if the filter does not parse
int32_t orc_matches_str(const char* f, uint32_t flen, const char* t, uint32_t tlen) {
    Topic ft;
    if (!parse_topic(std::string_view(f, flen), ft)) return -1;
    return matches_str(ft, std::string_view(t, tlen)) ? 1 : 0;
}

// ---- TopicTree<u64> ----------------------------------------------------------
void* orc_tree_new() { return new Tree(); }
void orc_tree_free(void* t) { delete static_cast<Tree*>(t); }
int32_t orc_tree_insert(void* tp, const char* f, uint32_t len, uint64_t v) {
    Topic t;
    if (!parse_topic(std::string_view(f, len), t)) return -1;
    return static_cast<Tree*>(tp)->insert(t, 0, v) ? 1 : 0;
}
int32_t orc_tree_remove(void* tp, const char* f, uint32_t len, uint64_t v) {
    Topic t;
    if (!parse_topic(std::string_view(f, len), t)) return -1;
    return static_cast<Tree*>(tp)->remove(t, 0, v) ? 1 : 0;
}
uint64_t orc_tree_values_size(void* tp) { return static_cast<Tree*>(tp)->values_size(); }
uint64_t orc_tree_nodes_size(void* tp) { return static_cast<Tree*>(tp)->nodes_size(); }

// Bulk insert.  Filters whose level-0 strings differ never share a node below the root, so the
// set is partitioned by hash(level 0) over `nthreads` private roots which are spliced afterwards —
// the resulting tree is identical to sequential insertion.
int64_t orc_tree_bulk_insert(void* tp, const char* blob, const uint32_t* offs, const uint32_t* vals, uint64_t n, int nthreads) {
    Tree* root = static_cast<Tree*>(tp);
    if (nthreads < 1) nthreads = 1;
    std::vector<Tree> parts(nthreads);
    std::vector<int64_t> ok(nthreads, 0);
    {
        std::vector<std::thread> th;
        for (int p = 0; p < nthreads; ++p) th.emplace_back([&, p] {
            Topic t;
            for (uint64_t i = 0; i < n; ++i) {
                std::string_view s(blob + offs[i], offs[i + 1] - offs[i]);
                size_t sl = s.find('/');
                std::string_view l0 = s.substr(0, sl);
                if (std::hash<std::string_view>()(l0) % static_cast<size_t>(nthreads) != static_cast<size_t>(p)) continue;
                if (!parse_topic(s, t)) continue;
                if (parts[p].insert(t, 0, vals[i])) ok[p]++;
            }
        });
        for (auto& t : th) t.join();
    }
    int64_t total = 0;
    for (int p = 0; p < nthreads; ++p) {
        total += ok[p];
        for (auto& kv : parts[p].branches) {
            auto it = root->branches.find(kv.first);
            if (it == root->branches.end()) root->branches.emplace(kv.first, std::move(kv.second));
            else {  // root child already present (pre-existing tree): fall back to re-insertion semantics
                std::function<void(TrieNode<uint64_t>*, TrieNode<uint64_t>*)> merge = [&](TrieNode<uint64_t>* dst, TrieNode<uint64_t>* src) {
                    for (auto& v : src->values) dst->values.insert(v);
                    for (auto& c : src->branches) {
                        auto d = dst->branches.find(c.first);
                        if (d == dst->branches.end()) dst->branches.emplace(c.first, std::move(c.second));
                        else merge(d->second.get(), c.second.get());
                    }
                };
                merge(it->second.get(), kv.second.get());
            }
        }
    }
    return total;
}

// Match one topic.  Writes up to cap values; returns total values, or -1 if the topic is invalid.
int64_t orc_tree_match(void* tp, const char* s, uint32_t len, uint64_t* out, uint64_t cap, orc_counters* ctr) {
    Topic t;
    if (!parse_topic(std::string_view(s, len), t)) return -1;
    Counters c; c.L = t.size(); c.B = len;
    std::vector<const Level*> sp;
    uint64_t n = 0;
    trie_walk(static_cast<Tree*>(tp), t, 0, sp, c, [&](const std::vector<const Level*>&, const std::set<uint64_t>& vs) {
        for (auto v : vs) { if (n < cap) out[n] = v; ++n; }
    });
    put(ctr, c);
    return static_cast<int64_t>(n);
}

// Batch match with `nthreads` readers over disjoint topic slices against the shared read-only tree
// (mirrors concurrent readers under the reference's RwLock, router.rs:166).
// counts[i] = number of values (or -1 invalid). If out_ids != NULL, out_offs (n+1, u64) must hold the
// exclusive prefix of max(counts,0) computed by a previous call with out_ids == NULL.
// Returns elapsed seconds of the matching loop.
double orc_tree_match_batch(void* tp, const char* blob, const uint32_t* offs, uint64_t n, int nthreads,
                            int64_t* counts, const uint64_t* out_offs, uint32_t* out_ids, orc_counters* ctr) {
    Tree* tree = static_cast<Tree*>(tp);
    std::vector<Counters> cs(std::max(1, nthreads));
    auto t0 = std::chrono::steady_clock::now();
    parallel_for(n, nthreads, [&](int tid, uint64_t b, uint64_t e) {
        Topic t; std::vector<const Level*> sp; Counters c;
        for (uint64_t i = b; i < e; ++i) {
            std::string_view s(blob + offs[i], offs[i + 1] - offs[i]);
            if (!parse_topic(s, t)) { if (counts) counts[i] = -1; continue; }
            c.L += t.size(); c.B += s.size();
            uint64_t k = 0; uint32_t* dst = out_ids ? out_ids + out_offs[i] : nullptr;
            sp.clear();
            trie_walk(tree, t, 0, sp, c, [&](const std::vector<const Level*>&, const std::set<uint64_t>& vs) {
                if (dst) for (auto v : vs) dst[k++] = static_cast<uint32_t>(v); else k += vs.size();
            });
            if (counts) counts[i] = static_cast<int64_t>(k);
        }
        cs[tid].add(c);
    });
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    Counters tot; for (auto& c : cs) tot.add(c);
    put(ctr, tot);
    return dt;
}

// ---- DefaultRouter restatement --------------------------------------------------
void* orc_router_new() { return new Router(); }
void orc_router_free(void* r) { delete static_cast<Router*>(r); }
int32_t orc_router_add(void* r, const char* f, uint32_t flen, const char* c, uint32_t clen, uint32_t rel_id, uint64_t id_tag) {
    return static_cast<Router*>(r)->add(std::string_view(f, flen), std::string_view(c, clen), rel_id, id_tag) ? 1 : -1;
}
int32_t orc_router_remove(void* r, const char* f, uint32_t flen, const char* c, uint32_t clen, uint64_t id_tag) {
    return static_cast<Router*>(r)->remove(std::string_view(f, flen), std::string_view(c, clen), id_tag);
}
int32_t orc_router_add_full(void* r, const char* f, uint32_t flen, const char* c, uint32_t clen, uint32_t rel_id, uint64_t id_tag, uint64_t node_id,
                            int32_t is_v5, int32_t no_local, uint32_t sub_id, const char* group, uint32_t glen) {
    Relation rel{rel_id, id_tag, 0, false, false, 0, {}};
    rel.node_id = node_id; rel.is_v5 = is_v5 != 0; rel.no_local = no_local != 0; rel.sub_id = sub_id;
    if (group) rel.group.assign(group, glen);
    return static_cast<Router*>(r)->add_full(std::string_view(f, flen), std::string_view(c, clen), rel) ? 1 : -1;
}
// canonical text (lines joined by '\n') into out (cap bytes); returns needed bytes, or -1 for an invalid topic
int64_t orc_router_match_full(void* r, const char* s, uint32_t len, uint64_t pub_node, const char* pub_client, uint32_t pclen, uint64_t pub_tag, char* out, uint64_t cap) {
    std::vector<std::string> lines;
    if (!static_cast<Router*>(r)->matches_full(std::string_view(s, len), pub_node, std::string_view(pub_client, pclen), pub_tag, lines)) return -1;
    std::string all;
    for (size_t i = 0; i < lines.size(); ++i) { if (i) all += '\n'; all += lines[i]; }
    if (all.size() <= cap) std::memcpy(out, all.data(), all.size());
    return static_cast<int64_t>(all.size());
}
// canonical text of a secondary reader (Router::readers) into out; returns needed bytes, or -1 for an invalid topic
int64_t orc_router_readers(void* r, const char* s, uint32_t len, int32_t kind, char* out, uint64_t cap) {
    std::vector<std::string> lines;
    if (!static_cast<Router*>(r)->readers(std::string_view(s, len), kind, lines)) return -1;
    std::string all;
    for (size_t i = 0; i < lines.size(); ++i) { if (i) all += '\n'; all += lines[i]; }
    if (all.size() <= cap) std::memcpy(out, all.data(), all.size());
    return static_cast<int64_t>(all.size());
}
int64_t orc_router_topics(void* r) { return static_cast<Router*>(r)->topics_count; }
int64_t orc_router_routes(void* r) { return static_cast<Router*>(r)->relations_count; }
uint64_t orc_router_topics_tree(void* r) { return static_cast<Router*>(r)->topics.values_size(); }
// bulk: subscription i = (filter i, client "c<i>" , rel_id = vals[i], id_tag = vals[i])
// Partitioned by hash(level 0) like orc_tree_bulk_insert (same argument: identical final state);
// only valid on an EMPTY router when nthreads > 1.
int64_t orc_router_bulk_add(void* rp, const char* blob, const uint32_t* offs, const uint32_t* vals, uint64_t n, int nthreads) {
    Router* r = static_cast<Router*>(rp);
    if (nthreads <= 1 || !r->relations.empty()) {
        int64_t ok = 0; char cid[32];
        for (uint64_t i = 0; i < n; ++i) {
            int cl = snprintf(cid, sizeof cid, "c%u", vals[i]);
            if (r->add(std::string_view(blob + offs[i], offs[i + 1] - offs[i]), std::string_view(cid, cl), vals[i], vals[i])) ok++;
        }
        return ok;
    }
    std::vector<Router> parts(nthreads);
    std::vector<int64_t> ok(nthreads, 0);
    std::vector<std::thread> th;
    for (int p = 0; p < nthreads; ++p) th.emplace_back([&, p] {
        char cid[32];
        for (uint64_t i = 0; i < n; ++i) {
            std::string_view s(blob + offs[i], offs[i + 1] - offs[i]);
            std::string_view l0 = s.substr(0, s.find('/'));
            if (std::hash<std::string_view>()(l0) % static_cast<size_t>(nthreads) != static_cast<size_t>(p)) continue;
            int cl = snprintf(cid, sizeof cid, "c%u", vals[i]);
            if (parts[p].add(s, std::string_view(cid, cl), vals[i], vals[i])) ok[p]++;
        }
    });
    for (auto& t : th) t.join();
    int64_t total = 0;
    for (int p = 0; p < nthreads; ++p) {
        total += ok[p];
        for (auto& kv : parts[p].topics.branches) r->topics.branches.emplace(kv.first, std::move(kv.second));
        r->relations.merge(parts[p].relations);
        r->topics_count += parts[p].topics_count;
        r->relations_count += parts[p].relations_count;
    }
    return total;
}
// Router::remove then Router::add of the same subscription (the pairs orc_router_bulk_add created), one thread — the
// reference's write-lock path (router.rs:417-479).  Returns seconds for n pairs (2n operations).
double orc_router_churn(void* rp, const char* blob, const uint32_t* offs, const uint32_t* vals, uint64_t n) {
    Router* r = static_cast<Router*>(rp);
    char cid[32];
    auto t0 = std::chrono::steady_clock::now();
    for (uint64_t i = 0; i < n; ++i) {
        std::string_view f(blob + offs[i], offs[i + 1] - offs[i]);
        int cl = snprintf(cid, sizeof cid, "c%u", vals[i]);
        r->remove(f, std::string_view(cid, cl), vals[i]);
        r->add(f, std::string_view(cid, cl), vals[i], vals[i]);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int64_t orc_router_match(void* r, const char* s, uint32_t len, uint32_t* out, uint64_t cap, orc_counters* ctr) {
    std::vector<uint32_t> v; Counters c;
    if (!static_cast<Router*>(r)->matches(std::string_view(s, len), v, c)) return -1;
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    put(ctr, c);
    return static_cast<int64_t>(v.size());
}
// Whole DefaultRouter::_matches restatement per topic, `nthreads` concurrent readers. Returns seconds.
double orc_router_match_batch(void* rp, const char* blob, const uint32_t* offs, uint64_t n, int nthreads,
                              int64_t* counts, uint64_t* total_ids, orc_counters* ctr) {
    Router* r = static_cast<Router*>(rp);
    std::vector<Counters> cs(std::max(1, nthreads));
    std::vector<uint64_t> tot(std::max(1, nthreads), 0);
    auto t0 = std::chrono::steady_clock::now();
    parallel_for(n, nthreads, [&](int tid, uint64_t b, uint64_t e) {
        std::vector<uint32_t> v; Counters c;
        for (uint64_t i = b; i < e; ++i) {
            v.clear();
            bool ok = r->matches(std::string_view(blob + offs[i], offs[i + 1] - offs[i]), v, c);
            if (counts) counts[i] = ok ? static_cast<int64_t>(v.size()) : -1;
            tot[tid] += v.size();
        }
        cs[tid].add(c);
    });
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    Counters t; uint64_t s = 0;
    for (size_t i = 0; i < cs.size(); ++i) { t.add(cs[i]); s += tot[i]; }
    put(ctr, t);
    if (total_ids) *total_ids = s;
    return dt;
}

// ---- RetainTree<i64> ---------------------------------------------------------------
void* orc_retain_new() { return new RTree(); }
void orc_retain_free(void* t) { delete static_cast<RTree*>(t); }
int32_t orc_retain_insert(void* tp, const char* s, uint32_t len, int64_t v) {
    Topic t;
    if (!parse_topic(std::string_view(s, len), t)) return -1;
    static_cast<RTree*>(tp)->insert(t, 0, v);
    return 1;
}
// 1 removed (old written), 0 nothing there, -1 invalid
int32_t orc_retain_remove(void* tp, const char* s, uint32_t len, int64_t* old) {
    Topic t;
    if (!parse_topic(std::string_view(s, len), t)) return -1;
    return static_cast<RTree*>(tp)->remove(t, 0, old) ? 1 : 0;
}
uint64_t orc_retain_values_size(void* tp) { return static_cast<RTree*>(tp)->values_size(); }
uint64_t orc_retain_nodes_size(void* tp) { return static_cast<RTree*>(tp)->nodes_size(); }
int64_t orc_retain_bulk_insert(void* tp, const char* blob, const uint32_t* offs, const uint32_t* vals, uint64_t n) {
    RTree* r = static_cast<RTree*>(tp); Topic t; int64_t ok = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (!parse_topic(std::string_view(blob + offs[i], offs[i + 1] - offs[i]), t)) continue;
        r->insert(t, 0, static_cast<int64_t>(vals[i])); ok++;
    }
    return ok;
}
int64_t orc_retain_match(void* tp, const char* s, uint32_t len, int64_t* out, uint64_t cap, orc_counters* ctr) {
    Topic t;
    if (!parse_topic(std::string_view(s, len), t)) return -1;
    Counters c; c.L = t.size(); c.B = len;
    std::vector<int64_t> v;
    static_cast<RTree*>(tp)->matches(t, 0, true, c, v);
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    put(ctr, c);
    return static_cast<int64_t>(v.size());
}
double orc_retain_match_batch(void* tp, const char* blob, const uint32_t* offs, uint64_t n, int nthreads,
                              int64_t* counts, const uint64_t* out_offs, uint32_t* out_ids, orc_counters* ctr) {
    RTree* tree = static_cast<RTree*>(tp);
    std::vector<Counters> cs(std::max(1, nthreads));
    auto t0 = std::chrono::steady_clock::now();
    parallel_for(n, nthreads, [&](int tid, uint64_t b, uint64_t e) {
        Topic t; Counters c; std::vector<int64_t> v;
        for (uint64_t i = b; i < e; ++i) {
            std::string_view s(blob + offs[i], offs[i + 1] - offs[i]);
            if (!parse_topic(s, t)) { if (counts) counts[i] = -1; continue; }
            c.L += t.size(); c.B += s.size();
            v.clear();
            tree->matches(t, 0, true, c, v);
            if (counts) counts[i] = static_cast<int64_t>(v.size());
            if (out_ids) for (size_t k = 0; k < v.size(); ++k) out_ids[out_offs[i] + k] = static_cast<uint32_t>(v[k]);
        }
        cs[tid].add(c);
    });
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    Counters tot; for (auto& c : cs) tot.add(c);
    put(ctr, tot);
    return dt;
}

// threads the batch functions should use: the CPUs this process is allowed on (not the machine's total)
int32_t orc_hardware_threads() {
    const size_t n = allowed_cpus().size();
    return static_cast<int32_t>(n ? n : std::thread::hardware_concurrency());
}

// Memory policy of the calling thread (inherited by the threads it creates): interleave new pages over all online
// NUMA nodes — a tree built by one socket's threads and read by both otherwise makes the multi-threaded baseline
// depend on where its pages happened to land.  Returns 0 on success, -1 if the kernel refused (harmless).
int32_t orc_numa_interleave(int32_t on) {
    unsigned long mask[16] = {0};
    unsigned long maxnode = 0;
    if (on) {
        FILE* f = std::fopen("/sys/devices/system/node/online", "r");
        if (!f) return -1;
        char buf[256] = {0};
        if (!std::fgets(buf, sizeof(buf), f)) { std::fclose(f); return -1; }
        std::fclose(f);
        for (char* p = buf; *p;) {                                    // "0-1,4" style list
            char* e;
            long a = std::strtol(p, &e, 10), b = a;
            if (e == p) break;
            if (*e == '-') { p = e + 1; b = std::strtol(p, &e, 10); }
            for (long k = a; k <= b && k < 1024; ++k) { mask[k / (8 * sizeof(long))] |= 1ul << (k % (8 * sizeof(long))); maxnode = std::max<unsigned long>(maxnode, k + 1); }
            p = (*e == ',') ? e + 1 : e;
            if (*e != ',') break;
        }
        if (maxnode < 2) return 0;                                    // one node: nothing to interleave
    }
    return syscall(SYS_set_mempolicy, on ? 3 /* MPOL_INTERLEAVE */ : 0 /* MPOL_DEFAULT */, on ? mask : nullptr, on ? maxnode + 1 : 0) == 0 ? 0 : -1;
}

}  // extern "C"
