// Host mirror of the device-resident subscription trie.
//
// Mutations (Router::add / Router::remove -> TopicTree::insert / remove, rmqtt/src/trie.rs:99-135,
// rmqtt/src/router.rs:417-479) are applied here first: the host owns an exact copy of every device
// table in the device layout (layout.h), remembers which 32-byte slots changed, and `flush` ships only
// those slots to HBM on a side stream (engine.cu).  Nothing in this file is on the match path.
#pragma once
#include <cstddef>
#include <string>
#include <unordered_map>
#include <vector>

#include "huge_alloc.h"
#include "layout.h"

namespace gm {

struct HNode {
    u32 parent = 0;
    u32 token = 0;
    u32 edge_slot = 0;       // slot in `edges` holding this node's record (non-root)
    u32 plus_child = 0;      // node id of the '+' child (0 = none)
    u32 hash_child = 0;      // node id of the '#' child (0 = none)
    u32 nvals = 0;           // size of the value set (BTreeSet<V>)
    u32 v0 = 0;              // the value when nvals == 1
    u32 ref = 0;             // value-set reference currently published to the device ...
    u32 cnt16 = 0;           // ... and its 16-bit count (layout.h)
    u32 live_children = 0;
    u32 lit_children = 0;    // literal child edges ever created
    u32 mask = 0;
    u8 alive = 0;            // reference semantics: pruned nodes (trie.rs:126-128) are "not alive"
    u8 dirty = 0;
    u8 wide = 0;             // more than WIDE_FANOUT literal children: its child edges live in the child filter
    u8 wtag = 0;             // window tag: which window of `edges` holds this node's child edges (layout.h)
    uint16_t depth = 0;           // levels from the root (root = 0)
};

enum ParseStatus { PARSE_OK = 0, PARSE_INVALID = 1, PARSE_TOO_DEEP = 2 };

class HostTrie {
  public:
    explicit HostTrie(u32 max_levels);

    // TopicTree::insert / remove.  Return ParseStatus; *changed mirrors the reference's bool.
    // `tree`: which TopicTree of this engine (0 = the subscription trie).  The reference keeps further, small TopicTree<V>s
    // that are asked on every PUBLISH / SUBSCRIBE too — ACL rule trees (rmqtt-plugins/rmqtt-acl/src/config.rs:291-326),
    // topic-rewrite rules, bridge routing tables: they live in the same device tables as extra ROOTS and are matched in
    // the same batch (every row of a batch names its tree).
    int insert(const char* filter, u32 len, u32 value, bool* changed, u32 tree = 0);
    int remove(const char* filter, u32 len, u32 value, bool* changed, u32 tree = 0);
    static constexpr u32 MAX_TREES = 4096;
    std::vector<u32> tree_slots{0xFFFFFFFFu};   // device mirror: tree id -> edge slot holding its root record (0xFFFFFFFF = no such tree); [0] unused
    bool trees_dirty = false;
    // n inserts (invalid filters are skipped); returns how many changed the set.  Same result as one-by-one insert().
    u64 insert_batch(const char* blob, const u32* offsets, const u32* values, u64 n);
    // Levels of a whole batch as tokens (new level strings are interned in first-occurrence order, as parse() would do one by
    // one): toks[lvl_off[i] .. +depth[i]) for filter i; depth 0 = invalid or too deep.
    struct TokenizedBatch { BigVec<u64> lvl_off; BigVec<u32> toks; BigVec<uint16_t> depth; double t_classify = 0, t_dictionary = 0, t_resolve = 0; };
    void tokenize_batch(const char* blob, const u32* offsets, u64 n, unsigned threads, TokenizedBatch& out);
    u64 insert_batch_parallel(const char* blob, const u32* offsets, const u32* values, u64 n, unsigned threads);   // same content, all host threads
    void reserve(u64 n_filters);

    // Resolve dirty nodes into slot patches (called by flush).  false: more than 2^32 live value words (references would wrap).
    bool sync();
    u64 values_epoch = 0;     // bumped whenever `values` / `ranges` were rebuilt from scratch (the device copy must be re-shipped whole)
    // Rebuild every table from the live content: drops pruned nodes, dead dictionary use and value garbage.
    // Tokens are re-assigned; the caller must treat all device tables as new (everything is marked dirty).
    // `keep` lists old tokens that other users of the dictionary (the retained tree) still hold: their strings are
    // re-interned and `remap` (if given) receives old token -> new token for them (0 for every other token).
    void compact(const std::vector<u32>* keep = nullptr, std::vector<u32>* remap = nullptr);

    // reference-visible statistics
    u64 values_size() const { return values_size_; }   // trie.rs:148-151
    u64 nodes_size() const { return live_nodes_; }     // trie.rs:154-157

    // ---- device mirror -------------------------------------------------------------------------
    ZeroTable<EdgeSlot> edges;
    StableVec<Range> ranges;         // [0] reserved.  Both arrays never move in memory (gm_values_view hands out their base)
    StableVec<u32> values;
    ZeroTable<DictSlot> dict;
    std::vector<u8> pool;
    std::vector<u32> cfilter;        // child filter of wide nodes (layout.h)
    bool cfilter_dirty = true;
    u32 root_plus = 0, root_hash_ref = 0, root_hash_cnt = 0, root_mask = 0;
    u32 max_depth = 0;
    // window geometry of `edges` (layout.h): windows = 1 << nwin_log2, each edges.size() >> nwin_log2 slots
    u32 nwin_log2 = 0;
    u32 table_log2 = 10;      // log2(edges.size()), kept by rehash_edges
    u32 win_shift() const { return table_log2 - nwin_log2; }
    u32 win_mask() const { return (1u << win_shift()) - 1u; }
    u32 nwin_mask() const { return (1u << nwin_log2) - 1u; }

    // ---- dirty tracking (consumed and cleared by the engine's flush) ---------------------------
    std::vector<u32> dirty_edges, dirty_dict;       // (`ranges` / `values` / `pool` are append-only between compactions: shipped as tails)
    bool full_edges = true, full_dict = true;   // table re-hashed / never uploaded: ship whole table
    bool any_dirty() const { return !dirty_nodes_.empty() || !dirty_edges.empty() ||
                                    !dirty_dict.empty() || full_edges || full_dict || root_dirty || cfilter_dirty || trees_dirty ||
                                    cfilter_rebuild_; }
    bool root_dirty = true;
    u64 garbage_values = 0;

    u64 edge_count() const { return edge_count_; }
    u64 plus_count() const { return plus_count_; }
    u64 dict_count() const { return dict_count_; }
    u64 node_count() const { return nodes_.size(); }

    // host-side tokeniser (tests compare the device tokeniser against it)
    u32 lookup_token(const char* s, u32 len) const;
    static u32 level0_hash(const char* s, u32 len);
    // Topic::from_str -> tokens (shared with the retained tree, which interns into the same dictionary)
    int parse(const char* f, u32 len, bool intern_new, std::vector<u32>& toks);
    bool token_is_dollar(u32 tok) const { return tok < tok_dollar_.size() && tok_dollar_[tok]; }

  private:
    u32 intern(const char* s, u32 len, bool create);
    u32 tree_root(u32 tree, bool create);               // node id of an extra tree's root (0 = none)
    std::vector<u32> tree_nodes_{0u};                    // tree id -> node id
    std::unordered_map<u32, u32> tree_of_token_;         // reserved level token of a tree root -> tree id
    u32 find_edge(u32 parent, u32 token, u32 wtag) const;   // returns slot index or ~0u; wtag = window tag of `parent`
    u32 add_edge(u32 parent, u32 token);                // creates the child node, returns its id
    bool add_value(u32 node, u32 value);                // BTreeSet::insert on the node's value set
    void grow_edges() { rehash_edges(edges.size() * 2); }
    void rehash_edges(size_t new_size);                 // re-places every edge (new size and / or new window count)
    void make_room(u32 wtag);                           // before an insertion into the window of `wtag`
    u32 pick_tag();                                     // least-loaded window tag for a new depth-2 subtree
    void grow_dict();
    void mark(u32 node);
    void mark_vals(u32 node);
    void compact_values();
    void make_ref(u32 node);
    void write_record(u32 node);
    void cfilter_insert(u32 parent, u32 token);
    void cfilter_rebuild();
    bool cfilter_rebuild_ = false;
    u64 cfilter_keys_ = 0;

    u32 max_levels_;
    std::vector<HNode, HugeAlloc<HNode>> nodes_;
    std::unordered_map<u32, std::vector<u32>> multi_;   // node -> sorted values when nvals > 1
    std::vector<u32> dirty_nodes_;
    std::vector<u32> scratch_toks_;
    std::vector<u8> tok_dollar_;      // token -> level string starts with '$' (Level::Metadata)
    u64 edge_count_ = 0, dict_count_ = 0, plus_count_ = 0, anchors_ = 0;
    u32 nwin_cap_log2_ = WIN_MAX_LOG2;
    u32 win_min_log2_ = WIN_MIN_SLOTS_LOG2;   // raised when one subtree outgrows its window
    std::vector<u64> tag_count_;         // edges per window tag (WTAG_COUNT entries)
    std::vector<u64> tag_anchors_;       // depth-2 subtrees assigned per window tag
    std::vector<u64> win_count_;         // edges per effective window
    u32 next_token_ = TOK_FIRST;
    u64 values_size_ = 0, live_nodes_ = 0;
};

}  // namespace gm
