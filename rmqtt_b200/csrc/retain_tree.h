// Host mirror + device layout of the retained-message tree (rmqtt/src/retain.rs:202-257).
//
// The tree is keyed by CONCRETE topics (one optional value per node); queries are topic FILTERS:
// '+' = every child, '#' = the whole subtree (retain.rs:298-367).  Device layout (rebuilt by `flatten`
// whenever the tree changed, shipped by the engine's flush):
//
//   rnodes : nodes in DFS PRE-ORDER, so the subtree of node n is the index range [n, sub_end(n)) and the
//            values below a node are ONE contiguous range of `rvals` -> a '#' match is a range copy
//   rkids  : per node, its children as a contiguous block of 32-byte entries {token, child, child's record}
//            -> a '+' match is a coalesced scan of one block and yields complete work items
//   redges : open-addressing hash (parent, token) -> child + the child's record for the exact-level steps
//   rvals  : the values of all valued nodes, in pre-order
#pragma once
#include <vector>

#include "host_trie.h"
#include "layout.h"

namespace gm {

constexpr u32 RVAL_NONE = 0xFFFFFFFFu;
constexpr u32 RF_LIT_PLUS = 1u;       // node has a child whose level is literally "+"  (shadows '+' expansion, retain.rs:313)
constexpr u32 RF_LIT_HASH = 2u;       // node has a child whose level is literally "#"
constexpr u32 RF_SUB_LIT_HASH = 4u;   // some node in this subtree (self included) has a literal "#" child:
                                      // the '#' recursion is shadowed somewhere below -> no range shortcut

struct alignas(32) RNode {
    u32 first_kid, nkids;
    u32 val;              // RVAL_NONE = no value
    u32 val_lo, val_hi;   // rvals[val_lo .. val_hi) = values of the subtree [n, sub_end) (own value first)
    u32 flags;
    u32 sub_end;
    u32 pad;
};
// A node's record as the walk needs it: {first_kid, nk_flags, val, val_lo, val_hi} where nk_flags = #kids (28 bits) |
// flags << 28 (RF_* and 8 = has value).  It travels WITH the reference to the node — inside the parent's child
// block entry and inside the (parent, token) hash slot — so visiting a node costs no extra memory access.
constexpr u32 RNK_MASK = 0x0FFFFFFFu;
struct alignas(32) RKid { u32 token, child, first_kid, nk_flags, val, val_lo, val_hi, pad; };
struct alignas(32) REdge { u32 parent, token, child, first_kid, nk_flags, val, val_lo, val_hi; };   // child == 0: empty (root is node 0)

struct RetainView {
    const RKid* kids;
    const REdge* edges;
    const u32* vals;
    u32 edge_mask;
    u32 root_first_kid, root_nk_flags;
    u32 root_plain_kids;     // root children whose level does not start with '$' (ordered first)
    u32 root_plain_val_hi;   // rvals[0 .. root_plain_val_hi) = values below those children
    u32 max_depth;
};

GM_HD u32 redge_hash(u32 parent, u32 token) { return fmix32(parent * 0x85EBCA77u + (token ^ 0x2545F491u) * 0x9E3779B1u); }

class RetainTreeHost {
  public:
    explicit RetainTreeHost(HostTrie* dict) : dict_(dict) { nodes_.emplace_back(); }

    // RetainTree::insert (retain.rs:221-234): value.replace(v).  Returns ParseStatus; *had_old / *old as remove.
    int set(const char* topic, u32 len, u32 value, bool* had_old, u32* old);
    // RetainTree::remove (retain.rs:237-257) with bottom-up pruning.
    int remove(const char* topic, u32 len, bool* had_old, u32* old);

    // Dictionary compaction (HostTrie::compact): the level tokens this tree holds, and their replacement.
    std::vector<u32> used_tokens() const;
    void remap_tokens(const std::vector<u32>& remap);

    u64 values_size() const { return n_values_; }     // retain.rs:385-392
    u64 nodes_size() const { return n_nodes_; }       // retain.rs:395-398

    bool dirty = true;
    void flatten();                                     // host tree -> device-layout arrays below
    std::vector<RNode> rnodes;
    std::vector<RKid> rkids;
    std::vector<REdge> redges;
    std::vector<u32> rvals;
    u32 root_plain_kids = 0, root_plain_val_hi = 0, max_depth = 0;

  private:
    struct HN {
        std::vector<std::pair<u32, u32>> kids;   // (token, node), sorted by token
        u32 val = RVAL_NONE;
        bool has_val = false;
        u32 parent = 0, token = 0;
    };
    int parse(const char* s, u32 len, bool create, std::vector<u32>& toks);
    u32 child_of(u32 node, u32 token) const;
    HostTrie* dict_;
    std::vector<HN> nodes_;
    std::vector<u32> free_;
    std::vector<u32> toks_;
    u64 n_values_ = 0, n_nodes_ = 0;
};

}  // namespace gm
