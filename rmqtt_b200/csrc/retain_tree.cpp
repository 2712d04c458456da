// Host mirror of the retained-message tree — see retain_tree.h.  Semantics follow
// rmqtt/src/retain.rs:221-257 (insert = value.replace, remove with bottom-up pruning).
#include "retain_tree.h"

#include <algorithm>

namespace gm {

int RetainTreeHost::parse(const char* s, u32 len, bool create, std::vector<u32>& toks) {
    return dict_->parse(s, len, create, toks);     // Topic::from_str (topic.rs:348-363)
}

u32 RetainTreeHost::child_of(u32 node, u32 token) const {
    const auto& k = nodes_[node].kids;
    auto it = std::lower_bound(k.begin(), k.end(), std::make_pair(token, 0u));
    return (it != k.end() && it->first == token) ? it->second : 0u;
}

int RetainTreeHost::set(const char* topic, u32 len, u32 value, bool* had_old, u32* old) {
    if (had_old) *had_old = false;
    int st = parse(topic, len, true, toks_);
    if (st != PARSE_OK) return st;
    u32 node = 0;
    for (u32 tok : toks_) {
        u32 c = child_of(node, tok);
        if (!c) {
            if (!free_.empty()) { c = free_.back(); free_.pop_back(); nodes_[c] = HN{}; }
            else { c = static_cast<u32>(nodes_.size()); nodes_.emplace_back(); }
            nodes_[c].parent = node; nodes_[c].token = tok;
            auto& k = nodes_[node].kids;
            k.insert(std::lower_bound(k.begin(), k.end(), std::make_pair(tok, 0u)), std::make_pair(tok, c));
            n_nodes_++;
        }
        node = c;
    }
    HN& n = nodes_[node];
    // Re-publishing a retained message on a topic that keeps its handle changes nothing on the device: the
    // device copy is only rebuilt when the set of retained topics (or a handle) actually changes.
    const bool same = n.has_val && n.val == value;
    if (n.has_val) { if (had_old) *had_old = true; if (old) *old = n.val; }
    else n_values_++;
    n.has_val = true; n.val = value;
    if (!same) dirty = true;
    return PARSE_OK;
}

int RetainTreeHost::remove(const char* topic, u32 len, bool* had_old, u32* old) {
    if (had_old) *had_old = false;
    int st = parse(topic, len, false, toks_);
    if (st != PARSE_OK) return st;
    u32 node = 0;
    for (u32 tok : toks_) {
        if (tok == TOK_UNKNOWN) return PARSE_OK;
        node = child_of(node, tok);
        if (!node) return PARSE_OK;
    }
    HN& n = nodes_[node];
    if (n.has_val) {
        if (had_old) *had_old = true;
        if (old) *old = n.val;
        n.has_val = false; n.val = RVAL_NONE;
        n_values_--;
        dirty = true;
    }
    // prune (retain.rs:247-249): value.is_none() && branches.is_empty()
    for (u32 x = node; x != 0;) {
        HN& c = nodes_[x];
        if (c.has_val || !c.kids.empty()) break;
        u32 p = c.parent;
        auto& k = nodes_[p].kids;
        k.erase(std::lower_bound(k.begin(), k.end(), std::make_pair(c.token, 0u)));
        free_.push_back(x);
        n_nodes_--;
        dirty = true;
        x = p;
    }
    return PARSE_OK;
}

std::vector<u32> RetainTreeHost::used_tokens() const {
    std::vector<u32> out;
    std::vector<u32> todo{0u};
    while (!todo.empty()) {
        const u32 n = todo.back(); todo.pop_back();
        for (const auto& kv : nodes_[n].kids) { if (kv.first >= TOK_FIRST) out.push_back(kv.first); todo.push_back(kv.second); }
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

void RetainTreeHost::remap_tokens(const std::vector<u32>& remap) {
    std::vector<u32> todo{0u};
    while (!todo.empty()) {
        const u32 n = todo.back(); todo.pop_back();
        for (auto& kv : nodes_[n].kids) {
            if (kv.first >= TOK_FIRST) kv.first = remap[kv.first];
            nodes_[kv.second].token = kv.first;
            todo.push_back(kv.second);
        }
        std::sort(nodes_[n].kids.begin(), nodes_[n].kids.end());
    }
    dirty = true;
}

void RetainTreeHost::flatten() {
    rnodes.clear(); rkids.clear(); rvals.clear();
    rnodes.reserve(n_nodes_ + 1); rkids.reserve(n_nodes_ + 1); rvals.reserve(n_values_);
    max_depth = 0;
    // iterative pre-order DFS; root children ordered plain-first, `$`-prefixed last (retain.rs:327-331, 345-349)
    struct Frame { u32 host, dev, next, depth; std::vector<std::pair<u32, u32>> order; };
    std::vector<Frame> stack;
    auto open = [&](u32 host, u32 depth) {
        const HN& h = nodes_[host];
        u32 dev = static_cast<u32>(rnodes.size());
        RNode r{};
        r.first_kid = static_cast<u32>(rkids.size());
        r.nkids = static_cast<u32>(h.kids.size());
        r.val = h.has_val ? h.val : RVAL_NONE;
        r.val_lo = static_cast<u32>(rvals.size());
        r.flags = h.has_val ? 8u : 0u;
        rnodes.push_back(r);
        if (h.has_val) rvals.push_back(h.val);
        Frame f{host, dev, 0, depth, h.kids};
        if (host == 0)
            std::stable_partition(f.order.begin(), f.order.end(), [&](const std::pair<u32, u32>& kv) { return !dict_->token_is_dollar(kv.first); });
        for (auto& kv : f.order) {
            if (kv.first == TOK_PLUS) rnodes[dev].flags |= RF_LIT_PLUS;
            if (kv.first == TOK_HASH) rnodes[dev].flags |= RF_LIT_HASH | RF_SUB_LIT_HASH;
        }
        rkids.resize(rkids.size() + f.order.size());
        max_depth = std::max(max_depth, depth);
        stack.push_back(std::move(f));
    };
    open(0, 0);
    root_plain_kids = 0;
    for (auto& kv : stack.back().order) if (!dict_->token_is_dollar(kv.first)) root_plain_kids++;
    root_plain_val_hi = 0;
    while (!stack.empty()) {
        Frame& f = stack.back();
        if (f.next < f.order.size()) {
            u32 j = f.next++;
            if (f.host == 0 && j == root_plain_kids) root_plain_val_hi = static_cast<u32>(rvals.size());
            u32 child_host = f.order[j].second;
            u32 child_dev = static_cast<u32>(rnodes.size());
            RKid& k = rkids[rnodes[f.dev].first_kid + j];
            k.token = f.order[j].first;
            k.child = child_dev;                    // the child's record is copied in below, once its subtree is complete
            open(child_host, f.depth + 1);          // invalidates `f`
        } else {
            RNode& r = rnodes[f.dev];
            r.sub_end = static_cast<u32>(rnodes.size());
            r.val_hi = static_cast<u32>(rvals.size());
            u32 flags = r.flags;
            bool root = f.host == 0;
            u32 parent_dev = 0;
            stack.pop_back();
            if (!stack.empty()) { parent_dev = stack.back().dev; if (flags & RF_SUB_LIT_HASH) rnodes[parent_dev].flags |= RF_SUB_LIT_HASH; }
            if (root && root_plain_kids == nodes_[0].kids.size()) root_plain_val_hi = static_cast<u32>(rvals.size());
        }
    }
    // child entries carry the complete record of the child (flags and value ranges are final only now)
    for (RKid& k : rkids) {
        const RNode& c = rnodes[k.child];
        k.first_kid = c.first_kid; k.nk_flags = c.nkids | (c.flags << 28); k.val = c.val; k.val_lo = c.val_lo; k.val_hi = c.val_hi; k.pad = 0;
    }
    // exact-step hash table, load <= 0.25
    size_t cap = 1024;
    while (cap < rkids.size() * 4) cap <<= 1;
    redges.assign(cap, REdge{0, 0, 0, 0, 0, 0, 0, 0});
    const u32 mask = static_cast<u32>(cap - 1);
    for (u32 n = 0; n < rnodes.size(); ++n) {
        const RNode& r = rnodes[n];
        for (u32 j = 0; j < r.nkids; ++j) {
            const RKid& k = rkids[r.first_kid + j];
            u32 i = redge_hash(n, k.token) & mask;
            while (redges[i].child != 0) i = (i + 1) & mask;
            redges[i] = REdge{n, k.token, k.child, k.first_kid, k.nk_flags, k.val, k.val_lo, k.val_hi};
        }
    }
    dirty = false;
}

}  // namespace gm
