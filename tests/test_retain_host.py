"""CPU tier: retained-tree host mirror + pre-order flattening (rmqtt_b200/csrc/retain_tree.cpp) against the
oracle's RetainTree restatement, through a pure-Python model of the retained-lookup kernels."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from rmqtt_b200.engine import Engine, GpuMqttError, pack

from _gen import rand_filter, rand_topic
from _retainwalk import RetainTables


def _tables(eng):
    eng.flush()
    t = eng.debug_tables()
    R = RetainTables(t)
    # root_plain_* are part of the view; recompute them the way flatten() defines them
    nodes, kids, vals = t["rnodes"], t["rkids"], t["rvals"]
    fk, nk = int(nodes[0, 0]), int(nodes[0, 1])
    plain = 0
    for j in range(nk):
        child = int(kids[fk + j, 1])
        # placeholder; _plain_split() sets the real values
        plain += 1
    R.set_root_plain(plain, int(nodes[0, 4]))
    return R, t


def _plain_split(eng, R, t, dollar_roots):
    """root_plain_kids / root_plain_val_hi given which level-0 strings start with '$'."""
    nodes, kids = t["rnodes"], t["rkids"]
    fk, nk = int(nodes[0, 0]), int(nodes[0, 1])
    toks = {R.T.token(s.encode()) for s in dollar_roots}
    plain = sum(1 for j in range(nk) if int(kids[fk + j, 0]) not in toks)
    # children are ordered plain-first
    assert all(int(kids[fk + j, 0]) not in toks for j in range(plain))
    hi = int(nodes[int(kids[fk + plain, 1]), 3]) if plain < nk else int(nodes[0, 4])
    R.set_root_plain(plain, hi)


def test_golden_retain_rs_451_475(golden):
    g = golden["retain_A3"]
    eng, tree = Engine(host_only=True), orc.RetainTree()
    for topic, v in g["inserts"] + g["more_inserts"]:
        eng.retain_set(topic, v & 0xFFFFFFFF)
        tree.insert(topic, v & 0xFFFFFFFF)
    R, t = _tables(eng)
    _plain_split(eng, R, t, [])
    for f, _ in g["matches"] + g["more_matches"] + golden["derived_A5"]["retain_on_A3"]:
        assert R.match(f.encode()) == tree.matches(f), f
    st = eng.stats()
    assert st["retained_values"] == tree.values_size() == 12 and st["retained_nodes"] == tree.nodes_size()


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_random_differential(seed):
    rng = random.Random(seed)
    eng, tree = Engine(host_only=True), orc.RetainTree()
    topics = []
    for step in range(700):
        if topics and rng.random() < 0.25:
            t = rng.choice(topics)
            try:
                got = eng.retain_remove(t)
            except Exception:
                continue
            assert got == tree.remove(t)
        else:
            t = rand_topic(rng, max_depth=5) if rng.random() < 0.9 else rand_filter(rng, 5)   # literal '+' / '#' levels too
            v = rng.randint(0, 2**32 - 2)
            try:
                old = eng.retain_set(t, v)
            except Exception:
                with pytest.raises(ValueError):
                    tree.insert(t, v)
                continue
            prev = tree.remove(t)
            tree.insert(t, v)
            assert old == prev
            topics.append(t)
        if step % 175 == 174:
            st = eng.stats()
            assert st["retained_values"] == tree.values_size() and st["retained_nodes"] == tree.nodes_size()
            R, t = _tables(eng)
            _plain_split(eng, R, t, ["$SYS", "$q", "$share"])
            for _ in range(200):
                f = rand_filter(rng, 6)
                want = tree.matches(f)
                got = R.match(f.encode())
                assert got == want, f


@pytest.mark.parametrize("threads", [2, 5, 64])
def test_bulk_load_builds_tree_and_image_level_by_level(monkeypatch, threads):
    """gm_retain_bulk_load into an empty tree builds the host tree and its device image together on all host threads
    (retain_tree.cpp set_batch_build: sorted (parent, token) keys per level, pre-order numbers from subtree sizes).  The image
    must be the one set() x n + flatten() produce — array for array — and the tree must keep working incrementally."""
    rng = random.Random(300 + threads)
    topics = [rand_topic(rng, max_depth=5) if rng.random() < 0.9 else rand_filter(rng, 5) for _ in range(5000)]
    topics += ["$SYS/broker/up", "$SYS/broker/load", "$q/x", "", "/", "a//b", "dup/t", "dup/t", "lit/#", "lit/+/x", "bad/#/x", "a/$b"]
    rng.shuffle(topics)
    vals = [rng.randint(0, 2**32 - 2) for _ in topics]
    one, tree = Engine(host_only=True), orc.RetainTree()
    n_ok = 0
    for t, v in zip(topics, vals):
        try:
            one.retain_set(t, v)
        except GpuMqttError:
            continue
        tree.remove(t); tree.insert(t, v)
        n_ok += 1
    ta = one.debug_tables()                               # flushes: flatten()
    monkeypatch.setenv("GM_HOST_PAR_MIN", "1")
    monkeypatch.setenv("GM_HOST_THREADS", str(threads))
    bulk = Engine(host_only=True)
    blob, offs = pack(topics)
    assert bulk.retain_bulk_load(blob, offs, np.asarray(vals, dtype=np.uint32)) == n_ok
    tb = bulk.debug_tables()
    assert int(tb["rstats"][0]) == 1 and int(tb["rstats"][5]) == 1       # built once, image valid: the flush had nothing to flatten
    for name in ("rnodes", "rkids", "rvals"):
        assert np.array_equal(ta[name], tb[name]), name
    ea, eb = ta["redges"], tb["redges"]                   # the hash table: the same entries (colliding ones may sit in other slots)
    ea, eb = ea[ea[:, 2] != 0], eb[eb[:, 2] != 0]
    assert np.array_equal(ea[np.lexsort(ea.T[::-1])], eb[np.lexsort(eb.T[::-1])])
    sa, sb = one.stats(), bulk.stats()
    assert sa["retained_values"] == sb["retained_values"] == tree.values_size() and sa["retained_nodes"] == sb["retained_nodes"] == tree.nodes_size()
    R, t = _tables(bulk)
    _plain_split(bulk, R, t, ["$SYS", "$q", "$share"])
    for _ in range(300):
        f = rand_filter(rng, 6)
        assert R.match(f.encode()) == tree.matches(f), f
    # incremental edits on the bulk-built tree (in-place patches, then a re-pack)
    for t_ in topics[:300]:
        try:
            got = bulk.retain_remove(t_)
        except GpuMqttError:
            continue
        assert got == tree.remove(t_)
    for k in range(200):
        t_, v = rand_topic(rng, max_depth=5), rng.randint(0, 2**32 - 2)
        try:
            old = bulk.retain_set(t_, v)
        except GpuMqttError:
            continue
        prev = tree.remove(t_); tree.insert(t_, v)
        assert old == prev
    R, t = _tables(bulk)
    _plain_split(bulk, R, t, ["$SYS", "$q", "$share"])
    assert bulk.stats()["retained_nodes"] == tree.nodes_size()
    for _ in range(300):
        f = rand_filter(rng, 6)
        assert R.match(f.encode()) == tree.matches(f), f


def test_republish_with_the_same_handle_does_not_dirty_the_device_copy():
    eng = Engine(host_only=True)
    eng.retain_set("a/b", 7)
    eng.flush()
    assert eng.stats()["pending"] in (0, 1)        # root record of the subscription trie may still be pending in host-only mode
    eng.retain_set("a/b", 7)                       # same topic, same handle (new payload lives on the host)
    t1 = eng.debug_tables()["rvals"].tolist()
    eng.retain_set("a/b", 8)
    t2 = eng.debug_tables()["rvals"].tolist()
    assert t1 == [7] and t2 == [8]


def test_compact_relabels_the_retained_tree():
    """gm_compact re-assigns every dictionary token; the retained tree shares the dictionary and must be re-labelled:
    both trees keep matching exactly like the oracle afterwards, and keep working incrementally."""
    import random
    from _gen import rand_filter, rand_topic
    from _tablewalk import Tables
    rng = random.Random(5)
    eng, sub, ret = Engine(host_only=True), orc.TopicTree(), orc.RetainTree()
    filters = []
    for i in range(800):
        f = rand_filter(rng)
        try:
            eng.add(f, i)
        except GpuMqttError:
            continue
        sub.insert(f, i)
        filters.append((f, i))
    topics = []
    for i in range(600):
        t = rand_topic(rng)
        try:
            eng.retain_set(t, i)
        except GpuMqttError:
            continue
        ret.insert(t, i)
        topics.append(t)
    for f, i in filters[:600]:            # most level strings of the subscription side become garbage
        eng.remove(f, i); sub.remove(f, i)
    for t in topics[::3]:
        eng.retain_remove(t); ret.remove(t)
    before = eng.stats()
    eng.compact()
    after = eng.stats()
    assert after["dict_entries"] <= before["dict_entries"]
    assert after["retained_values"] == ret.values_size() and after["retained_nodes"] == ret.nodes_size()
    eng.retain_set("fresh/topic/after", 424242); ret.insert("fresh/topic/after", 424242)
    eng.add("fresh/+/after", 7); sub.insert("fresh/+/after", 7)
    R, t = _tables(eng)
    _plain_split(eng, R, t, ["$SYS", "$q", "$share"])
    T = Tables(t)
    for _ in range(300):
        q = rand_filter(rng)
        want = ret.matches(q)
        if want is not None:
            assert R.match(q.encode()) == want, q
        tp = rand_topic(rng)
        assert T.match(tp.encode())[0] == sub.matches(tp), tp
    assert R.match(b"fresh/+/after") == ret.matches("fresh/+/after") == [424242]
    got = T.match(b"fresh/topic/after")[0]
    assert got == sub.matches("fresh/topic/after") and 7 in got


def test_incremental_updates_edit_the_image_in_place():
    """After the first flatten, set / remove must keep the device image current with in-place edits (no re-flatten)
    for plain topics — new leaves, new deep paths, value replacement, removal with pruning, revival of pruned paths,
    root children incl. `$` roots — and the image must match the oracle after every round; a literal `+` level or
    too much garbage falls back to a full re-flatten."""
    rng = random.Random(77)
    eng, tree = Engine(host_only=True), orc.RetainTree()
    lv = ["a", "b", "c", "d", "", "x" * 30]
    def topic():
        n = rng.randint(1, 5)
        t = [rng.choice(lv) for _ in range(n)]
        if rng.random() < 0.1:
            t[0] = rng.choice(["$SYS", "$q"])
        return "/".join(t)
    live = []
    for i in range(300):
        t = topic(); eng.retain_set(t, i); tree.remove(t); tree.insert(t, i); live.append(t)
    R, tb = _tables(eng)
    base = tb["rstats"].tolist()
    assert base[0] == 1 and base[5] == 1
    v = 1000
    for rnd in range(25):
        for _ in range(rng.randint(1, 12)):
            r = rng.random()
            if r < 0.35 and live:
                t = rng.choice(live)
                assert eng.retain_remove(t) == tree.remove(t)
            elif r < 0.55 and live:
                t = rng.choice(live); v += 1                     # replace (or revive) an existing path
                old = eng.retain_set(t, v); assert old == tree.remove(t); tree.insert(t, v)
            else:
                t = topic(); v += 1
                old = eng.retain_set(t, v); assert old == tree.remove(t); tree.insert(t, v); live.append(t)
        st = eng.stats()
        assert st["retained_values"] == tree.values_size() and st["retained_nodes"] == tree.nodes_size()
        R, tb = _tables(eng)
        _plain_split(eng, R, tb, ["$SYS", "$q"])
        for _ in range(120):
            f = rand_filter(rng, 5) if rng.random() < 0.5 else "/".join(rng.choice(lv + ["+", "+", "#"]) for _ in range(rng.randint(1, 5)))
            want = tree.matches(f)
            if want is not None:
                assert R.match(f.encode()) == want, (rnd, f)
    s = tb["rstats"].tolist()
    assert s[0] <= base[0] + 1 and s[1] > 0, s                   # at most one garbage-triggered re-pack; edits happened in place
    # a literal '+' level changes the shadowing flags: rebuilt at the next flush, still exact
    eng.retain_set("a/+/c", 5); tree.insert("a/+/c", 5)
    R, tb = _tables(eng)
    _plain_split(eng, R, tb, ["$SYS", "$q"])
    assert tb["rstats"].tolist()[0] == s[0] + 1
    for f in ("a/+/c", "a/#", "+/+/c", "#", "a/b/#"):
        assert R.match(f.encode()) == tree.matches(f), f


def test_batch_removal_is_the_expiry_sweep():
    """gm_retain_remove_batch == n calls of gm_retain_remove (remove_expired_messages, retain.rs:118-128)."""
    eng, tree = Engine(host_only=True), orc.RetainTree()
    topics = [f"a/{i % 7}/b{i}" for i in range(300)]
    for i, t in enumerate(topics):
        eng.retain_set(t, i); tree.insert(t, i)
    victims = topics[::3] + ["a/none/here", "bad/#/x", "a/0"]
    blob, offs = pack(victims)
    old, removed = eng.retain_remove_batch(blob, offs)
    want = []
    for t in victims:
        try:
            want.append(tree.remove(t))
        except ValueError:
            want.append(None)
    assert removed == sum(w is not None for w in want)
    assert [None if o == 0xFFFFFFFF else int(o) for o in old] == want
    st = eng.stats()
    assert st["retained_values"] == tree.values_size() and st["retained_nodes"] == tree.nodes_size()
