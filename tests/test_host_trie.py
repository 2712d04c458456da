"""CPU tier: the host-side trie builder (rmqtt_b200/csrc/host_trie.cpp) and the device table layout,
checked against the oracle through a pure-Python model of the kernels' walk (tests/_tablewalk.py)."""
import random

import pytest

from oracle import oracle as orc
from rmqtt_b200.engine import Engine, GpuMqttError
from rmqtt_b200 import _native as N

from _gen import rand_filter, rand_topic
from _tablewalk import Tables


def _apply(eng, tree, f, v, op):
    try:
        got = eng.add(f, v) if op == "add" else eng.remove(f, v)
    except GpuMqttError as ex:
        assert ex.code == N.GM_ERR_INVALID_TOPIC
        with pytest.raises(ValueError):
            tree.insert(f, v) if op == "add" else tree.remove(f, v)
        return
    want = tree.insert(f, v) if op == "add" else tree.remove(f, v)
    assert got == want, (op, f, v)


def test_golden_through_tables(golden):
    for key in ("trie_A1",):
        g = golden[key]
        eng, tree = Engine(host_only=True), orc.TopicTree()
        for f, v in g["inserts"]:
            _apply(eng, tree, f, v, "add")
        T = Tables(eng.debug_tables())
        for topic, want in g["matches"]:
            assert T.match(topic.encode())[0] == sorted(want), topic
        for f, v, want in g["removes"]:
            assert eng.remove(f, v) is want
        T = Tables(eng.debug_tables())
        for topic, want in g["after_remove_matches"]:
            assert T.match(topic.encode())[0] == sorted(want), topic
    g = golden["derived_A5"]["trie"]
    eng = Engine(host_only=True)
    for f, v in g["inserts"]:
        eng.add(f, v)
    T = Tables(eng.debug_tables())
    for topic, want in g["matches"]:
        assert T.match(topic.encode())[0] == sorted(want), topic


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_differential_tables_vs_oracle(seed):
    rng = random.Random(seed)
    eng, tree = Engine(host_only=True), orc.TopicTree()
    live = []
    for step in range(600):
        if live and rng.random() < 0.3:
            f, v = rng.choice(live)
            if rng.random() < 0.3:
                v = rng.randint(0, 5)
            _apply(eng, tree, f, v, "remove")
        else:
            f, v = rand_filter(rng), rng.choice([0, 1, 2, 3, 7, 2**31 + 5, 2**32 - 1, rng.randint(0, 50)])
            _apply(eng, tree, f, v, "add")
            live.append((f, v))
        if step % 150 == 149:
            st = eng.stats()
            assert st["values"] == tree.values_size()
            assert st["nodes"] == tree.nodes_size()
            T = Tables(eng.debug_tables())
            for _ in range(150):
                t = rand_topic(rng)
                got, ctr = T.match(t.encode())
                want, wctr = tree.matches(t, with_counters=True)
                assert got == want, t
                if want is not None:
                    # counters of the table walk equal the reference walk's, except that the device keeps pruned
                    # (dead) nodes and skips probes through the Bloom mask: V/E may differ, F and M never do.
                    assert (ctr["F"], ctr["M"], ctr["L"], ctr["B"]) == (wctr["F"], wctr["M"], wctr["L"], wctr["B"])


def test_many_values_one_filter_and_growth():
    eng, tree = Engine(host_only=True), orc.TopicTree()
    for v in range(1, 3000):
        assert eng.add("/iot/x", v) == tree.insert("/iot/x", v)
        assert eng.add(f"/iot/{v}", v) == tree.insert(f"/iot/{v}", v)
    eng.add("/iot/#", 3); tree.insert("/iot/#", 3)
    T = Tables(eng.debug_tables())
    assert T.match(b"/iot/x")[0] == tree.matches("/iot/x")
    assert T.match(b"/iot/17")[0] == tree.matches("/iot/17") == [3, 17]
    st = eng.stats()
    assert st["values"] == tree.values_size() and st["nodes"] == tree.nodes_size()
    assert st["edge_slots"] >= 2 * st["edges"]          # load factor <= 0.5 after growth


def test_too_deep_filter_is_rejected():
    eng = Engine(host_only=True, max_levels=4)
    assert eng.add("a/b/c/d", 1)
    with pytest.raises(GpuMqttError) as ei:
        eng.add("a/b/c/d/e", 1)
    assert ei.value.code == N.GM_ERR_TOO_DEEP


def test_huge_value_set_uses_the_range_table():
    eng, tree = Engine(host_only=True), orc.TopicTree()
    import numpy as np
    from rmqtt_b200.engine import pack
    n = 70_000                                   # >= 65535 values on one filter node -> CNT_BIG + ranges[]
    blob, offs = pack(["big/+"] * n)
    vals = np.arange(n, dtype=np.uint32)
    assert eng.bulk_load(blob, offs, vals) == n
    tree.bulk_insert(blob, offs, vals)
    eng.add("big/#", 7); tree.insert("big/#", 7)
    T = Tables(eng.debug_tables())
    got, ctr = T.match(b"big/x")
    assert got == tree.matches("big/x") and ctr["M"] == n + 1
    assert eng.remove("big/+", 5) and tree.remove("big/+", 5)
    T = Tables(eng.debug_tables())
    assert T.match(b"big/x")[0] == tree.matches("big/x")
    assert eng.stats()["garbage_value_words"] == n


def test_wide_nodes_child_filter_has_no_false_negatives_across_flushes():
    """Nodes with > 48 literal children register their edges in the child filter; edges added before the node
    became wide are back-filled by the rebuild, later ones are inserted directly, and the filter grows."""
    rng = random.Random(3)
    eng, tree = Engine(host_only=True), orc.TopicTree()
    names = [f"k{i}" for i in range(400)]
    rng.shuffle(names)
    v = 0
    for chunk in range(8):                       # 8 flushes, the parent 'w' (and 'w/+') crosses the threshold in chunk 0
        for nm in names[chunk * 50:(chunk + 1) * 50]:
            for f in (f"w/{nm}", f"w/+/{nm}", f"{nm}/leaf"):
                v += 1
                assert eng.add(f, v) == tree.insert(f, v)
        if chunk == 3:
            for nm in names[:30]:
                assert eng.remove(f"w/{nm}", 1 + 3 * names.index(nm)) == tree.remove(f"w/{nm}", 1 + 3 * names.index(nm))
        T = Tables(eng.debug_tables())           # debug_tables() syncs like a flush
        for nm in names[:(chunk + 1) * 50:7] + ["nope", "k9999"]:
            for t in (f"w/{nm}", f"w/x/{nm}", f"{nm}/leaf", f"{nm}"):
                assert T.match(t.encode())[0] == tree.matches(t), t
    assert len(T.cfilter) >= 1024


def test_compact_drops_dead_nodes_and_garbage_keeps_matches():
    rng = random.Random(17)
    eng, tree = Engine(host_only=True), orc.TopicTree()
    live = []
    for _ in range(1500):
        f, v = rand_filter(rng), rng.randint(0, 40)
        _apply(eng, tree, f, v, "add")
        live.append((f, v))
    for f, v in live[:1000]:
        _apply(eng, tree, f, v, "remove")
    eng.flush()
    before = eng.stats()
    assert before["device_nodes"] > before["nodes"]            # pruned nodes linger as dead records
    eng.compact()
    after = eng.stats()
    assert after["nodes"] == after["device_nodes"] == tree.nodes_size()
    assert after["values"] == tree.values_size() and after["garbage_value_words"] == 0
    assert after["dict_entries"] <= before["dict_entries"]
    T = Tables(eng.debug_tables())
    for _ in range(300):
        t = rand_topic(rng)
        assert T.match(t.encode())[0] == tree.matches(t), t
    # the compacted trie keeps working incrementally
    _apply(eng, tree, "a/+/zz", 5, "add")
    T = Tables(eng.debug_tables())
    assert T.match(b"a/b/zz")[0] == tree.matches("a/b/zz")


@pytest.mark.parametrize("win_min,cap", [(3, 8), (4, 2), (5, 8)])
def test_windowed_edge_table_vs_oracle(monkeypatch, win_min, cap):
    """Edge table cut into many tiny windows (layout.h): subtrees outgrow their window, windows get halved, the
    table grows — the table walk must stay identical to the oracle throughout, and one topic's probes stay in
    few windows (its exact (level0, level1) subtree, the '+' variants and the hot window 0)."""
    monkeypatch.setenv("GM_WIN_MIN_SLOTS_LOG2", str(win_min))
    monkeypatch.setenv("GM_EDGE_WINDOWS_LOG2", str(cap))
    rng = random.Random(100 + win_min)
    eng, tree = Engine(host_only=True), orc.TopicTree()
    fs = []
    for i in range(3000):
        a, b = rng.randrange(6), rng.randrange(6)
        lv = [f"r{a}", f"s{b}"] + [f"x{rng.randrange(5)}" for _ in range(rng.randrange(0, 4))]
        if rng.random() < 0.3:
            lv[rng.randrange(len(lv))] = "+"
        if rng.random() < 0.1:
            lv.append("#")
        fs.append(("/".join(lv), i))
    fs += [(f"big/one/{i}/{j}", 10_000 + i * 40 + j) for i in range(40) for j in range(40)]   # one subtree >> a window
    rng.shuffle(fs)
    for k, (f, v) in enumerate(fs):
        _apply(eng, tree, f, v, "add")
        if k % 1500 == 1499 or k == len(fs) - 1:
            T = Tables(eng.debug_tables())
            for _ in range(120):
                t = "/".join([f"r{rng.randrange(6)}", f"s{rng.randrange(6)}"] + [f"x{rng.randrange(5)}" for _ in range(rng.randrange(0, 4))])
                assert T.match(t.encode())[0] == tree.matches(t), t
                assert len(T.windows_touched) <= 5, (t, T.windows_touched)     # window 0 + {r,+} x {s,+}
            for i in (0, 17, 39):
                t = f"big/one/{i}/{i}"
                assert T.match(t.encode())[0] == tree.matches(t), t
    for f, v in fs[::3]:
        _apply(eng, tree, f, v, "remove")
    T = Tables(eng.debug_tables())
    assert T.nwin_mask + 1 <= 1 << cap
    for _ in range(200):
        t = "/".join([f"r{rng.randrange(6)}", f"s{rng.randrange(6)}"] + [f"x{rng.randrange(5)}" for _ in range(rng.randrange(0, 4))])
        assert T.match(t.encode())[0] == tree.matches(t), t


def test_bulk_load_equals_one_by_one_inserts():
    """gm_bulk_load walks groups of 64 filters level-synchronously (prefetching); the result must be the trie that
    one-by-one inserts build: same reference statistics, same matches, invalid filters skipped, duplicates kept once."""
    import numpy as np
    from rmqtt_b200.engine import pack
    rng = random.Random(41)
    fs = [rand_filter(rng) for _ in range(5000)] + ["a/b/#/c", "x/+y", "dup/f", "dup/f", "r/s/t", "r/s/t"]
    rng.shuffle(fs)
    vals = [rng.randint(0, 20) for _ in fs]
    tree = orc.TopicTree()
    one = Engine(host_only=True)
    n_changed = 0
    for f, v in zip(fs, vals):
        try:
            n_changed += bool(one.add(f, v))
        except GpuMqttError:
            continue
        tree.insert(f, v)
    blob, offs = pack(fs)
    bulk = Engine(host_only=True)
    assert bulk.bulk_load(blob, offs, np.asarray(vals, dtype=np.uint32)) == n_changed
    sa, sb = one.stats(), bulk.stats()
    for k in ("values", "nodes", "edges", "dict_entries", "plus_nodes", "max_depth"):
        assert sa[k] == sb[k], k
    assert sb["values"] == tree.values_size() and sb["nodes"] == tree.nodes_size()
    T = Tables(bulk.debug_tables())
    for _ in range(400):
        t = rand_topic(rng)
        assert T.match(t.encode())[0] == tree.matches(t), t


@pytest.mark.parametrize("threads", [2, 3, 8, 64])
def test_parallel_bulk_load_builds_the_same_trie(monkeypatch, threads):
    """gm_bulk_load of a big batch runs on all host threads (host_trie.cpp insert_batch_parallel: per-thread tokenising with a
    first-occurrence merge of the new level strings, level-synchronous edges with one owner thread per window of the table,
    per-node-owner values) and the first flush after it too (child filter, value references, records).  GM_HOST_PAR_MIN=1
    sends a small batch down those paths: same reference statistics and token numbering as one-by-one inserts, same matches
    as the oracle, also when the batch lands on a trie that already holds filters, pruned nodes and multi-value sets."""
    import numpy as np
    from rmqtt_b200.engine import pack
    rng = random.Random(97 + threads)
    first = [rand_filter(rng) for _ in range(1500)] + ["keep/a", "keep/a", "gone/x/y", "gone/x/z"]
    fs = [rand_filter(rng) for _ in range(6000)] + ["a/b/#/c", "x/+y", "dup/f", "dup/f", "r/s/t", "r/s/t", "gone/x/y/deeper", "keep/a", "$SYS/x", "a/$b", "+/+/#", "#", ""]
    rng.shuffle(fs)
    fvals = [rng.randint(0, 20) for _ in first]
    vals = [rng.randint(0, 20) for _ in fs]
    tree = orc.TopicTree()
    one = Engine(host_only=True)
    for f, v in zip(first, fvals):
        try:
            one.add(f, v)
        except GpuMqttError:
            continue
        tree.insert(f, v)
    for f in ("gone/x/y", "gone/x/z"):                     # prune a branch: the bulk load below revives part of it
        for v in range(21):
            one.remove(f, v); tree.remove(f, v)
    n_changed = 0
    for f, v in zip(fs, vals):
        try:
            n_changed += bool(one.add(f, v))
        except GpuMqttError:
            continue
        tree.insert(f, v)
    one.flush()

    monkeypatch.setenv("GM_HOST_PAR_MIN", "1")
    monkeypatch.setenv("GM_HOST_THREADS", str(threads))
    bulk = Engine(host_only=True)
    b0, o0 = pack(first)
    bulk.bulk_load(b0, o0, np.asarray(fvals, dtype=np.uint32))
    bulk.flush()
    for f in ("gone/x/y", "gone/x/z"):
        for v in range(21):
            bulk.remove(f, v)
    blob, offs = pack(fs)
    assert bulk.bulk_load(blob, offs, np.asarray(vals, dtype=np.uint32)) == n_changed
    bulk.flush()
    sa, sb = one.stats(), bulk.stats()
    for k in ("values", "nodes", "edges", "dict_entries", "plus_nodes", "max_depth"):
        assert sa[k] == sb[k], k
    assert sa["value_words"] - sa["garbage_value_words"] == sb["value_words"] - sb["garbage_value_words"]   # live words of the multi-value sets
    assert sb["values"] == tree.values_size() and sb["nodes"] == tree.nodes_size()
    ta, tb = one.debug_tables(), bulk.debug_tables()
    da, db = ta["dict"], tb["dict"]                        # tokens are numbered by first occurrence, as one-by-one inserts number them:
    da, db = da[da[:, 0] != 0], db[db[:, 0] != 0]          # the same (token, level string) rows, whatever the table sizes
    assert np.array_equal(da[np.argsort(da[:, 0])], db[np.argsort(db[:, 0])])
    assert np.array_equal(ta["pool"], tb["pool"]) and len(tb["pool"]) > 0     # long level strings: same pool layout (token order)
    T = Tables(tb)
    for _ in range(600):
        t = rand_topic(rng)
        assert T.match(t.encode())[0] == tree.matches(t), t
    # and mutations keep working on the bulk-built tables
    for f, v in list(zip(fs, vals))[:200]:
        try:
            got = bulk.remove(f, v)
        except GpuMqttError:
            continue
        assert got == tree.remove(f, v)
    bulk.flush()
    T = Tables(bulk.debug_tables())
    assert bulk.stats()["nodes"] == tree.nodes_size()
    for _ in range(300):
        t = rand_topic(rng)
        assert T.match(t.encode())[0] == tree.matches(t), t


def test_parallel_bulk_load_makes_room_by_rehashing(monkeypatch):
    """The parallel bulk path settles room for a whole level of new edges at once (more windows / wider windows / a bigger
    table, host_trie.cpp step B3): a deliberately tight table (3 slots per filter, 8-slot windows) must re-hash on the way
    and still hold every edge where the walk looks for it."""
    import numpy as np
    from rmqtt_b200.engine import pack
    monkeypatch.setenv("GM_HOST_PAR_MIN", "1")
    monkeypatch.setenv("GM_HOST_THREADS", "4")
    monkeypatch.setenv("GM_EDGE_SLOTS_PER_FILTER", "3")
    monkeypatch.setenv("GM_WIN_MIN_SLOTS_LOG2", "3")
    monkeypatch.setenv("GM_EDGE_WINDOWS_LOG2", "5")
    rng = random.Random(5)
    fs = ["/".join([f"r{rng.randrange(6)}", f"s{rng.randrange(6)}"] + [f"x{rng.randrange(40)}" for _ in range(rng.randrange(1, 5))]) for _ in range(4000)]
    vals = [rng.randint(0, 9) for _ in fs]
    tree = orc.TopicTree()
    for f, v in zip(fs, vals):
        tree.insert(f, v)
    eng = Engine(host_only=True)
    blob, offs = pack(fs)
    eng.bulk_load(blob, offs, np.asarray(vals, dtype=np.uint32))
    eng.flush()
    st = eng.stats()
    assert st["values"] == tree.values_size() and st["nodes"] == tree.nodes_size()
    assert st["edge_slots"] > 16384                        # it grew beyond what the hint reserved (3 x 4000 -> 16384 slots)
    T = Tables(eng.debug_tables())
    for _ in range(400):
        t = "/".join([f"r{rng.randrange(6)}", f"s{rng.randrange(6)}"] + [f"x{rng.randrange(40)}" for _ in range(rng.randrange(0, 5))])
        assert T.match(t.encode())[0] == tree.matches(t), t


def test_child_adds_do_not_recopy_the_value_set_and_churn_is_compacted():
    """ADVICE r1 (host_trie.cpp make_ref): a node's multi-value set is re-published only when the SET changed, not when
    the node merely gained a child edge; replaced copies are garbage that auto-compaction bounds."""
    import numpy as np
    from rmqtt_b200.engine import pack
    eng, tree = Engine(host_only=True), orc.TopicTree()
    n = 1000
    blob, offs = pack(["pop/+"] * n)
    vals = np.arange(n, dtype=np.uint32)
    assert eng.bulk_load(blob, offs, vals) == n
    tree.bulk_insert(blob, offs, vals)
    eng.flush()
    assert eng.stats()["value_words"] == n
    for k in range(200):                          # 200 child adds below the popular node, one flush each
        f = f"pop/+/c{k}"
        assert eng.add(f, 5000 + k) == tree.insert(f, 5000 + k)
        eng.flush()
    st = eng.stats()
    assert st["value_words"] == n and st["garbage_value_words"] == 0, st
    for k in range(300):                          # subscribe/unsubscribe churn on the popular filter
        assert eng.add("pop/+", 10_000 + k) == tree.insert("pop/+", 10_000 + k)
        eng.flush()
        assert eng.remove("pop/+", 10_000 + k) == tree.remove("pop/+", 10_000 + k)
        eng.flush()
    st = eng.stats()
    assert st["value_words"] <= 2 * n + 65536 + 2 * n and st["garbage_value_words"] <= n + 65536 + 2 * n, st   # bounded by auto-compaction
    T = Tables(eng.debug_tables())
    for t in ("pop/x", "pop/x/c7", "pop/x/c199", "pop"):
        assert T.match(t.encode())[0] == tree.matches(t), t


def test_extra_trees_do_not_leak_into_the_subscription_trie():
    """gm_sub_add_tree keeps further TopicTrees as extra roots in the same tables; tree 0 (walked here from the global
    root by the Python model of the kernels) must not see them, before and after a compaction."""
    rng = random.Random(23)
    eng, tree = Engine(host_only=True), orc.TopicTree()
    v = 0
    for _ in range(400):
        f = rand_filter(rng)
        v += 1
        try:
            if rng.random() < 0.5:
                assert eng.add(f, v) == tree.insert(f, v)
            else:
                eng.add_tree(1 + v % 3, f, v)
        except GpuMqttError:
            pass
    for rnd in range(2):
        T = Tables(eng.debug_tables())
        for _ in range(300):
            t = rand_topic(rng)
            want = tree.matches(t)
            if want is not None:
                assert T.match(t.encode())[0] == want, t
        eng.compact()
    assert eng.stats()["values"] > tree.values_size()            # the extra trees' values are counted too
