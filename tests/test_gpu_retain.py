"""GPU tier: retained-message lookup (gm_retain_match_batch) against the oracle's RetainTree restatement."""
import random

import numpy as np
import pytest

from oracle import oracle as orc
from rmqtt_b200 import _native as N
from rmqtt_b200 import workload as wl
from rmqtt_b200.engine import Engine, GpuMqttError, pack

from _gen import rand_filter, rand_topic
from test_gpu_parity import _assert_same

pytestmark = pytest.mark.gpu


def test_golden_retain_rs_451_475(golden):
    g = golden["retain_A3"]
    eng = Engine()
    for topic, v in g["inserts"]:
        eng.retain_set(topic, v & 0xFFFFFFFF)
    for f, want in g["matches"]:
        assert eng.retain_matches(f) == sorted(w & 0xFFFFFFFF for w in want), f
    for f, bad in g["not_matches"]:
        assert eng.retain_matches(f) != sorted(bad)
    for topic, v in g["more_inserts"]:
        eng.retain_set(topic, v & 0xFFFFFFFF)
    for f, want in g["more_matches"] + golden["derived_A5"]["retain_on_A3"]:
        assert eng.retain_matches(f) == sorted(w & 0xFFFFFFFF for w in want), f
    # retain(usize::MAX, |_| false): everything goes (retain.rs:479)
    for topic, _ in g["inserts"] + g["more_inserts"]:
        assert eng.retain_remove(topic) is not None
    st = eng.stats()
    assert st["retained_values"] == 0 and st["retained_nodes"] == 0
    assert eng.retain_matches("#") == []


@pytest.mark.parametrize("seed", [21, 22])
def test_random_differential_with_mutations(seed):
    rng = random.Random(seed)
    eng, tree = Engine(), orc.RetainTree()
    topics = []
    for rnd in range(5):
        for _ in range(500):
            if topics and rng.random() < 0.25:
                t = rng.choice(topics)
                try:
                    got = eng.retain_remove(t)
                except GpuMqttError:
                    continue
                assert got == tree.remove(t)
            else:
                t = rand_topic(rng, max_depth=6) if rng.random() < 0.9 else rand_filter(rng, 5)
                v = rng.randint(0, 2**32 - 2)
                try:
                    old = eng.retain_set(t, v)
                except GpuMqttError as ex:
                    assert ex.code == N.GM_ERR_INVALID_TOPIC
                    continue
                assert old == tree.remove(t)
                tree.insert(t, v)
                topics.append(t)
        filters = [rand_filter(rng, 7) for _ in range(1200)] + ["#", "+/#", "+", "$SYS/#", "+/+/+/+/+/+"]
        fb, fo = pack(filters)
        _assert_same(eng.retain_match_batch(fb, fo), tree.match_batch(fb, fo))
        st = eng.stats()
        assert st["retained_values"] == tree.values_size() and st["retained_nodes"] == tree.nodes_size()


def test_c4_scaled_parity():
    cfg = wl.C4.scaled(n_subs=400_000, n_topics=4_000)
    rb, ro, rv = wl.gen_retained(cfg)
    fb, fo = wl.gen_retain_filters(cfg)
    eng, tree = Engine(), orc.RetainTree()
    assert eng.retain_bulk_load(rb, ro, rv) == tree.bulk_insert(rb, ro, rv) == cfg.n_subs
    # a few whole-tree and root-level queries on top of the C4 mix
    extra_b, extra_o = pack(["#", "+/#", "reg-03/#", "+/+/+/sen-1/met-2/ch-0", "reg-01/+/+/+/+/+"])
    res = eng.retain_match_batch(fb, fo)
    want = tree.match_batch(fb, fo, nthreads=4)
    _assert_same(res, want)
    _assert_same(eng.retain_match_batch(extra_b, extra_o), tree.match_batch(extra_b, extra_o))
    assert res.needed == int(want["counts"].clip(0).sum())


def test_capacity_protocol_and_subscription_trie_coexist():
    eng = Engine()
    for i in range(300):
        eng.retain_set(f"a/{i}", i)
    eng.add("a/+", 7)                                  # the two trees share one dictionary
    fb, fo = pack(["a/+", "a/+/x", "a/#/b", "a/5"])
    with pytest.raises(GpuMqttError) as ei:
        eng.retain_match_batch(fb, fo, cap_ids=100)
    assert ei.value.code == N.GM_ERR_CAPACITY
    res = eng.retain_match_batch(fb, fo)
    assert res.sorted_list(0) == list(range(300)) and res.sorted_list(1) == [] and res.sorted_list(2) is None and res.sorted_list(3) == [5]
    assert eng.matches("a/17") == [7]


def test_message_storage_shape():
    """SURVEY §8(f) rank 3 — the message-storage plugin keeps a RetainTree<MsgID> keyed by `topic/<msg_id>`
    (rmqtt-plugins/rmqtt-message-storage/src/ram.rs:268-269, 294) and looks it up with the SUBSCRIBE filter plus one
    trailing `+` unless the filter ends in `#` (ram.rs:311-314): the same retained kernel with msg ids as values."""
    rng = random.Random(31)
    eng, tree = Engine(), orc.RetainTree()
    msg_id = 1
    stored = []
    for _ in range(3000):
        t = rand_topic(rng, 5)
        if orc.topic_parse(t) is None or "#" in t.split("/"):     # the reference pushes the id level without re-validating;
            continue                                              # through the string ABI a `#` level must stay last
        for _ in range(rng.randint(1, 3)):                       # several stored messages per topic
            key = f"{t}/{msg_id}"
            eng.retain_set(key, msg_id); tree.insert(key, msg_id)
            stored.append(key)
            msg_id += 1
    for key in stored[::4]:                                       # expiry (ram.rs:185)
        eng.retain_remove(key); tree.remove(key)
    filters = []
    for _ in range(1500):
        f = rand_filter(rng, 5)
        if orc.topic_parse(f) is None:
            continue
        filters.append(f if f.split("/")[-1] == "#" else f + "/+")
    fb, fo = pack(filters)
    _assert_same(eng.retain_match_batch(fb, fo), tree.match_batch(fb, fo))


def test_incremental_updates_on_device():
    """Small groups of set / remove between lookups: the retained image on the device is edited in place (32-byte entry
    and value-word patches shipped by the flush) instead of being re-flattened; every lookup stays bit-exact."""
    rng = random.Random(55)
    eng, tree = Engine(), orc.RetainTree()
    lv = ["a", "b", "c", "d", "e", "", "x" * 30, "dev-0000001"]
    def topic():
        t = [rng.choice(lv) for _ in range(rng.randint(1, 6))]
        if rng.random() < 0.1:
            t[0] = rng.choice(["$SYS", "$q"])
        return "/".join(t)
    live, v = [], 0
    for _ in range(2000):
        t = topic(); v += 1
        eng.retain_set(t, v); tree.remove(t); tree.insert(t, v); live.append(t)
    filters = ["/".join(rng.choice(lv + ["+", "+", "#"]) for _ in range(rng.randint(1, 6))) for _ in range(800)]
    filters = [f for f in filters if orc.topic_parse(f) is not None] + ["#", "+/#", "$SYS/#", "+/+/+"]
    fb, fo = pack(filters)
    _assert_same(eng.retain_match_batch(fb, fo), tree.match_batch(fb, fo))
    flattens0 = int(eng.debug_tables()["rstats"][0])
    for rnd in range(30):
        for _ in range(rng.randint(1, 25)):
            r = rng.random()
            if r < 0.35:
                t = rng.choice(live)
                assert eng.retain_remove(t) == tree.remove(t)
            else:
                t = rng.choice(live) if r < 0.55 else topic()
                v += 1
                assert eng.retain_set(t, v) == tree.remove(t)
                tree.insert(t, v); live.append(t)
        _assert_same(eng.retain_match_batch(fb, fo), tree.match_batch(fb, fo))
    st = eng.debug_tables()["rstats"]
    assert int(st[0]) <= flattens0 + 1 and int(st[1]) > 100, st.tolist()     # edited in place, at most one re-pack
