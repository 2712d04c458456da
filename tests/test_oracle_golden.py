"""Pins the CPU oracle (oracle/oracle.cpp) against every assertion the reference's own unit tests
make for the hot path: rmqtt/src/trie.rs:415-513, rmqtt/src/retain.rs:449-482, rmqtt/src/topic.rs:429-586.
"""
from oracle import oracle as orc


def test_parse_topic_rs_487_541(golden):
    for s, kinds in golden["parse_A0"]["ok"]:
        assert orc.topic_parse(s) == kinds, s
    for s in golden["parse_A0"]["err"]:
        assert orc.topic_parse(s) is None, s


def test_matches_str_topic_rs_554_586(golden):
    for filt, topic, want in golden["matches_str_A4"]:
        assert orc.matches_str(filt, topic) is want, (filt, topic)


def test_trie_rs_417_449(golden):
    g = golden["trie_A1"]
    t = orc.TopicTree()
    for f, v in g["inserts"]:
        t.insert(f, v)
    for topic, want in g["matches"]:
        assert t.matches(topic) == sorted(want), topic
    for topic, bad in g["not_matches"]:
        assert t.matches(topic) != sorted(bad)
    for f, v, want in g["removes"]:
        assert t.remove(f, v) is want, (f, v)
    for topic, bad in g["after_remove_not_matches"]:
        assert t.matches(topic) != sorted(bad)
    for topic, want in g["after_remove_matches"]:
        assert t.matches(topic) == sorted(want)


def test_trie_rs_452_498(golden):
    g = golden["trie_A2"]
    t = orc.TopicTree()
    for f, v in g["inserts"]:
        t.insert(f, v)
    r = g["range_inserts"]
    for v in range(r["lo"], r["hi"]):
        t.insert(r["pattern_each"].format(v=v), v)
    for v in range(r["lo"], r["hi"]):
        t.insert(r["pattern_same"], v)
    # 7 + 9999 (of which /iot/10 -> 10 and /iot/11 -> 11 already present) + 9999
    assert t.values_size() == 7 + 9999 - 2 + 9999
    for topic, want in g["matches"]:
        assert t.matches(topic) == sorted(want), topic
    for topic in g["is_match"]:
        assert len(t.matches(topic)) > 0
    assert t.matches("/iot/x") == sorted(list(range(1, 10000)) + [3])
    for f, v in g["stage2_inserts"]:
        t.insert(f, v)
    for topic, want in g["stage2_matches"]:
        assert t.matches(topic) == sorted(want), topic
    for f, v in g["stage3_inserts"]:
        t.insert(f, v)
    for topic, want in g["stage3_matches"]:
        assert t.matches(topic) == sorted(want), topic


def test_trie_rs_503_512_values_size(golden):
    # TopicTree<()>: all values are the unit value, so the set per node holds at most one element.
    t = orc.TopicTree()
    for f in golden["trie_values_size"]["inserts"]:
        t.insert(f, 0)
    assert t.values_size() == golden["trie_values_size"]["values_size_unit"]


def test_retain_rs_451_475(golden):
    g = golden["retain_A3"]
    t = orc.RetainTree()
    for topic, v in g["inserts"]:
        t.insert(topic, v)
    for f, want in g["matches"]:
        assert t.matches(f) == sorted(want), f
    for f, bad in g["not_matches"]:
        assert t.matches(f) != sorted(bad)
    for topic, v in g["more_inserts"]:
        t.insert(topic, v)
    for f, want in g["more_matches"]:
        assert t.matches(f) == sorted(want), f
    for f, want in golden["derived_A5"]["retain_on_A3"]:
        assert t.matches(f) == sorted(want), f
    # retain(usize::MAX, |_| false) removes everything (retain.rs:479)
    n = t.values_size()
    assert n == 12
    for topic, _ in g["inserts"] + g["more_inserts"]:
        assert t.remove(topic) is not None
    assert t.values_size() == 0 and t.nodes_size() == 0


def test_derived_a5_trie(golden):
    g = golden["derived_A5"]["trie"]
    t = orc.TopicTree()
    for f, v in g["inserts"]:
        t.insert(f, v)
    for topic, want in g["matches"]:
        assert t.matches(topic) == sorted(want), topic


def test_counters_definition():
    t = orc.TopicTree()
    t.insert("a/b", 1)
    t.insert("a/+", 2)
    t.insert("a/#", 3)
    res, c = t.matches("a/b", with_counters=True)
    assert res == [1, 2, 3]
    # visited: root, a, a/+ , a/b -> V=4; non-empty path at root and a -> E=2; F=3 filters, M=3 ids
    assert (c["V"], c["E"], c["F"], c["M"], c["L"], c["B"]) == (4, 2, 3, 3, 2, 3)


def test_router_restatement_router_rs_417_479():
    r = orc.Router()
    assert r.add("a/+", "c1", 1, 11)
    assert r.add("a/+", "c2", 2, 22)
    assert r.add("a/b", "c1", 3, 11)
    assert not r.add("a/#/b", "c1", 4, 11)
    assert (r.topics(), r.routes(), r.topics_tree()) == (2, 3, 2)
    assert r.matches("a/b") == [1, 2, 3]
    assert r.matches("a/b+") is None
    assert r.remove("a/+", "c1", 99) == 0          # Id mismatch (router.rs:444-451)
    assert r.remove("a/+", "c1", 11) == 1
    assert r.matches("a/b") == [2, 3]
    assert r.remove("a/+", "c2", 22) == 1          # last client -> filter pruned from the trie
    assert (r.topics(), r.routes(), r.topics_tree()) == (1, 1, 1)
    assert r.matches("a/b") == [3]
    # re-adding the same client replaces its relation (HashMap::insert), count unchanged
    assert r.add("a/b", "c1", 7, 11)
    assert r.routes() == 1 and r.matches("a/b") == [7]
