"""CPU tier, hypothesis-driven (SURVEY §8c "extra pins"): arbitrary interleavings of add / remove over a level
alphabet that contains every special case of SURVEY §8a (Blank levels, `$` roots, literal `+` / `#` in topic
names, ill-formed wildcards, strings longer than the inline dictionary slot, duplicate ids) — the host-side
builders and the device table layout (through the pure-Python models of the kernels' walks) against the oracle."""
import os

import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as orc
from rmqtt_b200 import _native as N
from rmqtt_b200.engine import Engine, GpuMqttError

from _tablewalk import Tables
from test_retain_host import _plain_split, _tables

LEVELS = ["a", "b", "c", "", "+", "#", "$SYS", "$q", "a+", "#b", "x" * 27, "y" * 28, "z" * 40, "é", "dev-0000001"]
level = st.sampled_from(LEVELS)
path = st.lists(level, min_size=1, max_size=6).map("/".join)
value = st.sampled_from([0, 1, 2, 7, 2**31 + 5, 2**32 - 1])
ops = st.lists(st.tuples(st.sampled_from(["add", "add", "add", "remove"]), path, value), min_size=1, max_size=60)
COMMON = dict(deadline=None, max_examples=int(os.environ.get("GM_HYP_EXAMPLES", "400")), suppress_health_check=[HealthCheck.too_slow])


@settings(**COMMON)
@given(ops=ops, topics=st.lists(path, min_size=1, max_size=25), win=st.sampled_from([None, "3", "5"]))
def test_subscription_trie_tables_vs_oracle(ops, topics, win):
    if win is None:
        os.environ.pop("GM_WIN_MIN_SLOTS_LOG2", None)
    else:
        os.environ["GM_WIN_MIN_SLOTS_LOG2"] = win
    try:
        eng, tree = Engine(host_only=True), orc.TopicTree()
    finally:
        os.environ.pop("GM_WIN_MIN_SLOTS_LOG2", None)
    for op, f, v in ops:
        try:
            got = eng.add(f, v) if op == "add" else eng.remove(f, v)
        except GpuMqttError as ex:
            assert ex.code == N.GM_ERR_INVALID_TOPIC
            with pytest.raises(ValueError):
                tree.insert(f, v) if op == "add" else tree.remove(f, v)
            continue
        assert got == (tree.insert(f, v) if op == "add" else tree.remove(f, v)), (op, f, v)
    st_ = eng.stats()
    assert st_["values"] == tree.values_size() and st_["nodes"] == tree.nodes_size()
    T = Tables(eng.debug_tables())
    for t in topics:
        assert T.match(t.encode())[0] == tree.matches(t), t


retain_ops = st.lists(st.tuples(st.sampled_from(["set", "set", "set", "set", "remove", "remove", "flush"]), path, st.integers(0, 2**32 - 2)), min_size=1, max_size=60)


@settings(**COMMON)
@given(ops=retain_ops, filters=st.lists(path, min_size=1, max_size=25))
def test_retained_tree_tables_vs_oracle(ops, filters):
    """`flush` points in the sequence make the later operations edit the flattened image in place (retain_tree.h)."""
    eng, tree = Engine(host_only=True), orc.RetainTree()
    for op, t, v in ops:
        if op == "flush":
            eng.flush()
            continue
        try:
            if op == "set":
                eng.retain_set(t, v)
            else:
                eng.retain_remove(t)
        except GpuMqttError as ex:
            assert ex.code == N.GM_ERR_INVALID_TOPIC
            with pytest.raises(ValueError):
                tree.insert(t, v) if op == "set" else tree.remove(t)
            continue
        tree.insert(t, v) if op == "set" else tree.remove(t)
    R, t = _tables(eng)
    _plain_split(eng, R, t, ["$SYS", "$q"])
    for f in filters:
        assert R.match(f.encode()) == tree.matches(f), f


@settings(**COMMON)
@given(s=path)
def test_parse_validity_matches_oracle(s):
    """Topic::from_str validity (topic.rs:348-363) as the host parser sees it == the oracle's restatement."""
    eng = Engine(host_only=True)
    try:
        eng.add(s, 1)
        ok = True
    except GpuMqttError as ex:
        assert ex.code == N.GM_ERR_INVALID_TOPIC
        ok = False
    assert ok == (orc.topic_parse(s) is not None)
