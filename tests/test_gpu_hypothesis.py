"""GPU tier, hypothesis-driven: the same arbitrary add / remove interleavings as tests/test_hypothesis_cpu.py, but
matched by the CUDA kernels through the C ABI (gm_match_batch / gm_retain_match_batch) and compared bit-exactly
(sorted multisets) with the oracle.  One long-lived engine per test keeps example cost at one flush + one batch."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle as orc
from rmqtt_b200 import _native as N
from rmqtt_b200.engine import Engine, GpuMqttError, pack

from test_hypothesis_cpu import ops, path, retain_ops

pytestmark = pytest.mark.gpu
COMMON = dict(deadline=None, max_examples=int(os.environ.get("GM_HYP_GPU_EXAMPLES", "60")), suppress_health_check=[HealthCheck.too_slow])


def _lists(res):
    return [res.sorted_list(i) for i in range(len(res))]


@settings(**COMMON)
@given(ops=ops, topics=st.lists(path, min_size=1, max_size=40), tiny=st.booleans())
def test_subscription_trie_cuda_vs_oracle(ops, topics, tiny):
    if tiny:
        os.environ["GM_WIN_MIN_SLOTS_LOG2"] = "3"
    try:
        eng, tree = Engine(), orc.TopicTree()
    finally:
        os.environ.pop("GM_WIN_MIN_SLOTS_LOG2", None)
    for op, f, v in ops:
        try:
            got = eng.add(f, v) if op == "add" else eng.remove(f, v)
        except GpuMqttError as ex:
            assert ex.code == N.GM_ERR_INVALID_TOPIC
            continue
        assert got == (tree.insert(f, v) if op == "add" else tree.remove(f, v)), (op, f, v)
    tb, to = pack(topics)
    got = _lists(eng.match_batch(tb, to))
    for t, g in zip(topics, got):
        assert g == tree.matches(t), t


@settings(**COMMON)
@given(ops=retain_ops, filters=st.lists(path, min_size=1, max_size=40))
def test_retained_tree_cuda_vs_oracle(ops, filters):
    eng, tree = Engine(), orc.RetainTree()
    for op, t, v in ops:
        if op == "flush":              # ship what is staged: later operations patch the device image in place
            eng.flush()
            continue
        try:
            eng.retain_set(t, v) if op == "set" else eng.retain_remove(t)
        except GpuMqttError as ex:
            assert ex.code == N.GM_ERR_INVALID_TOPIC
            continue
        tree.insert(t, v) if op == "set" else tree.remove(t)
    fb, fo = pack(filters)
    got = _lists(eng.retain_match_batch(fb, fo))
    for f, g in zip(filters, got):
        assert g == tree.matches(f), f
