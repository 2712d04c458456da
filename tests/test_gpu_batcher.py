"""GPU tier: the single-call front end (gm_batcher_*: MPSC queue + size/time window + dispatcher threads) and the
small-batch CUDA-graph path it rides on.  Router::matches is called once per PUBLISH in the reference
(rmqtt/src/router.rs:482-484); every answer delivered through the callback must equal the oracle's list."""
import ctypes as C
import random
import threading

import numpy as np
import pytest

from oracle import oracle as orc
from rmqtt_b200 import _native as N
from rmqtt_b200 import workload as wl
from rmqtt_b200.engine import Engine, pack

from _gen import rand_filter, rand_topic

pytestmark = pytest.mark.gpu


def _setup(n_filters=3000, seed=5):
    rng = random.Random(seed)
    eng, tree = Engine(), orc.TopicTree()
    for i in range(n_filters):
        f = rand_filter(rng, 5)
        try:
            eng.add(f, i)
        except Exception:
            continue
        tree.insert(f, i)
    return rng, eng, tree


@pytest.mark.parametrize("max_batch,wait_us,producers", [(1, 0, 1), (64, 200, 4), (4096, 500, 2)])
def test_batcher_delivers_every_topic_exactly_once_with_the_oracle_list(max_batch, wait_us, producers):
    rng, eng, tree = _setup()
    topics = [rand_topic(rng, 6) for _ in range(3000)] + ["bad/#/x", "a//b", "$SYS/x"]
    got = {}
    lock = threading.Lock()

    @N.GM_MATCH_CB
    def cb(user, cookie, status, ids, n_ids):
        lst = sorted(ids[i] for i in range(n_ids)) if status == 0 else None
        with lock:
            assert cookie not in got
            got[cookie] = lst

    cfg = N.GmBatcherConfig(C.sizeof(N.GmBatcherConfig), max_batch, wait_us, 2, cb, None)
    h = C.c_void_p()
    lib = N.lib()
    assert lib.gm_batcher_create(eng._h, C.byref(cfg), C.byref(h)) == 0

    def produce(k):
        for i in range(k, len(topics), producers):
            b = topics[i].encode()
            assert lib.gm_submit(h, b, len(b), i) == 0

    ths = [threading.Thread(target=produce, args=(k,)) for k in range(producers)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert lib.gm_batcher_drain(h) == 0
    lib.gm_batcher_destroy(h)
    assert len(got) == len(topics)
    for i, t in enumerate(topics):
        want = tree.matches(t)
        assert got[i] == (sorted(want) if want is not None else None), t


def test_small_batch_graph_path_equals_pipelined_path_across_flushes():
    """Batches of 1..2048 topics take the one-launch CUDA-graph path; it must give the same lists as the pipelined path
    (graphs off) and follow mutations (the graph is re-captured when a flush changed the table view)."""
    rng, eng, tree = _setup(2000, seed=9)
    topics = [rand_topic(rng, 6) for _ in range(1500)]
    tb, to = pack(topics)
    for rnd in range(3):
        for sz in (1, 7, 64, 65, 1500):
            sb, so = pack(topics[:sz])
            want = tree.match_batch(sb, so)
            eng.debug_knob("small_graphs", 1)
            c1, i1 = eng.match_batch(sb, so).canonical()
            d1, j1 = eng.match_batch_via_desc(sb, so).canonical()
            eng.debug_knob("small_graphs", 0)
            c0, i0 = eng.match_batch(sb, so).canonical()
            assert (c1 == c0).all() and (i1 == i0).all() and (d1 == c0).all() and (j1 == i0).all()
            assert (c1 == want["counts"]).all()
        # a batch beyond the graph tiers takes the pipelined path on the same context and re-allocates its (larger) scratch:
        # the captured graphs hold the old pointers and must be re-captured (regression: stale scratch after growth)
        big = (topics * (5 + rnd))[:6000 + 3000 * rnd]
        bb, bo = pack(big)
        eng.debug_knob("small_graphs", 1)
        cb_, ib_ = eng.match_batch(bb, bo).canonical()
        wb = tree.match_batch(bb, bo)
        assert (cb_ == wb["counts"]).all()
        for sz in (1, 64, 700):
            sb, so = pack(topics[:sz])
            c1, i1 = eng.match_batch(sb, so).canonical()
            assert (c1 == tree.match_batch(sb, so)["counts"]).all()
        for k in range(50):                       # mutate, so that the next round runs on a new view
            f = rand_filter(rng, 5)
            try:
                assert eng.add(f, 100_000 + rnd * 100 + k) == tree.insert(f, 100_000 + rnd * 100 + k)
            except Exception:
                pass
    eng.debug_knob("small_graphs", 1)


def test_latency_probe_reports_sane_numbers():
    cfg = wl.C2.scaled(n_subs=50_000, n_topics=4_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    eng = Engine(filters_hint=cfg.n_subs)
    eng.bulk_load(sb, so, sv)
    eng.flush()
    lat = N.GmLatency()
    for burst, rounds in ((1, 200), (32, 50), (1024, 10)):
        rc = N.lib().gm_batcher_probe(eng._h, tb.ctypes.data, to.ctypes.data, cfg.n_topics, burst, rounds, 0, C.byref(lat))
        assert rc == 0
        d = lat.as_dict()
        assert d["samples"] == burst * rounds and 0 < d["p50_us"] <= d["p99_us"] <= d["max_us"] < 5e6 and d["topics_per_s"] > 0


def test_submit_publish_takes_the_raw_packet():
    """gm_submit_publish: the topic is taken straight out of the PUBLISH packet (v3.1.1 / v5 layout) into the batch."""
    rng, eng, tree = _setup(800, seed=3)
    topics = [rand_topic(rng, 5) for _ in range(300)]
    got = {}

    @N.GM_MATCH_CB
    def cb(user, cookie, status, ids, n_ids):
        got[cookie] = sorted(ids[i] for i in range(n_ids)) if status == 0 else None

    cfg = N.GmBatcherConfig(C.sizeof(N.GmBatcherConfig), 32, 100, 2, cb, None)
    h = C.c_void_p()
    lib = N.lib()
    assert lib.gm_batcher_create(eng._h, C.byref(cfg), C.byref(h)) == 0
    for i, t in enumerate(topics):
        tb = t.encode()
        qos = i % 3
        var = len(tb).to_bytes(2, "big") + tb + (b"\x00\x09" if qos else b"") + (b"\x00" if i % 2 else b"") + b"payload" * (i % 5)
        rem, enc = len(var), bytearray()
        while True:
            b = rem & 0x7F
            rem >>= 7
            enc.append(b | (0x80 if rem else 0))
            if not rem:
                break
        pkt = bytes([0x30 | (qos << 1)]) + bytes(enc) + var
        buf = (C.c_uint8 * len(pkt)).from_buffer_copy(pkt)
        assert lib.gm_submit_publish(h, buf, len(pkt), i) == 0
    bad = (C.c_uint8 * 4).from_buffer_copy(b"\x82\x02\x00\x01")        # a SUBSCRIBE is not a PUBLISH
    assert lib.gm_submit_publish(h, bad, 4, 999) == N.GM_ERR_INVALID_ARG
    assert lib.gm_batcher_drain(h) == 0
    lib.gm_batcher_destroy(h)
    for i, t in enumerate(topics):
        want = tree.matches(t)
        assert got[i] == (sorted(want) if want is not None else None), t
