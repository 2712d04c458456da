#!/usr/bin/env python
"""Stress of the single-call front end: many producer threads, small windows, results checked against the engine's own
batch call; prints the status codes of any mismatch (diagnostics for intermittent failures)."""
import ctypes as C
import random
import sys
import threading
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from rmqtt_b200 import _native as N             # noqa: E402
from rmqtt_b200.engine import Engine, pack      # noqa: E402
from _gen import rand_filter, rand_topic        # noqa: E402

rng = random.Random(5)
eng = Engine()
for i in range(3000):
    try:
        eng.add(rand_filter(rng, 5), i)
    except Exception:
        pass
topics = [rand_topic(rng, 6) for _ in range(3000)]
tb, to = pack(topics)
ref = eng.match_batch(tb, to)
want = [ref.sorted_list(i) for i in range(len(topics))]
lib = N.lib()
bad = Counter()
for rnd in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    got = {}
    lock = threading.Lock()

    @N.GM_MATCH_CB
    def cb(user, cookie, status, ids, n_ids):
        lst = sorted(ids[i] for i in range(n_ids)) if status == 0 else ("ERR", status)
        with lock:
            got[cookie] = lst

    cfg = N.GmBatcherConfig(C.sizeof(N.GmBatcherConfig), rng.choice([1, 8, 64, 512]), rng.choice([0, 50, 200]), 2, cb, None)
    h = C.c_void_p()
    assert lib.gm_batcher_create(eng._h, C.byref(cfg), C.byref(h)) == 0
    P = 4

    def produce(k):
        for i in range(k, len(topics), P):
            b = topics[i].encode()
            lib.gm_submit(h, b, len(b), i)

    ths = [threading.Thread(target=produce, args=(k,)) for k in range(P)]
    [t.start() for t in ths]; [t.join() for t in ths]
    lib.gm_batcher_drain(h)
    lib.gm_batcher_destroy(h)
    for i in range(len(topics)):
        if got.get(i) != want[i]:
            bad[(rnd, str(got.get(i))[:40], str(want[i])[:40])] += 1
print("mismatches:", sum(bad.values()))
for k, v in list(bad.items())[:20]:
    print(v, k)
