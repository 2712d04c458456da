"""CPU tier: the host side of GpuRouter (rmqtt_b200/csrc/router_host.cpp) — relations, counters, handle allocation and
reuse, the Id-equality rule of `remove` (router.rs:417-479) — against the oracle's Router restatement.  Matching
itself needs the device; here the handles stored in the (host-only) engine's tables are walked by the Python table
model and resolved through gmr_relation, which is exactly what the device result would be resolved through."""
import random

import pytest

from oracle import oracle as orc
from rmqtt_b200 import _native as N
from rmqtt_b200.engine import Engine, GpuMqttError
from rmqtt_b200.router import GpuRouter, Id

from _gen import rand_filter, rand_topic
from _tablewalk import Tables


def test_counters_and_id_rule_host_only():
    r, o = GpuRouter(Engine(host_only=True)), orc.Router()
    a, b = Id(1, "c1", 11), Id(1, "c2", 22)
    r.add("a/+", a); o.add_full("a/+", "c1", 0, 11, 1)
    r.add("a/+", b); o.add_full("a/+", "c2", 1, 22, 1)
    r.add("a/b", a); o.add_full("a/b", "c1", 2, 11, 1)
    with pytest.raises(GpuMqttError) as ei:
        r.add("a/#/b", a)
    assert ei.value.code == N.GM_ERR_INVALID_TOPIC
    assert (r.topics(), r.routes(), r.topics_tree()) == (o.topics(), o.routes(), 3) == (2, 3, 3)
    assert not r.remove("a/+", Id(1, "c1", 99))            # Id mismatch: nothing happens (router.rs:444-451)
    assert r.remove("a/+", a) and o.remove("a/+", "c1", 11) == 1
    assert (r.topics(), r.routes()) == (o.topics(), o.routes()) == (2, 2)
    assert r.remove("a/+", b) and o.remove("a/+", "c2", 22) == 1
    assert (r.topics(), r.routes()) == (o.topics(), o.routes()) == (1, 1)
    with pytest.raises(GpuMqttError) as ei:
        r.matches(Id(9, "pub"), "a/b")                      # no CPU fallback: a host-only engine refuses to match
    assert ei.value.code == N.GM_ERR_NO_DEVICE


@pytest.mark.parametrize("seed", [5, 6])
def test_random_relations_vs_oracle_router(seed):
    rng = random.Random(seed)
    eng = Engine(host_only=True)
    r, o = GpuRouter(eng), orc.Router()
    subs = {}
    clients = [Id(rng.randint(1, 3), f"c{k}", 100 + k) for k in range(20)]
    rel = 0
    for step in range(900):
        if subs and rng.random() < 0.3:
            (f, c) = rng.choice(list(subs))
            ident = subs[(f, c)]
            use = Id(ident.node_id, ident.client_id, 7) if rng.random() < 0.2 else ident
            got = r.remove(f, use)
            assert got == (o.remove(f, ident.client_id, use.tag) == 1)
            if got:
                del subs[(f, c)]
        else:
            f, cid = rand_filter(rng, 4), rng.choice(clients)
            try:
                r.add(f, cid)
            except GpuMqttError:
                continue
            o.add_full(f, cid.client_id, rel, cid.tag, cid.node_id); rel += 1
            subs[(f, cid.client_id)] = cid
        if step % 300 == 299:
            assert (r.topics(), r.routes()) == (o.topics(), o.routes())
            T = Tables(eng.debug_tables())
            for _ in range(150):
                t = rand_topic(rng, 5)
                handles, _ = T.match(t.encode())
                want = o.matches_full(t)
                if handles is None:
                    assert want is None
                    continue
                got = sorted("3|%d|%s|%s" % (subs[r._relation(h)].node_id, *r._relation(h)) for h in handles)
                assert got == sorted(want), t
