"""CPU tier, world_size 2 over gloo: root-hash sharding + all-gatherv reproduce the unsharded result.
The per-rank matcher here is the ORACLE (no GPU on this tier); what is under test is the host logic of
rmqtt_b200/sharding.py — partitioning, replication of root-wildcard filters, the collective plumbing."""
import os
import random

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from rmqtt_b200 import sharding, workload as wl
from rmqtt_b200.engine import pack

from _gen import rand_filter, rand_topic


def _workload():
    rng = random.Random(99)
    cfg = wl.C2.scaled(n_subs=20_000, n_topics=3_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    extra_f = ["#", "+/#", "+/site-0003/#", "$SYS/#", "$SYS/+", "+"] + [rand_filter(rng) for _ in range(300)]
    extra_t = ["$SYS/x", "$SYS", "+", "#", "a/b"] + [rand_topic(rng) for _ in range(300)]
    filters = wl.unpack(sb, so) + [f.encode() for f in extra_f]
    values = np.concatenate([sv, np.arange(10**6, 10**6 + len(extra_f), dtype=np.uint32)])
    topics = wl.unpack(tb, to) + [t.encode() for t in extra_t]
    return filters, values, topics


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    filters, values, topics = _workload()
    fb, fo = pack(filters)
    tb, to = pack(topics)
    # this rank's shard of the subscription set and of the topic batch
    sfb, sfo, sval, _ = sharding.partition_filters(fb, fo, values, rank, world)
    stb, sto, tidx = sharding.partition_topics(tb, to, rank, world)
    tree = orc.TopicTree()
    for f, v in zip(wl.unpack(sfb, sfo), sval):
        try:
            tree.insert(f, int(v))
        except ValueError:
            pass
    res = tree.match_batch(stb, sto)
    counts = torch.from_numpy(res["counts"].copy())
    ids = torch.from_numpy(res["ids"].astype(np.int64))
    ti, ct, ia = sharding.all_gatherv_match_lists(torch.from_numpy(tidx.astype(np.int64)), counts, ids)
    q.put((rank, ti.numpy(), ct.numpy(), ia.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_unsharded_world2():
    world, port = 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank holds the same gathered result
    got.sort(key=lambda g: g[0])
    for a, b in zip(got[0][1:], got[1][1:]):
        assert (a == b).all()
    _, ti, ct, ia = got[0]
    # unsharded oracle
    filters, values, topics = _workload()
    tree = orc.TopicTree()
    for f, v in zip(filters, values):
        try:
            tree.insert(f, int(v))
        except ValueError:
            pass
    tb, to = pack(topics)
    want = tree.match_batch(tb, to)
    assert sorted(ti.tolist()) == list(range(len(topics)))          # every topic matched on exactly one shard
    starts = np.concatenate([[0], np.cumsum(np.maximum(ct, 0))])
    for k, t in enumerate(ti):
        assert ct[k] == want["counts"][t], topics[t]
        if ct[k] > 0:
            w = want["ids"][want["offsets"][t]:want["offsets"][t + 1]]
            assert sorted(ia[starts[k]:starts[k + 1]].tolist()) == sorted(w.tolist()), topics[t]


def test_shard_function_properties():
    from rmqtt_b200.engine import shard_of
    names = [wl.region_name(r) for r in range(64)]
    for g in (2, 4, 8):
        sh = [shard_of(n, g) for n in names]
        assert set(sh) <= set(range(g))
        assert max(np.bincount(sh, minlength=g)) <= 64 // g * 2.5       # no pathological imbalance on 64 roots
        assert shard_of(names[3] + b"/x/y", g) == sh[3]


# ---- retained path: retained topics sharded by root, root-wildcard filters answered by every shard ----
def _retain_workload():
    rng = random.Random(123)
    cfg = wl.C4.scaled(n_subs=15_000, n_topics=1_500)
    rb, ro, rv = wl.gen_retained(cfg)
    fb, fo = wl.gen_retain_filters(cfg)
    topics = wl.unpack(rb, ro) + [b"$SYS/broker/uptime", b"$SYS/x", b"$q/1", b"a//b", b"a"]
    values = np.concatenate([rv, np.arange(10**6, 10**6 + 5, dtype=np.uint32)])
    filters = wl.unpack(fb, fo) + [f.encode() for f in ["#", "+/#", "+", "$SYS/#", "$SYS/+", "+/+/+/+/+/+", "+/site-0002/#", "a/#", "a//+"]]
    filters += [f.encode() for f in (rand_filter(rng) for _ in range(200)) if orc.topic_parse(f) is not None]
    return topics, values, filters


def _retain_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    topics, values, filters = _retain_workload()
    rb, ro = pack(topics)
    fb, fo = pack(filters)
    srb, sro, sval, _ = sharding.partition_retained(rb, ro, values, rank, world)
    sfb, sfo, fidx = sharding.partition_retain_filters(fb, fo, rank, world)
    tree = orc.RetainTree()
    for t, v in zip(wl.unpack(srb, sro), sval):
        tree.insert(t, int(v))
    res = tree.match_batch(sfb, sfo)
    ti, ct, ia = sharding.all_gatherv_match_lists(torch.from_numpy(fidx.astype(np.int64)), torch.from_numpy(res["counts"].copy()),
                                                  torch.from_numpy(res["ids"].astype(np.int64)))
    q.put((rank, ti.numpy(), ct.numpy(), ia.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_retained_lookup_equals_unsharded_world2():
    world, port = 2, 31000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_retain_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda g: g[0])
    for a, b in zip(got[0][1:], got[1][1:]):
        assert (a == b).all()
    _, ti, ct, ia = got[0]
    topics, values, filters = _retain_workload()
    tree = orc.RetainTree()
    for t, v in zip(topics, values):
        tree.insert(t, int(v))
    # union over ranks of the per-filter hit lists == the unsharded lookup (wildcard roots are answered by both shards)
    merged = {i: [] for i in range(len(filters))}
    starts = np.concatenate([[0], np.cumsum(np.maximum(ct, 0))])
    answered = np.bincount(ti, minlength=len(filters))
    for k, f in enumerate(ti):
        merged[int(f)].extend(ia[starts[k]:starts[k + 1]].tolist())
    n_wild = 0
    for i, f in enumerate(filters):
        wild_root = f.split(b"/", 1)[0] in (b"+", b"#")
        n_wild += wild_root
        assert answered[i] == (world if wild_root else 1), f
        assert sorted(merged[i]) == tree.matches(f), f
    assert n_wild >= 6
    with pytest.raises(ValueError):
        sharding.partition_retained(*pack(["+/x"]), np.zeros(1, np.uint32), 0, 2)
