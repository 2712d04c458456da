"""GPU tier, 2 ranks (skipped on a 1-GPU box): root-hash sharded engines + NCCL all-gatherv == one unsharded engine."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rmqtt_b200 import sharding, workload as wl
from rmqtt_b200.engine import Engine, pack

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = wl.C2.scaled(n_subs=100_000, n_topics=20_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    fb, fo, fv, _ = sharding.partition_filters(sb, so, sv, rank, world)
    stb, sto, tidx = sharding.partition_topics(tb, to, rank, world)
    eng = Engine(device=rank)
    eng.bulk_load(fb, fo, fv)
    res = eng.match_batch(stb, sto)
    counts, ids = res.canonical()
    dev = torch.device("cuda", rank)
    ti, ct, ia = sharding.all_gatherv_match_lists(torch.from_numpy(tidx.astype(np.int64)).to(dev), torch.from_numpy(counts).to(dev),
                                                  torch.from_numpy(ids.astype(np.int64)).to(dev))
    q.put((rank, ti.cpu().numpy(), ct.cpu().numpy(), ia.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_equals_single_engine():
    world, port = 2, 29500 + os.getpid() % 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for a, b in zip(got[0][1:], got[1][1:]):
        assert (a == b).all()
    _, ti, ct, ia = got[0]
    cfg = wl.C2.scaled(n_subs=100_000, n_topics=20_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    eng = Engine(device=0)
    eng.bulk_load(sb, so, sv)
    counts, ids = eng.match_batch(tb, to).canonical()
    starts = np.concatenate([[0], np.cumsum(counts)])
    gst = np.concatenate([[0], np.cumsum(ct)])
    assert sorted(ti.tolist()) == list(range(len(counts)))
    for k, t in enumerate(ti):
        assert ct[k] == counts[t]
        assert (ia[gst[k]:gst[k + 1]] == ids[starts[t]:starts[t + 1]]).all()
