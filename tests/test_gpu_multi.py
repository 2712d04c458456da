"""GPU tier, 2 ranks (skipped on a 1-GPU box): root-hash sharded engines + the all-gatherv of libgpumqtt
(gm_partition_batch_device -> gm_match_batch_device_ex -> gm_allgatherv_device, all device buffers, NCCL inside the
library) == one unsharded engine, bit-exact per topic."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rmqtt_b200 import sharding, workload as wl
from rmqtt_b200.engine import Engine, MatchResult

pytestmark = pytest.mark.gpu

CFG = dict(n_subs=100_000, n_topics=20_000)


def sharded_match_and_gather(eng, rank, world, d_blob, d_offs, dev, stream):
    """The C5 data path of one rank; returns (index, spans, ids, sizes) of the gathered result (device tensors)."""
    n = d_offs.numel() - 1
    d_sel = torch.zeros(n, dtype=torch.int32, device=dev)
    k, counts = eng.partition_batch_device(d_blob, d_offs, world, rank, d_sel, stream)
    d_spans = torch.zeros((max(k, 1), 2), dtype=torch.int32, device=dev)
    d_status = torch.zeros(max(k, 1), dtype=torch.int32, device=dev)
    d_needed = torch.zeros(1, dtype=torch.int64, device=dev)
    d_ids = torch.empty(64 * max(k, 1) + 1024, dtype=torch.int32, device=dev)
    eng.match_batch_device_ex(d_blob, d_offs, d_spans, d_ids, d_needed, d_status, stream, d_sel=d_sel, n_sel=k)
    a_idx = torch.full((n,), -1, dtype=torch.int32, device=dev)
    a_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    a_ids = torch.empty(64 * n + 1024, dtype=torch.int32, device=dev)
    sizes = eng.allgatherv_device(d_sel, d_spans, k, d_ids, d_needed, a_idx, a_spans, a_ids, stream)
    torch.cuda.synchronize()
    return a_idx, a_spans, a_ids, sizes, counts


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", rank)
    stream = torch.cuda.current_stream().cuda_stream
    cfg = wl.C2.scaled(**CFG)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    fb, fo, fv, _ = sharding.partition_filters(sb, so, sv, rank, world)
    eng = Engine(device=rank)
    eng.bulk_load(fb, fo, fv)
    uid = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    d_blob, d_offs = torch.from_numpy(tb).to(dev), torch.from_numpy(to.view(np.int32)).to(dev)
    a_idx, a_spans, a_ids, sizes, counts = sharded_match_and_gather(eng, rank, world, d_blob, d_offs, dev, stream)
    m = int(sizes[:, 1].sum())
    # the same exchange fused into the match kernels over peer memory (CUDA IPC between the two processes)
    n = d_offs.numel() - 1
    h = eng.gather_create(world, rank, n, 64 * n)
    hs = [None] * world
    dist.all_gather_object(hs, h)
    eng.gather_connect(hs)
    d_sel = torch.zeros(n, dtype=torch.int32, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    fused = None
    for _ in range(3):
        k, _c = eng.partition_batch_device(d_blob, d_offs, world, rank, d_sel, stream)
        eng.match_gather_device(d_blob, d_offs, d_status, stream, d_sel=d_sel, n_sel=k)
        fused = eng.gather_result(stream)
        dist.barrier()                                     # nobody starts the next step while a peer still reads its block
    q.put((rank, a_idx.cpu().numpy(), a_spans.cpu().numpy().view(np.uint32), a_ids[:m].cpu().numpy().view(np.uint32), sizes, fused))
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
    for a, b in zip(got[0][1:5], got[1][1:5]):        # identical on every rank
        assert (np.asarray(a) == np.asarray(b)).all()
    _, idx, spans, ids, sizes, _f = got[0]
    cfg = wl.C2.scaled(**CFG)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    eng = Engine(device=0)
    eng.bulk_load(sb, so, sv)
    counts, want_ids = eng.match_batch(tb, to).canonical()
    assert sorted(idx.tolist()) == list(range(cfg.n_topics)) and int(sizes[:, 0].sum()) == cfg.n_topics
    order = np.argsort(idx)
    c, i = MatchResult(spans[order], ids, np.zeros(cfg.n_topics, np.int32), len(ids)).canonical()
    assert (c == counts).all() and (i == want_ids).all()
    for g in got:                                      # the fused (peer-memory) gather: complete and bit-exact on EVERY rank
        fcounts, fidx, fspans, fids = g[5]
        assert int(fcounts[:, 0].sum()) == cfg.n_topics and sorted(fidx.tolist()) == list(range(cfg.n_topics))
        fo = np.argsort(fidx)
        c2, i2 = MatchResult(fspans[fo], fids, np.zeros(cfg.n_topics, np.int32), int(fcounts[:, 1].sum())).canonical()
        assert (c2 == counts).all() and (i2 == want_ids).all()
