"""GPU tier: descriptor-mode output (gm_match_batch_desc / gm_values_view / gm_desc_expand), sub-batch selection
(gm_match_batch_device_ex with d_sel), the device partition kernel and the all-gatherv entry point of libgpumqtt.

The collective is exercised here on a one-rank NCCL communicator (a 1-GPU box cannot hold two ranks — NCCL refuses
two ranks on one device); its multi-rank correctness is checked in tests/test_gpu_multi.py (needs 2 GPUs) and by
bench.py's `parity_check` on every multi-GPU run."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from rmqtt_b200 import _native as N
from rmqtt_b200 import sharding, workload as wl
from rmqtt_b200.engine import Engine, MatchResult, pack

pytestmark = pytest.mark.gpu


def _canon_oracle(want):
    counts = want["counts"]
    seg = np.repeat(np.arange(len(counts), dtype=np.int64), np.maximum(counts, 0))
    return counts, want["ids"][np.lexsort((want["ids"], seg))]


def _assert_same(res, want):
    counts, ids = res.canonical()
    wc, wi = _canon_oracle(want)
    assert (counts == wc).all(), f"counts differ at {np.nonzero(counts != wc)[0][:5]}"
    assert len(ids) == len(wi) and (ids == wi).all()


def test_descriptor_mode_equals_oracle_incl_deferred_and_huge_sets():
    cfg = wl.C2.scaled(n_subs=150_000, n_topics=30_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    eng, tree = Engine(filters_hint=cfg.n_subs), orc.TopicTree()
    eng.bulk_load(sb, so, sv); tree.bulk_insert(sb, so, sv)
    # a heavy hitter beyond the 16-bit count (ranges[]), a deep filter (deferred kernel), single-value sets, a '$' root
    n_big = 70_000
    bb, bo = pack(["reg-00/+/#"] * n_big)
    bv = np.arange(n_big, dtype=np.uint32) + 20_000_000
    eng.bulk_load(bb, bo, bv); tree.bulk_insert(bb, bo, bv)
    deep = "/".join(f"l{i}" for i in range(14))
    for i, f in enumerate([deep, deep + "/#", "$SYS/#", "+/+/#", "#"]):
        assert eng.add(f, 30_000_000 + i) == tree.insert(f, 30_000_000 + i)
    xb, xo = pack([deep, deep + "/x", "$SYS/a", "reg-00/site-0001", "reg-00/x/y/z", "a//b", "bad/#/x"])
    for blob, offs in ((tb, to), (xb, xo)):
        want = tree.match_batch(blob, offs)
        _assert_same(eng.match_batch_via_desc(blob, offs), want)
        _assert_same(eng.match_batch(blob, offs), want)
    # descriptors are one per matched FILTER: their number equals the oracle's F counter
    spans, descs, status, needed = eng.match_batch_desc(tb, to)
    want = tree.match_batch(tb, to, want_ids=False)
    assert needed == want["counters"]["F"] and int(spans[:, 1].sum()) == needed
    # capacity protocol in descriptor units
    from rmqtt_b200.engine import GpuMqttError
    from rmqtt_b200 import _native as N
    with pytest.raises(GpuMqttError) as ei:
        eng.match_batch_desc(tb, to, cap=needed - 1)
    assert ei.value.code == N.GM_ERR_CAPACITY


def test_values_view_is_stable_across_appends_and_epoch_moves_on_compaction():
    eng = Engine()
    for v in range(10):
        eng.add("a/+", v)
    eng.flush()
    vals, rng, ep = eng.values_view()
    base = vals.ctypes.data
    spans, descs, _, _ = eng.match_batch_desc(*pack(["a/b"]))
    assert descs[0, 1] == 10 and sorted(vals[descs[0, 0]:descs[0, 0] + 10].tolist()) == list(range(10))
    for v in range(100_000):                      # grows `values` far beyond its first pages: the base must not move
        eng.add(f"g/{v % 50}/+", v)
    eng.flush()
    vals2, _, ep2 = eng.values_view()
    assert vals2.ctypes.data == base and ep2 == ep
    assert sorted(vals2[descs[0, 0]:descs[0, 0] + 10].tolist()) == list(range(10))      # the old reference still reads the same set
    eng.compact()
    assert eng.values_view()[2] != ep or eng.values_view()[0].ctypes.data == base


def test_selection_and_device_partition_match_the_host_partition():
    cfg = wl.C2.scaled(n_subs=80_000, n_topics=20_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    n = cfg.n_topics
    dev = torch.device("cuda")
    stream = torch.cuda.current_stream().cuda_stream
    eng = Engine(filters_hint=cfg.n_subs)
    eng.bulk_load(sb, so, sv)
    full_counts, full_ids = eng.match_batch(tb, to).canonical()
    starts = np.concatenate([[0], np.cumsum(full_counts)])
    d_blob, d_offs = torch.from_numpy(tb).to(dev), torch.from_numpy(to.view(np.int32)).to(dev)
    d_sel = torch.zeros(n, dtype=torch.int32, device=dev)
    d_shard = torch.zeros(n, dtype=torch.int32, device=dev)
    world = 4
    host_shard = sharding.shard_ids(tb, to, world)
    seen = []
    for rank in range(world):
        k, counts = eng.partition_batch_device(d_blob, d_offs, world, rank, d_sel, stream, d_shard=d_shard)
        assert (d_shard.cpu().numpy().view(np.uint32) == host_shard).all()
        assert (counts == np.bincount(host_shard, minlength=world)).all() and k == counts[rank]
        sel = d_sel[:k].cpu().numpy()
        assert sorted(sel.tolist()) == np.nonzero(host_shard == rank)[0].tolist()
        seen.append(sel)
        # match only the selected rows
        d_spans = torch.zeros((k, 2), dtype=torch.int32, device=dev)
        d_status = torch.zeros(k, dtype=torch.int32, device=dev)
        d_needed = torch.zeros(1, dtype=torch.int64, device=dev)
        d_ids = torch.empty(64 * k + 1024, dtype=torch.int32, device=dev)
        eng.match_batch_device_ex(d_blob, d_offs, d_spans, d_ids, d_needed, d_status, stream, d_sel=d_sel, n_sel=k)
        torch.cuda.synchronize()
        res = MatchResult(d_spans.cpu().numpy().view(np.uint32), d_ids.cpu().numpy().view(np.uint32), d_status.cpu().numpy(), int(d_needed.item()))
        c, i = res.canonical()
        assert (c == full_counts[sel]).all()
        st = np.concatenate([[0], np.cumsum(c)])
        for row in range(0, k, max(1, k // 500)):
            t = sel[row]
            assert (i[st[row]:st[row + 1]] == full_ids[starts[t]:starts[t + 1]]).all()
    assert sorted(np.concatenate(seen).tolist()) == list(range(n))


def test_allgatherv_entry_point_on_a_one_rank_communicator():
    cfg = wl.C2.scaled(n_subs=50_000, n_topics=5_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    n = cfg.n_topics
    dev = torch.device("cuda")
    stream = torch.cuda.current_stream().cuda_stream
    eng = Engine(filters_hint=cfg.n_subs)
    eng.bulk_load(sb, so, sv)
    eng.comm_init(Engine.comm_unique_id(), 0, 1)
    d_blob, d_offs = torch.from_numpy(tb).to(dev), torch.from_numpy(to.view(np.int32)).to(dev)
    d_sel = torch.zeros(n, dtype=torch.int32, device=dev)
    k, _ = eng.partition_batch_device(d_blob, d_offs, 1, 0, d_sel, stream)
    assert k == n
    d_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_needed = torch.zeros(1, dtype=torch.int64, device=dev)
    d_ids = torch.empty(64 * n, dtype=torch.int32, device=dev)
    eng.match_batch_device_ex(d_blob, d_offs, d_spans, d_ids, d_needed, d_status, stream, d_sel=d_sel, n_sel=k)
    a_idx = torch.full((n,), -1, dtype=torch.int32, device=dev)
    a_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    a_ids = torch.empty(64 * n, dtype=torch.int32, device=dev)
    sizes = eng.allgatherv_device(d_sel, d_spans, k, d_ids, d_needed, a_idx, a_spans, a_ids, stream)
    torch.cuda.synchronize()
    m = int(d_needed.item())
    assert sizes.tolist() == [[n, m]]
    assert (a_idx.cpu().numpy() == d_sel.cpu().numpy()).all() and (a_spans.cpu().numpy() == d_spans.cpu().numpy()).all()
    assert (a_ids[:m].cpu().numpy() == d_ids[:m].cpu().numpy()).all()
    # and the gathered lists are the oracle's lists, addressed by global topic index
    tree = orc.TopicTree(); tree.bulk_insert(sb, so, sv)
    want = tree.match_batch(tb, to)
    idx = a_idx.cpu().numpy()
    sp = a_spans.cpu().numpy().view(np.uint32)
    ids = a_ids.cpu().numpy().view(np.uint32)
    order = np.argsort(idx)
    res = MatchResult(sp[order], ids, np.zeros(n, np.int32), m)
    _assert_same(res, want)


def test_fused_gather_on_one_rank_equals_the_oracle():
    """gm_match_gather_device: the match kernels publish straight into the gathered block (world 1: this rank's own block;
    the peer-memory path proper needs 2 GPUs, tests/test_gpu_multi.py).  Includes deferred (slow-kernel) topics."""
    cfg = wl.C2.scaled(n_subs=60_000, n_topics=6_000)
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics(cfg)
    eng, tree = Engine(filters_hint=cfg.n_subs), orc.TopicTree()
    eng.bulk_load(sb, so, sv); tree.bulk_insert(sb, so, sv)
    deep = "/".join(f"l{i}" for i in range(12))
    for i, f in enumerate([deep, deep + "/#", "#"]):
        assert eng.add(f, 9_000_000 + i) == tree.insert(f, 9_000_000 + i)
    xb, xo = pack(wl.unpack(tb, to) + [deep.encode(), (deep + "/x").encode(), b"bad/#/x"])
    n = len(xo) - 1
    dev = torch.device("cuda")
    stream = torch.cuda.current_stream().cuda_stream
    d_blob, d_offs = torch.from_numpy(xb).to(dev), torch.from_numpy(xo.view(np.int32)).to(dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_sel = torch.zeros(n, dtype=torch.int32, device=dev)
    k, _ = eng.partition_batch_device(d_blob, d_offs, 1, 0, d_sel, stream)
    eng.gather_create(1, 0, n, 64 * n)
    eng.gather_connect([b"\0" * N.GM_IPC_HANDLE_BYTES])
    for _ in range(2):                                       # two steps: the epoch barrier must pass both times
        eng.match_gather_device(d_blob, d_offs, d_status, stream, d_sel=d_sel, n_sel=k)
    counts, idx, spans, ids = eng.gather_result(stream)
    assert counts.tolist() == [[n, int(spans[:, 1].sum())]]
    order = np.argsort(idx)
    assert (idx[order] == np.arange(n)).all()
    st = d_status.cpu().numpy()[np.argsort(d_sel.cpu().numpy()[:k])]      # status is per ROW; back to topic order
    res = MatchResult(spans[order], ids, st, int(counts[0, 1]))
    _assert_same(res, tree.match_batch(xb, xo))
    eng.gather_destroy()
