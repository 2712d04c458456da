#!/usr/bin/env python
"""Secondary measurements for BASELINE.md §3: configs C1, C2, C4 (C3 is bench.py's headline).
For each config: GPU kernels (device-resident, CUDA events via the engine's timing ring), CPU oracle on the
same box (1 thread and all threads), parity of counts.  Prints one JSON object per config."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc            # noqa: E402  (this script is a measurement harness, like bench.py's cpu_baseline leg)
from rmqtt_b200 import workload as wl       # noqa: E402
from rmqtt_b200.engine import Engine        # noqa: E402

dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
threads = orc.hardware_threads()


def publish_config(cfg, reps=20, zipf=False):
    sb, so, sv = wl.gen_subs(cfg)
    tb, to = wl.gen_topics_zipf(cfg) if zipf else wl.gen_topics(cfg)
    n = len(to) - 1
    eng = Engine(filters_hint=len(sv))
    t0 = time.time(); eng.bulk_load(sb, so, sv); eng.flush(); build = time.time() - t0
    d_blob, d_offs = torch.from_numpy(tb).to(dev), torch.from_numpy(to.view(np.int32)).to(dev)
    d_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev); d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_needed = torch.zeros(1, dtype=torch.int64, device=dev)
    d_ids = torch.empty(64 * n + 1024, dtype=torch.int32, device=dev)
    w = eng.match_batch_device(d_blob, d_offs, d_spans, d_ids, d_needed, d_status, stream, work=True)
    need = int(d_needed.item())
    if need > d_ids.numel():
        d_ids = torch.empty(need + 1024, dtype=torch.int32, device=dev)
    for _ in range(reps + 3):
        eng.match_batch_device(d_blob, d_offs, d_spans, d_ids, d_needed, d_status, stream)
    torch.cuda.synchronize()
    k = eng.kernel_ms(reps).mean(axis=0)
    tree = orc.TopicTree(); tree.bulk_insert(sb, so, sv, nthreads=min(threads, 64))
    o1 = tree.match_batch(tb, to, nthreads=1, want_ids=False)
    oN = tree.match_batch(tb, to, nthreads=threads, want_ids=False)
    counts = d_spans.cpu().numpy()[:, 1].astype(np.int64)
    parity = bool((counts == o1["counts"]).all())
    ids_exact = None
    if FULL_PARITY:      # bit-exact sorted multisets of every topic at full size (slow: sorts all ids on the host)
        from rmqtt_b200.engine import MatchResult
        res = MatchResult(d_spans.cpu().numpy().view(np.uint32), d_ids.cpu().numpy().view(np.uint32)[:need], d_status.cpu().numpy(), need)
        _, gids = res.canonical()
        ow = tree.match_batch(tb, to, nthreads=threads, want_ids=True)
        seg = np.repeat(np.arange(n, dtype=np.int64), np.maximum(ow["counts"], 0))
        oids = ow["ids"][np.lexsort((ow["ids"], seg))]
        ids_exact = bool(parity and len(gids) == len(oids) and (gids == oids).all())
    alg = w["bytes"] + 8 * n + 16 * w["levels"] + 16 * w["visited"] + 16 * w["probed"] + 8 * w["filters"] + 4 * w["ids"] + 8 * n
    ms = float(k.sum())
    return {"config": cfg.name + ("-zipf" if zipf else ""), "subs": cfg.n_subs, "topics": n, "gpu_ms": {"tokenize": float(k[0]), "match": float(k[1]), "deferred": float(k[2])},
            "gpu_topics_per_s": n / (ms * 1e-3), "pairs_per_s": w["ids"] / (ms * 1e-3), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
            "cpu_1thr_topics_per_s": n / o1["seconds"], f"cpu_{threads}thr_topics_per_s": n / oN["seconds"], "cpu_kind": "oracle TopicTree::matches restatement",
            "ids_per_topic": w["ids"] / n, "visited_per_topic": w["visited"] / n, "count_parity": parity, "ids_bit_exact": ids_exact, "build_s": round(build, 2)}


def retain_config(cfg, reps=10):
    rb, ro, rv = wl.gen_retained(cfg)
    fb, fo = wl.gen_retain_filters(cfg)
    n = len(fo) - 1
    eng = Engine()
    t0 = time.time(); eng.retain_bulk_load(rb, ro, rv); eng.flush(); build = time.time() - t0
    d_blob, d_offs = torch.from_numpy(fb).to(dev), torch.from_numpy(fo.view(np.int32)).to(dev)
    d_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev); d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_ids = torch.empty(1 << 20, dtype=torch.int32, device=dev)
    try:
        need = eng.retain_match_batch_device(d_blob, d_offs, d_spans, d_ids, d_status, stream)
    except Exception:
        need = None
    if need is None or need > d_ids.numel():
        import ctypes as C
        from rmqtt_b200 import _native as N
        nd = C.c_uint64(0)
        N.lib().gm_retain_match_batch_device(eng._h, d_blob.data_ptr(), d_blob.numel(), d_offs.data_ptr(), n, d_spans.data_ptr(), d_ids.data_ptr(), d_ids.numel(), C.byref(nd), d_status.data_ptr(), stream)
        need = int(nd.value)
        d_ids = torch.empty(need + 1024, dtype=torch.int32, device=dev)
    for _ in range(reps + 2):
        need = eng.retain_match_batch_device(d_blob, d_offs, d_spans, d_ids, d_status, stream)
    k = eng.kernel_ms(reps).mean(axis=0)
    tree = orc.RetainTree(); tree.bulk_insert(rb, ro, rv)
    o1 = tree.match_batch(fb, fo, nthreads=1, want_ids=False)
    oN = tree.match_batch(fb, fo, nthreads=threads, want_ids=False)
    counts = d_spans.cpu().numpy()[:, 1].astype(np.int64)
    parity = bool((counts == o1["counts"]).all())
    c = o1["counters"]
    alg = c["B"] + 8 * n + 16 * c["L"] + 32 * c["V"] + 16 * c["E"] + 4 * c["M"] + 8 * n       # SURVEY §8(d) retained form
    ms = float(k.sum())
    return {"config": cfg.name, "retained_topics": cfg.n_subs, "filters": n, "gpu_ms": {"tokenize": float(k[0]), "walk": float(k[1]), "publish": float(k[2])},
            "gpu_filters_per_s": n / (ms * 1e-3), "hits_per_s": need / (ms * 1e-3), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
            "cpu_1thr_filters_per_s": n / o1["seconds"], f"cpu_{threads}thr_filters_per_s": n / oN["seconds"], "cpu_kind": "oracle RetainTree::matches restatement",
            "hits_per_filter": need / n, "visited_per_filter": c["V"] / n, "count_parity": parity, "build_s": round(build, 2)}


FULL_PARITY = False

if __name__ == "__main__":
    if "--full-parity" in sys.argv:
        FULL_PARITY = True
        sys.argv.remove("--full-parity")
    which = sys.argv[1:] or ["C1", "C2", "C4"]
    for name in which:
        zipf = name.endswith("Z")                   # e.g. C3Z: C3 subscriptions, Zipf(1.0)-over-devices publish batch
        cfg = wl.CONFIGS[name.rstrip("Z")]
        r = retain_config(cfg) if name == "C4" else publish_config(cfg, zipf=zipf)
        print(json.dumps(r), flush=True)
