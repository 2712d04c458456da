#!/usr/bin/env python
"""A/B harness (round 1): edge-table windows (layout.h) x tile-chunk scheduling of k_match_fast on workload C3.
Builds the 10 M-subscription trie once per window setting, then times the match kernels (engine timing ring) for
every tile_chunk; match counts must be identical across all settings (the oracle check itself is bench_configs.py).
usage: ab_windows.py [windows_log2 ...]   e.g.  ab_windows.py 0 8"""
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rmqtt_b200 import workload as wl       # noqa: E402
from rmqtt_b200.engine import Engine        # noqa: E402

dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
cfg = wl.CONFIGS[os.environ.get("AB_CONFIG", "C3")]
chunks = [int(x) for x in os.environ.get("AB_CHUNKS", "1,4,16,64").split(",")]
knob_sets = [dict(kv.split("=") for kv in ks.split("+") if kv) for ks in os.environ.get("AB_KNOBS", "").split(",")]   # e.g. "prefetch_values=0,prefetch_values=1"
reps = int(os.environ.get("AB_REPS", "30"))
sb, so, sv = wl.gen_subs(cfg)
batches = [wl.gen_topics(cfg, stream=k) for k in range(3)]      # distinct batches, rotated (device tables >> L2 anyway)
ref_digest = None
for arg in (sys.argv[1:] or ["0", "8"]):          # "W" or "W:S" = windows_log2 [: edge slots per filter]
    wlog = int(arg.split(":")[0])
    os.environ["GM_EDGE_WINDOWS_LOG2"] = str(wlog)
    if ":" in arg:
        os.environ["GM_EDGE_SLOTS_PER_FILTER"] = arg.split(":")[1]
    eng = Engine(filters_hint=len(sv))
    t0 = time.time(); eng.bulk_load(sb, so, sv); eng.flush(); build = time.time() - t0
    dbs = []
    for tb, to in batches:
        n = len(to) - 1
        dbs.append((torch.from_numpy(tb).to(dev), torch.from_numpy(to.view(np.int32)).to(dev), n))
    n = dbs[0][2]
    d_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev); d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_needed = torch.zeros(1, dtype=torch.int64, device=dev)
    d_ids = torch.empty(64 * n + 1024, dtype=torch.int32, device=dev)
    for chunk, knobs in [(c, k) for c in chunks for k in knob_sets]:
        eng.debug_knob("tile_chunk", chunk)
        for kn, kv in knobs.items():
            eng.debug_knob(kn, int(kv))
        for i in range(reps + 3):
            b = dbs[i % len(dbs)]
            eng.match_batch_device(b[0], b[1], d_spans, d_ids, d_needed, d_status, stream)
        torch.cuda.synchronize()
        k = eng.kernel_ms(reps)
        b = dbs[0]
        eng.match_batch_device(b[0], b[1], d_spans, d_ids, d_needed, d_status, stream)
        torch.cuda.synchronize()
        digest = hashlib.sha256(d_spans.cpu().numpy()[:, 1].tobytes()).hexdigest()[:16] + ":" + str(int(d_needed.item()))
        if ref_digest is None:
            ref_digest = digest
        print(json.dumps({"windows_log2": wlog, "arg": arg, "tile_chunk": chunk, "knobs": knobs, "k_tok_sort_ms": round(float(k[:, 0].mean()), 4), "k_match_fast_ms": round(float(k[:, 1].mean()), 4),
                          "k_match_fast_min_ms": round(float(k[:, 1].min()), 4), "counts_same": digest == ref_digest, "build_s": round(build, 1), "stats": {kk: vv for kk, vv in eng.stats().items() if "edge" in kk or "bytes" in kk}}), flush=True)
    del eng
    torch.cuda.empty_cache()
