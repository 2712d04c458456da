#!/usr/bin/env python
"""One retained lookup of config C4 (for ncu captures / the stats knob): python tools/c4_once.py [stats]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from rmqtt_b200 import workload as wl        # noqa: E402
from rmqtt_b200.engine import Engine         # noqa: E402

dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
cfg = wl.C4
rb, ro, rv = wl.gen_retained(cfg)
fb, fo = wl.gen_retain_filters(cfg)
n = len(fo) - 1
eng = Engine()
eng.retain_bulk_load(rb, ro, rv)
eng.flush()
d_blob, d_offs = torch.from_numpy(fb).to(dev), torch.from_numpy(fo.view(np.int32)).to(dev)
d_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev)
d_status = torch.zeros(n, dtype=torch.int32, device=dev)
d_ids = torch.empty(8 << 20, dtype=torch.int32, device=dev)
eng.retain_match_batch_device(d_blob, d_offs, d_spans, d_ids, d_status, stream)      # sizes the scratch queues
if "stats" in sys.argv:
    eng.debug_knob("retain_stats", 1)
hits = eng.retain_match_batch_device(d_blob, d_offs, d_spans, d_ids, d_status, stream)
torch.cuda.synchronize()
print("hits", hits, "kernel ms", eng.kernel_ms(1))
# the heaviest filters
cnt = d_spans.cpu().numpy()[:, 1]
top = np.argsort(cnt)[-5:]
txt = wl.unpack(fb, fo)
print("top filters by hits:", [(txt[i].decode(), int(cnt[i])) for i in top])
