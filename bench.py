#!/usr/bin/env python
"""Benchmark of the north-star path: Router::matches for a batch of PUBLISH topics at 10 M subscriptions.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]

One "step" = one pass of the hot path (tokenise -> trie walk -> per-topic match lists) over one batch of
synthetic topics (workload C3 of BASELINE.json: 10 M subscriptions, 30 % '+', 5 % '#', 6-level IoT topics,
1 M-topic batch).  N > 1 is launched by torchrun, one process per GPU: the subscription set is sharded by
topic-root hash (root-wildcard filters replicated), every rank matches its own batch of topics that belong to
its shard, there is no data-path collective -> "scaling": "weak".

Prints ONE JSON line (rank 0).  See DESIGN.md §"Measurement" for every key.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--subs", type=int, default=None, help="override subscription count (non-default => not the headline config)")
    ap.add_argument("--topics", type=int, default=None, help="override topics per batch")
    ap.add_argument("--batches", type=int, default=4, help="distinct topic batches rotated through the timed loop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    return ap.parse_args()


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _regions_of_rank(cfg, rank: int, world: int):
    from rmqtt_b200 import workload as wl
    from rmqtt_b200.engine import shard_of
    if world == 1:
        return list(range(cfg.R))
    return [r for r in range(cfg.R) if shard_of(wl.region_name(r), world) == rank]


def _cfg(args):
    from rmqtt_b200 import workload as wl
    cfg = wl.C3
    if args.subs or args.topics:
        cfg = cfg.scaled(n_subs=args.subs, n_topics=args.topics, name="C3-scaled")
    return cfg


def _workload_desc(cfg, world):
    return (f"{cfg.name}: {cfg.n_subs} subscriptions (30% '+', 5% '#', 0.6% root '+'), 6-level IoT topics "
            f"reg/site/dev/sen/met/ch over R{cfg.R}xS{cfg.S}xD{cfg.D}xK{cfg.K}xM{cfg.M}xF{cfg.F}, "
            f"{cfg.n_topics}-topic uniform batch per GPU, seed {cfg.seed:#x}"
            + (f", subscriptions sharded by topic-root hash over {world} GPUs (root-wildcards replicated)" if world > 1 else ""))


# ======================================================================================================
def run_reference(args):
    """The reference's own CPU implementation of the path.  rmqtt is Rust and cannot be built here (no
    cargo/rustc), so this arm times the C++ restatement of DefaultRouter::_matches (oracle/oracle.cpp) on
    all host cores against the SAME subscription set; each step matches a bounded sample of the batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from rmqtt_b200 import workload as wl
    cfg = _cfg(args)
    threads = orc.hardware_threads()
    sb, so, sv = wl.gen_subs(cfg)
    router = orc.Router()
    t0 = time.time()
    router.bulk_add(sb, so, sv, nthreads=min(threads, 64))
    build_s = time.time() - t0
    # bounded sample per step: sized so that warmup + steps stay within a few minutes on the host cores
    sample = min(cfg.n_topics, max(2_000, 25_000_000 // max(1, args.steps + args.warmup)))
    tb, to = wl.gen_topics(cfg, sample)
    for _ in range(max(1, args.warmup)):
        router.match_batch(tb, to, nthreads=threads)
    secs, ids = 0.0, 0
    for _ in range(args.steps):
        r = router.match_batch(tb, to, nthreads=threads)
        secs += r["seconds"]
        ids = r["total_ids"]
    value = sample * args.steps / secs
    line = {
        "impl": "reference", "metric": "topic-matches/sec @10M subs", "value": value, "unit": "topics/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": _workload_desc(cfg, 1), "sample": f"{sample} topics of the batch per step", "oracle_build_s": round(build_s, 1),
                   "matched_ids_per_topic": ids / sample},
        "cpu_baseline": {"value": value, "unit": "topics/s", "cores": threads, "kind": "port",
                         "sample": f"{sample}-topic sample x {args.steps} steps, C++ restatement of DefaultRouter::_matches (Rust reference not buildable here: no cargo)"},
        "e2e": {"value": value, "unit": "topics/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ======================================================================================================
def run_own(args):
    import torch
    import torch.distributed as dist
    from rmqtt_b200 import workload as wl
    from rmqtt_b200 import _native as N
    from rmqtt_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = _cfg(args)
    regions = _regions_of_rank(cfg, rank, world)

    # ---- build the device-resident trie for this rank's shard -------------------------------------
    t0 = time.time()
    if world == 1:
        sb, so, sv = wl.gen_subs(cfg)
    else:
        sb, so, sv = wl.gen_subs_sharded(cfg, regions)
    eng = Engine(device=local, filters_hint=len(sv))
    eng.bulk_load(sb, so, sv)
    eng.flush()
    build_s = time.time() - t0
    st = eng.stats()
    del sb, so

    # ---- topic batches, resident in HBM before the timed region ---------------------------------
    n = cfg.n_topics
    B = max(1, args.batches)
    host_batches = [wl.gen_topics(cfg, n, regions=regions if world > 1 else None, stream=rank * 1000 + b) for b in range(B)]
    d_batches = [(torch.from_numpy(tb).to(dev), torch.from_numpy(to.view(np.int32)).to(dev)) for tb, to in host_batches]
    d_spans = torch.zeros((n, 2), dtype=torch.int32, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_needed = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    cap = 64 * n
    d_ids = torch.empty(cap, dtype=torch.int32, device=dev)
    needed_max, works = 0, []
    for tb, to in d_batches:   # untimed instrumented pass: exact work counters + output sizing
        while True:
            w = eng.match_batch_device(tb, to, d_spans, d_ids, d_needed, d_status, stream, work=True)
            need = int(d_needed.item())
            if need <= d_ids.numel():
                break
            d_ids = torch.empty(int(need * 1.1) + 1024, dtype=torch.int32, device=dev)
        works.append(w)
        needed_max = max(needed_max, need)
    if d_ids.numel() > 2 * needed_max + 1024:
        d_ids = torch.empty(int(needed_max * 1.25) + 1024, dtype=torch.int32, device=dev)

    def step(k):
        tb, to = d_batches[k % B]
        eng.match_batch_device(tb, to, d_spans, d_ids, d_needed, d_status, stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for k in range(max(3, args.warmup)):
        step(k)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.2)
    launches0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.kernel_launches() - launches0
    kms = eng.kernel_ms(min(64, args.steps))
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * n * args.steps / (ms / 1e3)

    # ---- end to end through the C ABI with (pinned) HOST buffers: H2D + kernels + D2H every step ---
    lib = N.lib()
    hb, ho = host_batches[0]
    ids_cap = int(needed_max * 1.25) + 1024

    def pinned(nbytes, dtype, shape):
        p = lib.gm_host_alloc(nbytes)
        assert p, "gm_host_alloc failed"
        return p, np.frombuffer((C.c_uint8 * nbytes).from_address(p), dtype=dtype).reshape(shape)

    p_blob, a_blob = pinned(len(hb), np.uint8, (len(hb),))
    p_offs, a_offs = pinned(4 * (n + 1), np.uint32, (n + 1,))
    p_spans, a_spans = pinned(8 * n, np.uint32, (n, 2))
    p_ids, a_ids = pinned(4 * ids_cap, np.uint32, (ids_cap,))
    p_status, a_status = pinned(4 * n, np.int32, (n,))
    a_blob[:] = hb
    a_offs[:] = ho
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))
    need = C.c_uint64(0)

    def e2e_step():
        rc = lib.gm_match_batch(eng._h, p_blob, p_offs, n, p_spans, p_ids, ids_cap, C.byref(need), p_status)
        assert rc == 0, lib.gm_last_error(eng._h)

    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        ts = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_s = float(ts.item())
    e2e_value = world * n * e2e_steps / e2e_s
    h2d = int(len(hb) + 4 * (n + 1))
    d2h = int(8 + 4 * n + 8 * n + 4 * need.value)
    for p in (p_blob, p_offs, p_spans, p_ids, p_status):
        lib.gm_host_free(p)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_match_fast) -------------------------------------------
    peak, peak_src = _peaks()
    W = {k: sum(w[k] for w in works) / len(works) for k in works[0] if not isinstance(works[0][k], list)}   # mean per launch over the rotated batches
    diag = {"probes_by_depth": works[0]["probes_by_depth"], "misses_by_depth": works[0]["misses_by_depth"], "slot_loads": works[0]["slot_loads"]}
    k2_bytes = 16 * W["visited"] + 16 * W["probed"] + 8 * W["filters"] + 4 * W["ids"] + 8 * n     # SURVEY §8(d), walk terms
    k1_bytes = W["bytes"] + 8 * n + 16 * W["levels"]                                              # SURVEY §8(d), tokeniser terms
    k_mean = kms.mean(axis=0) if len(kms) else np.zeros(3)
    k2_ms = float(k_mean[1])
    achieved = k2_bytes / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else None
    traffic = None
    prof = ROOT / "profiles" / "k_match_fast_traffic.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            pass
    if traffic is None:
        traffic = 771856128   # dram__bytes_read.sum + dram__bytes_write.sum of k_match_fast, one C3 launch, profiles/r1_k2_windows.ncu-rep
    roofline = {"bound": "hbm", "kernel": "k_match_fast", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k2_bytes, "kernel_ms": {"k_tokenize+k_bucket_scan+k_bucket_scatter": float(k_mean[0]), "k_match_fast": k2_ms, "k_match_slow": float(k_mean[2])},
                "pipeline": {"algorithmic_bytes_per_step": k1_bytes + k2_bytes,
                             "achieved": (k1_bytes + k2_bytes) / (float(k_mean.sum()) * 1e-3) / 1e9 if k_mean.sum() > 0 else None}}

    # ---- CPU baseline: the oracle's DefaultRouter::_matches restatement on the host cores -----------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        threads = orc.hardware_threads()
        sb, so, sv2 = wl.gen_subs(cfg)
        router = orc.Router()
        router.bulk_add(sb, so, sv2, nthreads=min(threads, 64))
        sample = min(n, 250_000)
        stb, sto = wl.gen_topics(cfg, sample)
        router.match_batch(stb, sto, nthreads=threads)
        reps, secs = 0, 0.0
        while secs < 4.0 and reps < 50:
            secs += router.match_batch(stb, sto, nthreads=threads)["seconds"]
            reps += 1
        one = router.match_batch(stb[:int(sto[20000])], sto[:20001], nthreads=1)
        cpu = {"value": sample * reps / secs, "unit": "topics/s", "cores": threads, "kind": "port",
               "sample": f"{sample}-topic sample x {reps} reps of the same workload; C++ restatement of DefaultRouter::_matches "
                         f"(oracle/oracle.cpp; the Rust reference cannot be built here: no cargo)",
               "single_thread_value": 20000 / one["seconds"]}

    line = {
        "metric": "topic-matches/sec @10M subs", "value": value, "unit": "topics/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": _workload_desc(cfg, world), "l2": f"device tables {st['device_bytes'] / 1e9:.2f} GB >> 126 MB L2; {B} distinct topic batches rotated",
                   "matched_ids_per_topic": W["ids"] / n, "visited_nodes_per_topic": W["visited"] / n, "deferred_topics_per_batch": W["deferred"], "probe_diag": diag,
                   "trie": {k: st[k] for k in ("values", "nodes", "edges", "edge_slots", "dict_entries", "plus_nodes", "device_bytes", "max_depth")},
                   "build_s": round(build_s, 1), "e2e_timing": "perf_counter around synchronous gm_match_batch calls (pinned host buffers)"},
        "e2e": {"value": e2e_value, "unit": "topics/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = sys.stdout


def _emit(line: dict) -> None:
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


if __name__ == "__main__":
    a = _args()
    # stdout carries exactly ONE line (the JSON result): keep a private handle to it and point fd 1 at stderr so that
    # library chatter written by native code (e.g. NCCL's "NCCL version ..." banner) cannot land on it.
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_own(a)
