"""Multi-GPU partitioning of the path: one process per GPU, the subscription set sharded by
the hash of the topic root (level 0), filters whose level 0 is '+' or '#' replicated on every shard, every
topic routed to the single shard of its own root.  The `$`-rule (trie.rs:312-318) is evaluated locally and
identically, so the union over shards of per-topic match lists equals the unsharded result.

The data path of the multi-GPU layout (device partition of a mixed batch, match of the share, all-gatherv of the
match lists — NCCL or fused into the match kernels over peer memory) lives in libgpumqtt (include/gpumqtt.h:
gm_partition_batch_device, gm_allgatherv_device, gm_gather_*).  What is here is the HOST-side placement logic (which
filters / retained topics a rank stores, which queries it answers) and a torch.distributed reference of the
all-gatherv that the CPU tests (gloo, world size 2) use to check that the union over shards equals the unsharded result.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .engine import shard_of

REPLICATED = 0xFFFFFFFF


def shard_ids(blob: np.ndarray, offs: np.ndarray, n_shards: int) -> np.ndarray:
    """Shard of every topic / filter in a packed batch (REPLICATED for root-wildcard filters): gm_shard_of_batch."""
    from . import _native as N
    n = len(offs) - 1
    out = np.empty(n, dtype=np.uint32)
    if n == 0:
        return out
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint32)
    if len(blob) == 0:
        blob = np.zeros(1, dtype=np.uint8)
    rc = N.lib().gm_shard_of_batch(blob.ctypes.data, offs.ctypes.data, n, int(n_shards), out.ctypes.data)
    if rc != 0:
        raise ValueError(f"gm_shard_of_batch failed: {rc}")
    return out


def select(blob: np.ndarray, offs: np.ndarray, idx: np.ndarray):
    """Sub-batch made of the listed entries (order kept)."""
    lens = (offs[1:] - offs[:-1]).astype(np.int64)[idx]
    new_offs = np.zeros(len(idx) + 1, dtype=np.uint32)
    np.cumsum(lens, out=new_offs[1:])
    total = int(new_offs[-1])
    src = np.repeat(offs[:-1].astype(np.int64)[idx], lens) + (np.arange(total, dtype=np.int64) - np.repeat(new_offs[:-1].astype(np.int64), lens))
    return blob[src] if total else np.zeros(0, np.uint8), new_offs


def partition_filters(blob, offs, values, rank: int, world: int):
    """Filters this rank stores: its own roots plus every root-wildcard filter."""
    sh = shard_ids(blob, offs, world)
    idx = np.nonzero((sh == rank) | (sh == REPLICATED))[0]
    b, o = select(blob, offs, idx)
    return b, o, values[idx], idx


def partition_topics(blob, offs, rank: int, world: int):
    """Topics this rank matches (publish topics never have wildcard roots that parse as such; a literal '+'
    or '#' root goes to shard 0 — any shard holds the replicated root wildcards and no literal root equals them)."""
    sh = shard_ids(blob, offs, world)
    sh = np.where(sh == REPLICATED, 0, sh)
    idx = np.nonzero(sh == rank)[0]
    b, o = select(blob, offs, idx)
    return b, o, idx


def partition_retained(blob, offs, values, rank: int, world: int):
    """Retained topics this rank stores: concrete topics sharded by their root.  A retained topic
    whose level 0 is literally '+' or '#' (tolerated by the reference, never valid MQTT) cannot be sharded — it would
    have to shadow the root-level wildcard expansion of every shard (retain.rs:313) — and is refused."""
    sh = shard_ids(blob, offs, world)
    if (sh == REPLICATED).any():
        raise ValueError("a retained topic with a literal '+' / '#' root cannot be sharded")
    idx = np.nonzero(sh == rank)[0]
    b, o = select(blob, offs, idx)
    return b, o, values[idx], idx


def partition_retain_filters(blob, offs, rank: int, world: int):
    """SUBSCRIBE filters this rank answers: a filter with a literal root goes to the shard of that root; a filter whose
    level 0 is '+' or '#' is answered by EVERY shard for its own roots (the `$`-exclusion at the root, retain.rs:327-331
    and 345-349, is local), and the complete hit list is the concatenation over ranks (all_gatherv_match_lists keeps one
    entry per (rank, filter))."""
    sh = shard_ids(blob, offs, world)
    idx = np.nonzero((sh == rank) | (sh == REPLICATED))[0]
    b, o = select(blob, offs, idx)
    return b, o, idx


def all_gatherv_match_lists(topic_index: torch.Tensor, counts: torch.Tensor, ids: torch.Tensor, group=None):
    """All-gatherv of per-rank match lists — host-side reference used by the gloo tests; the product path is
    gm_allgatherv_device / gm_match_gather_device inside the library.

    topic_index int64[k]  global index of each local topic, counts int64[k] matches per local topic,
    ids int32/uint32[sum(counts)] the match lists concatenated in local topic order.
    Returns (topic_index_all, counts_all, ids_all) concatenated over ranks in rank order, identical on every rank.
    Implementation: one all_gather of the (k, m) sizes, then one broadcast per rank into a pre-sized buffer
    (NCCL has no native all-gatherv)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = ids.device
    sizes = torch.tensor([topic_index.numel(), ids.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    ks = [int(s[0]) for s in all_sizes]
    ms = [int(s[1]) for s in all_sizes]
    ti = torch.empty(sum(ks), dtype=torch.int64, device=dev)
    ct = torch.empty(sum(ks), dtype=torch.int64, device=dev)
    ia = torch.empty(sum(ms), dtype=ids.dtype, device=dev)
    ko = np.concatenate([[0], np.cumsum(ks)])
    mo = np.concatenate([[0], np.cumsum(ms)])
    ti[ko[rank]:ko[rank + 1]] = topic_index
    ct[ko[rank]:ko[rank + 1]] = counts
    ia[mo[rank]:mo[rank + 1]] = ids
    for r in range(world):
        src = dist.get_global_rank(group, r) if group is not None else r
        if ks[r]:
            dist.broadcast(ti[ko[r]:ko[r + 1]], src=src, group=group)
            dist.broadcast(ct[ko[r]:ko[r + 1]], src=src, group=group)
        if ms[r]:
            dist.broadcast(ia[mo[r]:mo[r + 1]], src=src, group=group)
    return ti, ct, ia
