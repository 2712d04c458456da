"""ctypes wrapper over oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
The product package (rmqtt_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("V", "E", "F", "M", "L", "B")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        build()
    L = C.CDLL(str(_LIB_PATH))
    vp, cp, u32, u64, i64, i32 = C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint64, C.c_int64, C.c_int32
    P = C.POINTER
    sig = {
        "orc_topic_parse": (i64, [cp, u32, P(C.c_uint8), u32]),
        "orc_matches_str": (i32, [cp, u32, cp, u32]),
        "orc_tree_new": (vp, []),
        "orc_tree_free": (None, [vp]),
        "orc_tree_insert": (i32, [vp, cp, u32, u64]),
        "orc_tree_remove": (i32, [vp, cp, u32, u64]),
        "orc_tree_values_size": (u64, [vp]),
        "orc_tree_nodes_size": (u64, [vp]),
        "orc_tree_bulk_insert": (i64, [vp, vp, vp, vp, u64, C.c_int]),
        "orc_tree_match": (i64, [vp, cp, u32, P(u64), u64, P(Counters)]),
        "orc_tree_match_batch": (C.c_double, [vp, vp, vp, u64, C.c_int, vp, vp, vp, P(Counters)]),
        "orc_router_new": (vp, []),
        "orc_router_free": (None, [vp]),
        "orc_router_add": (i32, [vp, cp, u32, cp, u32, u32, u64]),
        "orc_router_remove": (i32, [vp, cp, u32, cp, u32, u64]),
        "orc_router_add_full": (i32, [vp, cp, u32, cp, u32, u32, u64, u64, i32, i32, u32, cp, u32]),
        "orc_router_match_full": (i64, [vp, cp, u32, u64, cp, u32, u64, vp, u64]),
        "orc_router_topics": (i64, [vp]),
        "orc_router_routes": (i64, [vp]),
        "orc_router_topics_tree": (u64, [vp]),
        "orc_router_bulk_add": (i64, [vp, vp, vp, vp, u64, C.c_int]),
        "orc_router_match": (i64, [vp, cp, u32, P(u32), u64, P(Counters)]),
        "orc_router_match_batch": (C.c_double, [vp, vp, vp, u64, C.c_int, vp, P(u64), P(Counters)]),
        "orc_retain_new": (vp, []),
        "orc_retain_free": (None, [vp]),
        "orc_retain_insert": (i32, [vp, cp, u32, i64]),
        "orc_retain_remove": (i32, [vp, cp, u32, P(i64)]),
        "orc_retain_values_size": (u64, [vp]),
        "orc_retain_nodes_size": (u64, [vp]),
        "orc_retain_bulk_insert": (i64, [vp, vp, vp, vp, u64]),
        "orc_retain_match": (i64, [vp, cp, u32, P(i64), u64, P(Counters)]),
        "orc_retain_match_batch": (C.c_double, [vp, vp, vp, u64, C.c_int, vp, vp, vp, P(Counters)]),
        "orc_hardware_threads": (i32, []),
        "orc_numa_interleave": (i32, [i32]),
        "orc_router_readers": (i64, [vp, cp, u32, i32, vp, u64]),
        "orc_router_churn": (C.c_double, [vp, vp, vp, vp, u64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else s.encode("utf-8")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pack(strings):
    """list[str|bytes] -> (blob uint8[], offsets uint32[n+1])"""
    bs = [_b(s) for s in strings]
    offs = np.zeros(len(bs) + 1, dtype=np.uint32)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64).astype(np.uint32)
    blob = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return blob, offs


KINDS = ("Normal", "Metadata", "Blank", "Single", "Multi")


def topic_parse(s):
    """Returns list of level kind names, or None if Topic::from_str would be Err."""
    b = _b(s)
    kinds = (C.c_uint8 * 256)()
    n = lib().orc_topic_parse(b, len(b), kinds, 256)
    if n < 0:
        return None
    return [KINDS[kinds[i]] for i in range(min(n, 256))]


def matches_str(filt, topic):
    f, t = _b(filt), _b(topic)
    r = lib().orc_matches_str(f, len(f), t, len(t))
    if r < 0:
        raise ValueError("invalid filter")
    return bool(r)


def hardware_threads() -> int:
    """CPUs this process may run on (affinity mask / cpuset) — what the multi-threaded batch calls should use."""
    return max(1, int(lib().orc_hardware_threads()))


def numa_interleave(on: bool = True) -> bool:
    """Interleave this thread's (and its future threads') new pages over all NUMA nodes; call before building a tree."""
    return int(lib().orc_numa_interleave(1 if on else 0)) == 0


class TopicTree:
    """Restatement of rmqtt::trie::TopicTree<u64> (rmqtt/src/trie.rs)."""

    def __init__(self):
        self._h = lib().orc_tree_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tree_free(self._h)
            self._h = None

    def insert(self, filt, value: int) -> bool:
        b = _b(filt)
        r = lib().orc_tree_insert(self._h, b, len(b), value)
        if r < 0:
            raise ValueError(f"invalid topic filter {filt!r}")
        return bool(r)

    def remove(self, filt, value: int) -> bool:
        b = _b(filt)
        r = lib().orc_tree_remove(self._h, b, len(b), value)
        if r < 0:
            raise ValueError(f"invalid topic filter {filt!r}")
        return bool(r)

    def bulk_insert(self, blob: np.ndarray, offs: np.ndarray, vals: np.ndarray, nthreads: int = 1) -> int:
        assert offs.dtype == np.uint32 and vals.dtype == np.uint32 and blob.dtype == np.uint8
        return int(lib().orc_tree_bulk_insert(self._h, _ptr(blob), _ptr(offs), _ptr(vals), len(vals), nthreads))

    def values_size(self) -> int:
        return int(lib().orc_tree_values_size(self._h))

    def nodes_size(self) -> int:
        return int(lib().orc_tree_nodes_size(self._h))

    def matches(self, topic, with_counters=False):
        """Sorted list of matched values (multiset), or None if the topic is invalid."""
        b = _b(topic)
        ctr = Counters()
        cap = 1 << 12
        while True:
            out = (C.c_uint64 * cap)()
            n = lib().orc_tree_match(self._h, b, len(b), out, cap, C.byref(ctr))
            if n < 0:
                return (None, None) if with_counters else None
            if n <= cap:
                res = sorted(out[i] for i in range(n))
                return (res, ctr.as_dict()) if with_counters else res
            cap = int(n)

    def match_batch(self, blob: np.ndarray, offs: np.ndarray, nthreads: int = 1, want_ids: bool = True):
        """Returns dict(counts int64[n] (-1 invalid), offsets uint64[n+1], ids uint32[], counters, seconds)."""
        n = len(offs) - 1
        counts = np.zeros(n, dtype=np.int64)
        ctr = Counters()
        dt = lib().orc_tree_match_batch(self._h, _ptr(blob), _ptr(offs), n, nthreads, _ptr(counts), None, None, C.byref(ctr))
        res = {"counts": counts, "counters": ctr.as_dict(), "seconds": dt}
        if want_ids:
            o = np.zeros(n + 1, dtype=np.uint64)
            np.cumsum(np.maximum(counts, 0), out=o[1:])
            ids = np.zeros(int(o[-1]), dtype=np.uint32)
            lib().orc_tree_match_batch(self._h, _ptr(blob), _ptr(offs), n, nthreads, None, _ptr(o), _ptr(ids), None)
            res["offsets"], res["ids"] = o, ids
        return res


class Router:
    """Restatement of rmqtt::router::DefaultRouter add/remove/_matches (rmqtt/src/router.rs)."""

    def __init__(self):
        self._h = lib().orc_router_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_router_free(self._h)
            self._h = None

    def add(self, filt, client, rel_id: int, id_tag: int = 0) -> bool:
        f, c = _b(filt), _b(client)
        return lib().orc_router_add(self._h, f, len(f), c, len(c), rel_id, id_tag) > 0

    def remove(self, filt, client, id_tag: int = 0) -> int:
        f, c = _b(filt), _b(client)
        return int(lib().orc_router_remove(self._h, f, len(f), c, len(c), id_tag))

    def add_full(self, filt, client, rel_id: int, id_tag: int, node_id: int, is_v5=False, no_local=False, sub_id=0, group="") -> bool:
        f, c, g = _b(filt), _b(client), _b(group)
        return lib().orc_router_add_full(self._h, f, len(f), c, len(c), rel_id, id_tag, node_id, int(is_v5), int(no_local), sub_id, g, len(g)) > 0

    def matches_full(self, topic, pub_node=0, pub_client="", pub_tag=0):
        """Canonical sorted text lines of DefaultRouter::_matches (see oracle.cpp Router::matches_full); None if invalid."""
        t, pc = _b(topic), _b(pub_client)
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            n = lib().orc_router_match_full(self._h, t, len(t), pub_node, pc, len(pc), pub_tag, buf, cap)
            if n < 0:
                return None
            if n <= cap:
                txt = buf.raw[:n].decode()
                return txt.split("\n") if txt else []
            cap = int(n)

    def bulk_add(self, blob, offs, vals, nthreads: int = 1) -> int:
        return int(lib().orc_router_bulk_add(self._h, _ptr(blob), _ptr(offs), _ptr(vals), len(vals), nthreads))

    def topics(self):
        return int(lib().orc_router_topics(self._h))

    def routes(self):
        return int(lib().orc_router_routes(self._h))

    def topics_tree(self):
        return int(lib().orc_router_topics_tree(self._h))

    def matches(self, topic):
        b = _b(topic)
        cap = 1 << 12
        while True:
            out = (C.c_uint32 * cap)()
            n = lib().orc_router_match(self._h, b, len(b), out, cap, None)
            if n < 0:
                return None
            if n <= cap:
                return sorted(out[i] for i in range(n))
            cap = int(n)

    def readers(self, topic, kind: int):
        """Secondary readers, canonical sorted lines: kind 0 _has_matches (["1"] or []), 1 _get_routes (unique matched filters),
        2 get ("node|filter").  None for an invalid topic."""
        t = _b(topic)
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            n = lib().orc_router_readers(self._h, t, len(t), kind, buf, cap)
            if n < 0:
                return None
            if n <= cap:
                return buf.raw[:n].decode().split("\n") if n else []
            cap = int(n)

    def churn(self, blob, offs, vals) -> float:
        """remove + re-add every listed subscription (single thread, the write-lock path); returns seconds."""
        return float(lib().orc_router_churn(self._h, _ptr(blob), _ptr(offs), _ptr(vals), len(vals)))

    def match_batch(self, blob, offs, nthreads: int = 1):
        n = len(offs) - 1
        counts = np.zeros(n, dtype=np.int64)
        ctr = Counters()
        tot = C.c_uint64(0)
        dt = lib().orc_router_match_batch(self._h, _ptr(blob), _ptr(offs), n, nthreads, _ptr(counts), C.byref(tot), C.byref(ctr))
        return {"counts": counts, "total_ids": int(tot.value), "counters": ctr.as_dict(), "seconds": dt}


class RetainTree:
    """Restatement of rmqtt::retain::RetainTree<i64> (rmqtt/src/retain.rs)."""

    def __init__(self):
        self._h = lib().orc_retain_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_retain_free(self._h)
            self._h = None

    def insert(self, topic, value: int):
        b = _b(topic)
        if lib().orc_retain_insert(self._h, b, len(b), value) < 0:
            raise ValueError(f"invalid topic {topic!r}")

    def remove(self, topic):
        b = _b(topic)
        old = C.c_int64(0)
        r = lib().orc_retain_remove(self._h, b, len(b), C.byref(old))
        if r < 0:
            raise ValueError(f"invalid topic {topic!r}")
        return int(old.value) if r else None

    def bulk_insert(self, blob, offs, vals) -> int:
        return int(lib().orc_retain_bulk_insert(self._h, _ptr(blob), _ptr(offs), _ptr(vals), len(vals)))

    def values_size(self):
        return int(lib().orc_retain_values_size(self._h))

    def nodes_size(self):
        return int(lib().orc_retain_nodes_size(self._h))

    def matches(self, filt, with_counters=False):
        b = _b(filt)
        ctr = Counters()
        cap = 1 << 12
        while True:
            out = (C.c_int64 * cap)()
            n = lib().orc_retain_match(self._h, b, len(b), out, cap, C.byref(ctr))
            if n < 0:
                return (None, None) if with_counters else None
            if n <= cap:
                res = sorted(out[i] for i in range(n))
                return (res, ctr.as_dict()) if with_counters else res
            cap = int(n)

    def match_batch(self, blob, offs, nthreads: int = 1, want_ids: bool = True):
        n = len(offs) - 1
        counts = np.zeros(n, dtype=np.int64)
        ctr = Counters()
        dt = lib().orc_retain_match_batch(self._h, _ptr(blob), _ptr(offs), n, nthreads, _ptr(counts), None, None, C.byref(ctr))
        res = {"counts": counts, "counters": ctr.as_dict(), "seconds": dt}
        if want_ids:
            o = np.zeros(n + 1, dtype=np.uint64)
            np.cumsum(np.maximum(counts, 0), out=o[1:])
            ids = np.zeros(int(o[-1]), dtype=np.uint32)
            lib().orc_retain_match_batch(self._h, _ptr(blob), _ptr(offs), n, nthreads, None, _ptr(o), _ptr(ids), None)
            res["offsets"], res["ids"] = o, ids
        return res
