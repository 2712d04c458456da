"""Deterministic `iot6` workload generator — ctypes wrapper over libgmworkload.so.

Configs C1..C5 of BASELINE.json.  The same bytes feed the oracle, the GPU engine and bench.py.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
WL_PATH = PKG / "libgmworkload.so"


class _WlParams(C.Structure):
    _fields_ = [("R", C.c_uint32), ("S", C.c_uint32), ("D", C.c_uint32), ("K", C.c_uint32), ("M", C.c_uint32), ("F", C.c_uint32),
                ("p_plus", C.c_double), ("p_hash", C.c_double), ("p_root_plus", C.c_double), ("seed", C.c_uint64)]


_lib = None


def _wl():
    global _lib
    if _lib is None:
        if not WL_PATH.exists():
            from . import _build
            _build.build_workload()
        L = C.CDLL(str(WL_PATH))
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        P = C.POINTER(_WlParams)
        L.wl_space.restype, L.wl_space.argtypes = u64, [P]
        L.wl_max_len.restype, L.wl_max_len.argtypes = u32, []
        L.wl_gen_subs.restype, L.wl_gen_subs.argtypes = u64, [P, u64, u64, vp, vp]
        L.wl_gen_subs_sharded.restype, L.wl_gen_subs_sharded.argtypes = u64, [P, u64, u64, vp, vp, vp, vp, C.POINTER(u64)]
        L.wl_gen_topics.restype, L.wl_gen_topics.argtypes = u64, [P, u64, u64, C.c_double, u64, vp, u32, u64, vp, vp]
        L.wl_gen_topics_zipf.restype, L.wl_gen_topics_zipf.argtypes = u64, [P, u64, u64, u64, vp, vp]
        L.wl_gen_retained.restype, L.wl_gen_retained.argtypes = u64, [P, u64, u64, vp, vp]
        L.wl_gen_retain_filters.restype, L.wl_gen_retain_filters.argtypes = u64, [P, u64, u64, vp, vp]
        L.wl_region_name.restype, L.wl_region_name.argtypes = u32, [u32, vp]
        _lib = L
    return _lib


@dataclass(frozen=True)
class Config:
    name: str
    R: int
    S: int
    D: int
    K: int
    M: int
    F: int
    n_subs: int
    n_topics: int
    seed: int
    p_plus: float = 0.30
    p_hash: float = 0.05
    p_root_plus: float = 0.02
    frac_from_subs: float = 0.0

    def params(self) -> _WlParams:
        return _WlParams(self.R, self.S, self.D, self.K, self.M, self.F, self.p_plus, self.p_hash, self.p_root_plus, self.seed)

    @property
    def space(self) -> int:
        return self.R * self.S * self.D * self.K * self.M * self.F

    def scaled(self, n_subs=None, n_topics=None, name=None) -> "Config":
        d = dict(self.__dict__)
        if n_subs is not None:
            d["n_subs"] = int(n_subs)
        if n_topics is not None:
            d["n_topics"] = int(n_topics)
        if name:
            d["name"] = name
        return Config(**d)


# BASELINE.json configs
C1 = Config("C1", 4, 4, 8, 4, 2, 1, 1_000, 10_000, 0xC1, p_plus=0.0, p_hash=0.0, p_root_plus=0.0, frac_from_subs=0.5)
C2 = Config("C2", 16, 32, 256, 8, 4, 2, 1_000_000, 100_000, 0xC2)
C3 = Config("C3", 64, 64, 256, 8, 4, 2, 10_000_000, 1_000_000, 0xC3)
C4 = Config("C4", 64, 64, 256, 8, 4, 2, 5_000_000, 100_000, 0xC4)   # n_subs = retained topics, n_topics = SUBSCRIBE filters
CONFIGS = {c.name: c for c in (C1, C2, C3, C4)}


def _alloc(n: int):
    blob = np.empty(int(n) * int(_wl().wl_max_len()), dtype=np.uint8)
    offs = np.empty(int(n) + 1, dtype=np.uint32)
    return blob, offs


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def gen_subs(cfg: Config, n: int | None = None, first: int = 0):
    """-> (blob uint8[], offsets uint32[n+1], values uint32[n]); subscription i carries value i."""
    n = cfg.n_subs if n is None else int(n)
    blob, offs = _alloc(n)
    p = cfg.params()
    used = _wl().wl_gen_subs(C.byref(p), first, n, _vp(blob), _vp(offs))
    return blob[:used].copy(), offs, np.arange(first, first + n, dtype=np.uint32)


def gen_subs_sharded(cfg: Config, regions, n: int | None = None, first: int = 0):
    """Shard of the subscription set: filters whose region is in `regions` plus every root-'+' filter.
    -> (blob, offsets uint32[kept+1], values uint32[kept] = original subscription indices)"""
    n = cfg.n_subs if n is None else int(n)
    blob, offs = _alloc(n)
    vals = np.empty(n, dtype=np.uint32)
    keep = np.zeros(cfg.R, dtype=np.uint8)
    keep[np.asarray(list(regions), dtype=np.int64)] = 1
    p = cfg.params()
    nbytes = C.c_uint64(0)
    kept = int(_wl().wl_gen_subs_sharded(C.byref(p), first, n, _vp(keep), _vp(blob), _vp(offs), _vp(vals), C.byref(nbytes)))
    return blob[:int(nbytes.value)].copy(), offs[:kept + 1].copy(), vals[:kept].copy()


def gen_topics(cfg: Config, n: int | None = None, first: int = 0, regions=None, stream: int = 0):
    n = cfg.n_topics if n is None else int(n)
    blob, offs = _alloc(n)
    p = cfg.params()
    reg = np.ascontiguousarray(regions, dtype=np.uint32) if regions is not None and len(regions) else None
    used = _wl().wl_gen_topics(C.byref(p), first, n, cfg.frac_from_subs, cfg.n_subs if cfg.frac_from_subs > 0 else 0,
                               _vp(reg) if reg is not None else None, len(reg) if reg is not None else 0, stream, _vp(blob), _vp(offs))
    return blob[:used].copy(), offs


def gen_topics_zipf(cfg: Config, n: int | None = None, first: int = 0, stream: int = 0):
    """Publish topics whose DEVICE follows a Zipf(1.0) popularity (secondary workload)."""
    n = cfg.n_topics if n is None else int(n)
    blob, offs = _alloc(n)
    p = cfg.params()
    used = _wl().wl_gen_topics_zipf(C.byref(p), first, n, stream, _vp(blob), _vp(offs))
    return blob[:used].copy(), offs


def gen_retained(cfg: Config, n: int | None = None, first: int = 0):
    n = cfg.n_subs if n is None else int(n)
    blob, offs = _alloc(n)
    p = cfg.params()
    used = _wl().wl_gen_retained(C.byref(p), first, n, _vp(blob), _vp(offs))
    return blob[:used].copy(), offs, np.arange(first, first + n, dtype=np.uint32)


def gen_retain_filters(cfg: Config, n: int | None = None, first: int = 0):
    n = cfg.n_topics if n is None else int(n)
    blob, offs = _alloc(n)
    p = cfg.params()
    used = _wl().wl_gen_retain_filters(C.byref(p), first, n, _vp(blob), _vp(offs))
    return blob[:used].copy(), offs


def region_name(r: int) -> bytes:
    buf = (C.c_char * 16)()
    n = _wl().wl_region_name(r, buf)
    return bytes(buf[:n])


def unpack(blob: np.ndarray, offs: np.ndarray):
    b = blob.tobytes()
    return [b[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
