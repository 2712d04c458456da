"""CPU tier: the C-ABI library loads and exports every symbol include/gpumqtt.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from rmqtt_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    src = (ROOT / "include" / "gpumqtt.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmr?_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 18
    lib = C.CDLL(str(N.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gpumqtt.h but not exported"
        assert n in N.SIGNATURES, f"{n} has no ctypes signature in rmqtt_b200/_native.py"
    assert set(N.SIGNATURES) <= set(names)


def test_version_and_shard_fn():
    lib = N.lib()
    assert b"sm_100a" in lib.gm_version()
    s = lib.gm_shard_of(b"reg-01/x", 8, 8)
    assert 0 <= s < 8
    assert lib.gm_shard_of(b"reg-01", 6, 8) == s            # only level 0 counts
    assert lib.gm_shard_of(b"+/x", 3, 8) == 0xFFFFFFFF       # root wildcards are replicated
    assert lib.gm_shard_of(b"#", 1, 8) == 0xFFFFFFFF


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_cpu_fallback_without_a_device():
    lib = N.lib()
    h = C.c_void_p()
    assert lib.gm_create(None, C.byref(h)) == N.GM_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.gm_last_error(None)
