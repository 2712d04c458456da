"""CPU tier: the host-side builders (host_trie.cpp, retain_tree.cpp — trie mutation, windowed re-hash, bulk insert,
compaction, in-place retained updates) under AddressSanitizer + UndefinedBehaviorSanitizer.  No oracle here: random
operation soup, memory safety and UB only (the differential tests cover results)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "rmqtt_b200" / "csrc"


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_host_builders_under_asan_ubsan(tmp_path):
    exe = tmp_path / "san_stress"
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           f"-I{CSRC}", "-o", str(exe), str(ROOT / "tests" / "native" / "san_stress.cpp"), str(CSRC / "host_trie.cpp"), str(CSRC / "retain_tree.cpp")]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env={"ASAN_OPTIONS": "detect_leaks=1", "PATH": "/usr/bin:/bin"})
    assert run.returncode == 0, (run.stdout[-1000:], run.stderr[-3000:])
    assert run.stdout.count(" ok:") == 4
