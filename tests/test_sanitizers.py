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


def _build(tmp_path, name, san_flags):
    exe = tmp_path / name
    cmd = ["g++", "-O1", "-g", "-std=c++17", *san_flags, "-fno-omit-frame-pointer", "-pthread",
           f"-I{CSRC}", "-o", str(exe), str(ROOT / "tests" / "native" / "san_stress.cpp"), str(CSRC / "host_trie.cpp"), str(CSRC / "retain_tree.cpp")]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_parallel_bulk_paths_under_sanitizers(tmp_path, san):
    """The all-host-threads bulk paths (insert_batch_parallel + parallel first flush, the retained tree's level-by-level build)
    forced onto small inputs with 6 worker threads: memory safety / UB under ASan + UBSan, data races under ThreadSanitizer."""
    flags = [f"-fsanitize={san}"] + (["-fno-sanitize-recover=undefined"] if "undefined" in san else [])
    exe = _build(tmp_path, "san_par_" + san.split(",")[0], flags)
    env = {"PATH": "/usr/bin:/bin", "ASAN_OPTIONS": "detect_leaks=1", "TSAN_OPTIONS": "halt_on_error=1 second_deadlock_stack=1"}
    run = subprocess.run([str(exe), "par"], capture_output=True, text=True, timeout=900, env=env)
    assert run.returncode == 0, (run.stdout[-1000:], run.stderr[-4000:])
    assert run.stdout.count(" ok:") == 3 and "WARNING: ThreadSanitizer" not in run.stderr


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_batcher_front_end_under_sanitizers(tmp_path, san):
    """rmqtt_b200/csrc/batcher.cpp (gm_submit / dispatchers / callbacks / drain / destroy) against a stubbed engine: several
    producer threads, every cookie answered exactly once with its own topic's ids, the capacity-retry and failing-batch paths."""
    exe = tmp_path / ("batcher_" + san.split(",")[0])
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread", f"-fsanitize={san}", "-fno-omit-frame-pointer", "-o", str(exe),
           str(ROOT / "tests" / "native" / "batcher_stress.cpp"), str(CSRC / "batcher.cpp")]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    env = {"PATH": "/usr/bin:/bin", "ASAN_OPTIONS": "detect_leaks=1", "TSAN_OPTIONS": "halt_on_error=1"}
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, (run.stdout[-1500:], run.stderr[-3000:])
    assert run.stdout.count("-> ok:") == 4 and "WARNING: ThreadSanitizer" not in run.stderr
