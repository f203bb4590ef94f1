"""The engine source under AddressSanitizer + UndefinedBehaviorSanitizer (CPU debug build; tests/hostsim/sanitize_run.py).
compute-sanitizer is closed on the GPU pool: this is the memory-safety check of the indexing logic the kernels share."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_engine_source_under_asan_ubsan(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    ubsan = subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(asan) and os.path.exists(asan) and os.path.exists(ubsan)):
        pytest.skip("libasan / libubsan not available")
    lib = tmp_path / "libjsl_hostsim_asan.so"
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-DJSL_LEG_LOG", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", str(lib),
                           str(ROOT / "tests" / "hostsim" / "jsl_hostsim.cpp")])
    env = dict(os.environ, LD_PRELOAD=f"{asan}:{ubsan}", ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "hostsim" / "sanitize_run.py"), str(lib), "120"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "sanitizers ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
