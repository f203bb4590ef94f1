"""Runs the engine SOURCE (CPU debug build) under AddressSanitizer + UndefinedBehaviorSanitizer: differential fuzz of
every specialisation level incl. wide instances, plus golden replays of every observation kind.  compute-sanitizer is
closed on the GPU pool, so this is the memory-safety net for the indexing logic the kernels share with this build.

    LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
        python tests/hostsim/sanitize_run.py <path of the sanitizer build> [n_fuzz]

(tests/test_engine_sanitizers.py builds the library and launches this.)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from hostsim import hostsim_py as hs  # noqa: E402

hs.LIB = Path(sys.argv[1])
hs.build = lambda force=False: hs.LIB
import _golden as G  # noqa: E402
import fuzz_engine as F  # noqa: E402
import test_engine_hostsim as T  # noqa: E402

n_fuzz = int(sys.argv[2]) if len(sys.argv) > 2 else 150
rng = np.random.default_rng(5)
steps = 0
for i in range(n_fuzz):
    sub = np.random.default_rng(rng.integers(0, 2**63))
    steps += F.run_one(sub, simple=(None, 3, 1, 2, 4)[i % 5])[0]
for i in range(6):
    sub = np.random.default_rng(rng.integers(0, 2**63))
    steps += F.run_one(sub, max_steps=200, simple=(3, None, 4)[i % 3], jrange=((33, 64), (65, 101))[(i // 3) % 2])[0]
replays = 0
for group in ("classic", "transport", "buffers", "features", "oparray", "tassel", "truncation"):
    for name in G.names(group)[::9]:
        sc = G.Scenario(group, name)
        if sc.lowered.is_deterministic:
            T.replay(sc)
            replays += 1
print(f"sanitizers ok: {steps} fuzz steps, {replays} golden replays")
