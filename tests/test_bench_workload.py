"""bench.py's synthetic workload: vectorised Taillard bounds == the compiler's (reference utils.py:310-352),
instance slices are shard-invariant, and the CPU reference arm prints a valid line."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from jobshoplab_b200 import compiler  # noqa: E402


def test_bounds_match_compiler():
    om, od = bench.make_instances(5)
    lb, mat = bench.taillard_bounds(om, od)
    for i in range(5):
        jos = np.arange(0, bench.J * bench.M + 1, bench.M)
        assert compiler._taillard_lower_bound(jos, om[i], od[i]) == lb[i]
        assert od[i].sum() == mat[i]
        assert sorted(om[i, :bench.M].tolist()) == list(range(bench.M))
    assert od.min() >= 1 and od.max() <= 99


def test_instances_are_shard_invariant():
    a_m, a_d = bench.make_instances(bench.BLOCK + 300)
    b_m, b_d = bench.make_instances(200, first=bench.BLOCK - 50)
    assert np.array_equal(a_m[bench.BLOCK - 50:bench.BLOCK + 150], b_m)
    assert np.array_equal(a_d[bench.BLOCK - 50:bench.BLOCK + 150], b_d)
    c_m, _ = bench.make_instances(3, first=10)
    assert np.array_equal(a_m[10:13], c_m)


def test_reference_arm_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["value"] > 0
    from oracle import ref_bench
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_bench.available() else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"] == bench.WORKLOAD
    assert d["config"]["envs"] > 0 and "envs_per_gpu" not in d["config"]  # the line prints what it ran


def test_torch_bounds_equal_numpy():
    import torch

    om, od = bench.make_instances(3000, first=500)
    lb, mat = bench.taillard_bounds(om, od)
    lt, mt = bench.taillard_bounds_torch(torch.from_numpy(om), torch.from_numpy(od), chunk=1024)
    assert np.array_equal(lb, lt.numpy()) and np.array_equal(mat, mt.numpy())
    a, b = bench.make_instances(2 * bench.BLOCK + 7, first=11, threads=1)
    c, d = bench.make_instances(2 * bench.BLOCK + 7, first=11, threads=4)
    assert np.array_equal(a, c) and np.array_equal(b, d)


def test_reference_pool_matches_oracle():
    """The bench's CPU baseline steps the UNMODIFIED reference on the bench's own instances and Philox action
    streams; its finished episodes must be the oracle's (and, in bench.py, the GPU's)."""
    import pytest
    from oracle import ref_bench

    if not ref_bench.available():
        pytest.skip("reference not installed (oracle/make_ref.py)")
    pool = bench.reference_pool(2, episodes_per_worker=1)
    r = pool.run(0)
    pool.close()
    ora, _ = bench.cpu_run(2, 2)
    got = {i: (mk, ns) for i, mk, ns, term in r["episodes"]}
    assert got == {i: (int(ora["makespan"][i]), int(ora["n_steps"][i])) for i in range(2)}
