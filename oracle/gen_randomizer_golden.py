"""Golden vectors for InstanceRandomizer -- TEST INFRASTRUCTURE (run in the build container only).

Runs the UNMODIFIED reference manipulator (jobshoplab/compiler/manipulators.py:99-150) behind the import
shims, with seeded global generators, and stores
  * exact results for a few (instance, numpy seed, random seed) triples: the host-side manipulator
    (jobshoplab_b200.compiler.randomize_instance) must reproduce them bit for bit;
  * 3000 unseeded-in-sequence draws on ft06 (op machines + durations): the distribution the device generator
    (jsl_batch_randomize) is compared with (two-sample chi-square).

    python oracle/gen_randomizer_golden.py      -> tests/golden/randomizer.npz
"""
import random
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_harness as H  # noqa: E402
from jobshoplab.compiler.manipulators import InstanceRandomizer  # noqa: E402
from jobshoplab.utils.utils import calculate_lower_bound, get_max_allowed_time  # noqa: E402

SPEC = H.REF_ROOT / "data" / "jssp_instances" / "spec_files"


def compiled(name):
    cfg = H.make_config()
    repo = H.SpecRepository(SPEC / name, loglevel="warning", config=cfg)
    return cfg, H.Compiler(cfg, loglevel="warning", repo=repo).compile()


def tables(instance):
    ops = [o for j in instance.instance.specification for o in j.operations]
    return (np.array([int(o.machine.split("-")[1]) for o in ops], np.int32),
            np.array([int(o.duration.time) for o in ops], np.int32))


def main():
    out = {}
    names = []
    for name, ns, rs in (("ft06", 1, 2), ("ft06", 7, 7), ("la01", 3, 11), ("ft10", 5, 0), ("ta01", 2024, 99)):
        cfg, (inst, init) = compiled(name)
        np.random.seed(ns)
        random.seed(rs)
        new, _ = InstanceRandomizer("warning", cfg).manipulate(inst, init)
        om, od = tables(new)
        key = f"exact/{name}/{ns}/{rs}"
        names.append(key)
        out[key + "/op_machine"], out[key + "/op_duration"] = om, od
        out[key + "/bounds"] = np.array([calculate_lower_bound(new), get_max_allowed_time(new)], np.int64)
    cfg, (inst, init) = compiled("ft06")
    np.random.seed(12345)
    random.seed(54321)
    R = 3000
    man = InstanceRandomizer("warning", cfg)
    oms, ods = [], []
    for _ in range(R):
        new, _ = man.manipulate(inst, init)  # always from the same base instance: same duration range
        om, od = tables(new)
        oms.append(om)
        ods.append(od)
    out["dist/ft06/op_machine"], out["dist/ft06/op_duration"] = np.stack(oms), np.stack(ods)
    out["names"] = np.array(names)
    dst = HERE.parent / "tests" / "golden" / "randomizer.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
