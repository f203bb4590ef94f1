"""Bit-exact makespans on ALL bundled deterministic instances (BASELINE.json north_star).

tests/golden/rules_all.npz (oracle/gen_rule_table.py) holds what the UNMODIFIED reference's rule drivers
(jobshoplab/utils/dipatching_benchmark.py:17-67, 131-137) produce with the default Config() on every spec file,
every transport YAML and every deterministic DSL YAML the reference ships: makespan, env steps, no-op steps and the
sum of rewards for fifo (always accept) / spt / "mwkr".

Two shipped files are rejected by the reference's own compiler (spec_files/swv01: a comment banner its parser
chokes on; transport/ta01.yaml: a malformed travel matrix) and four DSL files carry stochastic times; the table lists
them under meta["skipped"].

  * CPU (`-m "not gpu"`): the product compiler lowers each source text and the C oracle reproduces every entry --
    this pins the oracle (and the compiler) on the full instance set;
  * GPU (`-m gpu`): `jsl_rollout` of the THROUGHPUT build (libjsl_b200.so, the binary bench.py measures) with the
    in-kernel policies reproduces every entry on every env of a small batch.
"""
import json
from functools import lru_cache
from pathlib import Path

import numpy as np
import pytest

from jobshoplab_b200 import abi, compiler

TABLE = Path(__file__).resolve().parent / "golden" / "rules_all.npz"
POLICY = {"fifo": abi.POLICY_FIFO, "spt": abi.POLICY_SPT, "mwkr": abi.POLICY_MWKR}


@lru_cache(maxsize=None)
def table():
    z = np.load(TABLE, allow_pickle=False)
    names, kinds, texts, rules = (json.loads(str(z[k])) for k in ("names", "kinds", "texts", "rules"))
    return names, kinds, texts, rules, z["table"], z["ret"], json.loads(str(z["meta"]))


@lru_cache(maxsize=None)
def lowered(i):
    names, kinds, texts = table()[:3]
    return compiler.compile_spec_text(texts[i]) if kinds[i] == "spec" else compiler.compile_dsl_text(texts[i])


def test_table_covers_every_bundled_instance():
    names, kinds, _, rules, tab, ret, meta = table()
    assert rules == ["fifo", "spt", "mwkr"]
    # every file under data/jssp_instances/{spec_files,transport,dsl} is either in the table or listed with the reason
    # the reference itself gives for it: 167 spec files, 45 transport YAMLs, 7 DSL YAMLs
    skipped = meta["skipped"]
    assert sum(n.startswith("spec/") for n in names) == 166 and skipped["spec/swv01"] == "compile: ValueError"
    assert sum(n.startswith("transport/") for n in names) == 44 and skipped["transport/ta01"] == "compile: InstanceSchemaError"
    assert sum(n.startswith("dsl/") for n in names) == 3
    assert sorted(k for k, v in skipped.items() if v == "stochastic") == [
        "dsl/instance_3x3_full", "dsl/real_world_deterministic_6x6_instance", "dsl/real_world_instance",
        "dsl/real_world_stochastic_6x6_instance"]  # time behaviours: covered by the features / ks fixtures instead
    assert tab.shape == (len(names), 3, 5) and ret.shape == (len(names), 3)
    assert (tab[:, :, 3] == 1).all() and (tab[:, :, 4] == 0).all()  # every rule rollout terminates
    # the figures the reference publishes / the survey measured
    want = {"spec/ft06": (68, 88, 83), "spec/ft10": (1262, 1074, 1262), "spec/la16": (1230, 1156, 1268),
            "spec/ta01": (1830, 1462, 1501), "spec/ta41": (2973, 2465, 3120), "transport/ft06": (83, 85, 85),
            "transport/ft10": (1463, 1167, 1286), "transport/ft20": (1591, 1364, 1484),
            "transport/la01": (973, 883, 1001), "transport/la02": (903, 796, 953)}
    for k, mk in want.items():
        assert tuple(tab[names.index(k), :, 0]) == mk, k


def test_oracle_reproduces_every_entry():
    from oracle import oracle_py as op

    names, _, _, rules, tab, ret, meta = table()
    for i, name in enumerate(names):
        li = lowered(i)
        assert [li.n_jobs, li.n_machines, li.n_transports, li.n_sbuf] == meta[name]["shape"], name
        assert [li.lower_bound, li.max_allowed_time] == meta[name]["lb_mat"], name
        for r, rule in enumerate(rules):
            o = op.OracleEnv(li)
            o.rollout(POLICY[rule])
            got = (o.time, o.n_steps, o.noop_count, int(o.terminated), int(o.truncated))
            assert got == tuple(int(x) for x in tab[i, r]), (name, rule, got, tab[i, r].tolist())
            assert o.ret == float(ret[i, r]), (name, rule, "return")


@pytest.mark.gpu
def test_bit_exact_makespans_on_all_bundled_deterministic_instances():
    torch = pytest.importorskip("torch")
    from jobshoplab_b200 import engine

    names, _, _, rules, tab, ret, _ = table()
    B = 3
    checked = 0
    for i, name in enumerate(names):
        batch = engine.Batch(lowered(i), B)          # throughput build
        assert batch.L is engine.load_library(False)
        for r, rule in enumerate(rules):
            res = batch.rollout(rule, reset_first=True)
            got = np.stack([res.makespan.cpu().numpy(), res.n_steps.cpu().numpy(), res.no_op_count.cpu().numpy()], axis=1)
            st = res.status.cpu().numpy()
            assert (st == abi.ST_DONE).all(), (name, rule, st)
            assert (got == tab[i, r, :3]).all(), (name, rule, got.tolist(), tab[i, r].tolist())
            assert (res.ret.cpu().numpy() == float(ret[i, r])).all(), (name, rule, "return")
            checked += 1
        batch.close()
    assert checked == 3 * len(names)
