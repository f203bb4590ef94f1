"""jobshoplab_b200.compiler (spec file / YAML DSL -> SoA) against the lowering of the reference's own
Compiler.compile() output stored in tests/golden (mapper.py:1201, :1387), plus the reference's
lower-bound known answers (tests/unit_tests/test_lower_bound_calculation.py:10-72)."""
import numpy as np
import pytest

import _golden as G
from jobshoplab_b200 import compiler as C
from jobshoplab_b200 import lowered as L

FIELDS = ("job_op_start", "op_machine", "op_tool", "pre_type", "pre_cap", "post_type", "post_cap", "sbuf_type",
          "sbuf_cap", "sbuf_role", "buf_ref_id", "agv_init_place", "job_init_buf", "m_outage_start")
TIMED = (("op_duration", "op_spec"), ("setup_time", "setup_spec"), ("travel_time", "travel_spec"),
         ("m_outage_freq", "m_ofreq_spec"), ("m_outage_dur", "m_odur_spec"),
         ("t_outage_freq", "t_ofreq_spec"), ("t_outage_dur", "t_odur_spec"))


def check(sc: G.Scenario):
    ref = sc.lowered
    mine = C.compile_spec_text(sc.source_text) if sc.kind == "spec" else C.compile_dsl_text(sc.source_text)
    for k in ("n_jobs", "n_machines", "n_transports", "n_sbuf", "n_tools", "default_tool", "start_time"):
        assert getattr(mine, k) == getattr(ref, k), k
    assert mine.tools == ref.tools
    for k in FIELDS:
        assert np.array_equal(getattr(mine, k), getattr(ref, k)), k
    for arr, spec in TIMED:
        a, b = getattr(mine, arr), getattr(ref, arr)
        sa, sb = getattr(mine, spec), getattr(ref, spec)
        assert a.shape == b.shape, arr
        if sb.kind is None:
            assert sa.kind is None, spec
            assert np.array_equal(a, b), arr
        else:
            static = sb.kind == L.TIME_STATIC
            assert np.array_equal(sa.kind, sb.kind) and np.array_equal(sa.base, sb.base), spec
            assert np.allclose(sa.param, sb.param), spec
            assert np.array_equal(a[static], b[static]), arr  # stochastic samples differ by construction
    if ref.is_deterministic:
        assert mine.lower_bound == ref.lower_bound and mine.max_allowed_time == ref.max_allowed_time


@pytest.mark.parametrize("group", G.GROUPS)
def test_compiler_matches_reference_lowering(group):
    seen = set()
    for name in G.names(group):
        sc = G.Scenario(group, name)
        key = hash(sc.source_text)
        if key in seen:
            continue
        seen.add(key)
        try:
            check(sc)
        except AssertionError as ex:
            raise AssertionError(f"{group}:{name}: {ex}") from ex


def test_lower_bound_known_answers():
    kat = {"ft06": 52, "ft10": 796, "ta01": 1005}
    for name, lb in kat.items():
        sc = G.Scenario("classic", f"{name}/fifo")
        assert C.compile_spec_text(sc.source_text).lower_bound == lb
    for name, lb in {"la01": 666, "la02": 655}.items():
        sc = G.Scenario("solutions", f"{name}/solution")
        assert C.compile_spec_text(sc.source_text).lower_bound == lb
    mat = {"ft06": 197, "ft10": 5109, "la16": 5351, "ta01": 11671}
    for name, v in mat.items():
        assert G.Scenario("classic", f"{name}/fifo").lowered.max_allowed_time == v


def test_repositories(tmp_path):
    sc = G.Scenario("classic", "ft06/fifo")
    p = tmp_path / "ft06"
    p.write_text(sc.source_text)
    li = C.Compiler(repo=C.SpecRepository(p)).compile()
    assert (li.n_jobs, li.n_machines, li.n_transports, li.n_sbuf) == (6, 6, 6, 2)
    assert li.job_init_buf.tolist() == [24] * 6  # input buffer b-24
    sc = G.Scenario("transport", "ft06-t/fifo")
    q = tmp_path / "ft06.yaml"
    q.write_text(sc.source_text)
    a = C.Compiler(repo=C.DslRepository(q)).compile()
    b = C.Compiler(repo=C.DslStrRepository(sc.source_text)).compile()
    assert np.array_equal(a.travel_time, b.travel_time) and a.travel_time.max() == 9
    with pytest.raises(FileNotFoundError):
        C.SpecRepository(tmp_path / "missing")
