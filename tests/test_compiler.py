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


def test_lowering_of_reference_objects():
    """jobshoplab_b200.ref_lowering (what the seam-2 middleware does with the `(InstanceConfig, State)` the reference
    env hands it) == the lowering stored in the goldens, for every distinct golden source incl. stochastic ones
    (time behaviours, numpy start seeds).  Needs the reference (oracle/_ref or /root/reference)."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
    from oracle import ref_bench
    if not ref_bench.available():
        pytest.skip("the reference is not installed (python oracle/make_ref.py)")
    import ref_harness as rh
    from jobshoplab_b200 import ref_lowering as RL
    import dataclasses

    seen, n = set(), 0
    for group in ("classic", "transport", "buffers", "features"):
        for name in G.names(group)[::4]:
            sc = G.Scenario(group, name)
            if hash(sc.source_text) in seen:
                continue
            seen.add(hash(sc.source_text))
            env = rh.make_env("spec" if sc.kind == "spec" else "dslstr", _as_file(sc) if sc.kind == "spec" else sc.source_text)
            inst, init = env.jsl_rc.inner.compile()
            a, b = rh.lower_reference(inst, init), RL.lower_reference(inst, init)
            for f in dataclasses.fields(a):
                x, y = getattr(a, f.name), getattr(b, f.name)
                if f.name.endswith("_spec"):
                    for k in ("kind", "base", "param"):
                        xx, yy = getattr(x, k), getattr(y, k)
                        assert (xx is None) == (yy is None) and (xx is None or np.array_equal(xx, yy)), (name, f.name, k)
                elif isinstance(x, (np.ndarray, list)):
                    assert np.array_equal(np.asarray(x), np.asarray(y)), (name, f.name)
                else:
                    assert x == y, (name, f.name)
            assert np.array_equal(a.start_seeds, b.start_seeds), name
            n += 1
    assert n >= 20


def _as_file(sc):
    import tempfile
    from pathlib import Path

    d = Path(tempfile.mkdtemp())
    p = d / "inst"
    p.write_text(sc.source_text)
    return p
