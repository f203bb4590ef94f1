"""Golden vectors for the dashboard data of env.render() -- TEST INFRASTRUCTURE (build container only).

Replays stored golden action tapes in the UNMODIFIED reference and records what its dashboard would draw,
DashboardDataMapper.map_states_to_gant_data(env.history) (jobshoplab/env/rendering/gant_dashboard.py:414-429),
half way through the episode and at its end.  The engine's jobshoplab_b200.state_view.gantt_data must return the
same dictionaries from the state digest and the transport-task log.

    python oracle/gen_gantt_golden.py      -> tests/golden/gantt.json
"""
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent / "tests"))
sys.path.insert(0, str(HERE.parent))
import ref_harness as H  # noqa: E402
import _golden as G  # noqa: E402
from jobshoplab.env.rendering.gant_dashboard import DashboardDataMapper  # noqa: E402

PICK = [("classic", "3x3/fifo"), ("classic", "ft06/random0"), ("classic", "ft10/spt"),
        ("transport", "ft06-t/fifo"), ("transport", "ft06-t/noearly/random0"), ("transport", "ft06-t/random1"),
        ("transport", "ft10-t/spt"), ("transport", "la01-t/noearly/mwkr"), ("transport", "ft20-t/random1"),
        ("buffers", "la01-t/fifo1/random0"), ("buffers", "la01-t/fifo2/fifo")]


def clean(d):
    return {k: [dict(x) for x in v] for k, v in d.items()}


def main():
    out = {}
    extra = [("buffers", n) for n in G.names("buffers") if "lifo" in n][:2]
    for group, name in PICK + extra:
        sc = G.Scenario(group, name)
        if not sc.lowered.is_deterministic:
            continue
        cfg = dict(sc.meta.get("cfg", {}))
        if sc.kind == "spec":
            env = H.make_env("spec", sc.meta["src"], **cfg)
        else:
            env = H.make_env("dslstr", sc.source_text, **cfg)
        n = int(sc.meta["n_steps"])
        half = max(1, n // 2)
        rec = {}
        for i in range(n):
            env.step(int(sc.actions[i]))
            if i + 1 == half:
                try:  # (the reference raises when nothing has been started yet, or on a TimeDependency)
                    rec["half"] = {"steps": half, "data": clean(DashboardDataMapper.map_states_to_gant_data(env.history))}
                except (ValueError, AttributeError) as ex:
                    print("  half-way snapshot skipped:", type(ex).__name__)
        try:
            rec["final"] = {"steps": n, "data": clean(DashboardDataMapper.map_states_to_gant_data(env.history))}
        except (ValueError, AttributeError) as ex:  # e.g. an episode that ended in BufferFullError before any op ran
            print("  final snapshot skipped:", type(ex).__name__)
            if not rec:
                continue
        out[f"{group}:{name}"] = rec
        last = rec.get("final", rec.get("half"))
        print(group, name, n, len(last["data"]["schedules"]), len(last["data"]["transports"]))
    dst = HERE.parent / "tests" / "golden" / "gantt.json"
    dst.write_text(json.dumps(out))
    print("wrote", dst, dst.stat().st_size)


if __name__ == "__main__":
    main()
