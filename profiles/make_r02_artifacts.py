"""Turn the round-2 captures brought back in gpurun_out/ into the tracked evidence under profiles/.

    python profiles/make_r02_artifacts.py

Inputs (made on the B200 by the commands in profiles/README.md):
  gpurun_out/r02_traffic.csv        ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
                                    smsp__inst_executed.sum --csv of `prof_workload.py 262144 620`
  gpurun_out/r02_full_<name>.ncu-rep   ncu --set full --import-source on captures (step, rollout, queues, stoch)
Outputs:
  profiles/r02_traffic.json                 per-env-step DRAM bytes / warp-instructions of k_rollout and k_step (read by bench.py)
  profiles/r02_ncu_full_<name>.csv          `--page raw --csv` of each capture
  profiles/r02_lines_<name>.txt             per-source-line / per-function instruction and stall summary (ncu_lines.py)
  profiles/r02_sass_<kernel>.txt            cuobjdump -sass opcode histogram of the kernels in the shipped library
"""
import csv
import json
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
GO, PR = ROOT / "gpurun_out", ROOT / "profiles"


def traffic():
    p = GO / "r02_traffic.csv"
    if not p.exists():
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    hdr = rows[0]
    ci = {n: i for i, n in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if not r[ci["ID"]].isdigit():
            continue
        k = "k_rollout" if "k_rollout" in r[ci["Kernel Name"]] else "k_step" if "k_step" in r[ci["Kernel Name"]] else None
        if k is None:
            continue
        per.setdefault((k, r[ci["ID"]]), {})[r[ci["Metric Name"]]] = (float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]])
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "inst": 1}
    agg = {}
    for (k, _), m in per.items():
        a = agg.setdefault(k, {"n": 0, "read": 0.0, "write": 0.0, "sec": 0.0, "inst": 0.0})
        a["n"] += 1
        a["read"] += m["dram__bytes_read.sum"][0] * unit[m["dram__bytes_read.sum"][1]]
        a["write"] += m["dram__bytes_write.sum"][0] * unit[m["dram__bytes_write.sum"][1]]
        a["sec"] += m["gpu__time_duration.sum"][0] * unit[m["gpu__time_duration.sum"][1]]
        a["inst"] += m["smsp__inst_executed.sum"][0]
    meta = json.loads((GO / "r02_traffic_meta.json").read_text())  # {"envs":..., "rollout_env_steps_per_launch":...}
    envs, rsteps = meta["envs"], meta["rollout_env_steps_per_launch"]
    ro, st = agg["k_rollout"], agg["k_step"]
    out = {
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum "
                  f"on profiles/prof_workload.py {envs} 620 (3 rollout launches, 620 step launches)",
        "envs": envs,
        "rollout_env_steps_per_launch": rsteps,
        "rollout_dram_bytes_per_launch": (ro["read"] + ro["write"]) / ro["n"],
        "rollout_dram_bytes_per_env_step": (ro["read"] + ro["write"]) / ro["n"] / rsteps,
        "rollout_warp_instr_per_env_step": ro["inst"] / ro["n"] / rsteps,
        "rollout_launch_ms_under_ncu": 1e3 * ro["sec"] / ro["n"],
        "step_launches": st["n"],
        "step_dram_bytes_per_env_step": (st["read"] + st["write"]) / st["n"] / envs,
        "step_dram_read_per_env_step": st["read"] / st["n"] / envs,
        "step_dram_write_per_env_step": st["write"] / st["n"] / envs,
        "step_warp_instr_per_env_step": st["inst"] / st["n"] / envs,
        "step_launch_us_under_ncu": 1e6 * st["sec"] / st["n"],
    }
    (PR / "r02_traffic.json").write_text(json.dumps(out, indent=1))
    print("r02_traffic.json:", json.dumps(out)[:400])


def full_captures():
    for rep in sorted(GO.glob("r02_full_*.ncu-rep")):
        name = rep.stem[len("r02_full_"):]
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        (PR / f"r02_ncu_full_{name}.csv").write_text(raw)
        src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--print-source", "cuda,sass", "--csv"],
                             capture_output=True, text=True).stdout
        tmp = GO / f"r02_src_{name}.csv"
        tmp.write_text(src)
        units = "1"
        mp = GO / f"r02_full_{name}.units"
        if mp.exists():
            units = mp.read_text().strip()
        txt = subprocess.run([sys.executable, str(PR / "ncu_lines.py"), str(tmp), units, "40",
                              str(ROOT / "jobshoplab_b200" / "csrc" / "jsl_engine.cuh")], capture_output=True, text=True).stdout
        (PR / f"r02_lines_{name}.txt").write_text(f"# ncu source page of {rep.name}, per {units} units (env-steps of the captured launch)\n" + txt)
        print("capture", name, "->", f"r02_ncu_full_{name}.csv, r02_lines_{name}.txt")


def sass_histograms():
    obj = ROOT / "jobshoplab_b200" / "csrc" / "libjsl_b200.so"
    want = {"k_step_1_0_2": "_ZN3jsl6k_stepILi1ELb0ELi2ELb0EEEvNS_8BatchDevEPKh12jsl_step_outi",
            "k_rollout_1_0_2": "_ZN3jsl9k_rolloutILi1ELb0ELi2EEEvNS_8BatchDevEiPKhlii15jsl_rollout_out",
            "k_rollout_1_0_3": "_ZN3jsl9k_rolloutILi1ELb0ELi3EEEvNS_8BatchDevEiPKhlii15jsl_rollout_out",
            "k_rollout_1_1_4": "_ZN3jsl9k_rolloutILi1ELb1ELi4EEEvNS_8BatchDevEiPKhlii15jsl_rollout_out"}
    for name, sym in want.items():
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, str(obj)], capture_output=True, text=True).stdout
        ops = Counter()
        n = 0
        for line in sass.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(2).split(".")[0]] += 1
                n += 1
        ptx = ""
        log = ROOT / "jobshoplab_b200" / "csrc" / "jsl_kernels_jw1.ptxas.log"
        if log.exists():
            lines = log.read_text().splitlines()
            for i, l in enumerate(lines):
                if f"Function properties for {sym}" in l:
                    ptx = " | ".join(x.strip().replace("ptxas info    : ", "") for x in lines[i + 1:i + 3])
        body = [f"# cuobjdump -sass -fun {sym} libjsl_b200.so: {n} instructions ({n * 16 / 1024:.1f} KB)", f"# ptxas: {ptx}"]
        body += [f"{c:6d} {100 * c / max(n, 1):5.1f} %  {op}" for op, c in ops.most_common()]
        (PR / f"r02_sass_{name}.txt").write_text("\n".join(body) + "\n")
        print(name, n, "instructions;", ptx)


if __name__ == "__main__":
    traffic()
    full_captures()
    sass_histograms()
