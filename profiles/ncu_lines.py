"""Per-source-line summary of an ncu capture taken with --import-source on (kernels built with -lineinfo).

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv
    python profiles/ncu_lines.py X.csv [units] [top]

`units` = env-steps of the captured launch (so counts print per env-step).  Only the first launch in the file is read.
Prints, per source file, the lines with the most executed warp-instructions and their stall samples, and the
totals per function-sized region of jsl_engine.cuh / jsl_kernels.cuh.
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45

rows = list(csv.reader(open(path, newline="")))
files, cur, hdr = [], None, None
seen_first = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        if r[1] in seen_first and cur is not None and files and files[0][0] == r[1]:
            break  # second launch starts
        seen_first.add(r[1])
        cur = (r[1], [])
        files.append(cur)
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if cur is not None and r[0] not in ("", "Line No"):
        cur[1].append(r)

ci = {n: i for i, n in enumerate(hdr)}
# the header repeats "Source": line text is column 1
I_EXEC, I_SAMP = ci["Instructions Executed"], ci["# Samples"]
stall_cols = [(n, i) for n, i in ci.items() if n.startswith("stall_") and "Not Issued" not in n]

total = 0
total_samples = 0
for f, ls in files:
    for r in ls:
        try:
            total += int(r[I_EXEC]); total_samples += int(r[I_SAMP])
        except ValueError:
            pass
print(f"total warp-instructions {total} = {total / units:.1f} per unit; samples {total_samples}")
for f, ls in files:
    sub = sum(int(r[I_EXEC]) for r in ls if r[I_EXEC].isdigit())
    print(f"\n== {f}: {sub / units:.1f} instr/unit ({100 * sub / max(total, 1):.1f} %)")
    ls2 = sorted((r for r in ls if r[I_EXEC].isdigit()), key=lambda r: -int(r[I_EXEC]))[:top]
    for r in ls2:
        st = sorted(((int(r[i]), n[6:]) for n, i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
        print(f"  L{r[0]:>5} {int(r[I_EXEC]) / units:8.2f} instr  {int(r[I_SAMP]):6d} smp  {' '.join(f'{n}:{v}' for v, n in st):<40} | {r[1][:100]}")

# stall totals
agg = defaultdict(int)
for f, ls in files:
    for r in ls:
        for n, i in stall_cols:
            if r[i].isdigit():
                agg[n] += int(r[i])
print("\nstall samples:", ", ".join(f"{n[6:]}={v}" for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))

# ---- per-function totals (optional 4th argument: the source file the capture was compiled from)
if len(sys.argv) > 4:
    import re
    src = open(sys.argv[4]).read().splitlines()
    starts = []
    pat = re.compile(r"^(?:JSL_DN?|JSL_M|JSL_HD|template|static|__device__|__global__|inline)?.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$")
    depth = 0
    for n, line in enumerate(src, 1):
        if depth <= 1 and line and not line.startswith((" ", "\t", "//", "#", "}", "template", "struct", "constexpr", "namespace")) and "(" in line:
            m = pat.match(line)
            if m:
                starts.append((n, m.group(1)))
        depth += line.count("{") - line.count("}")
    fn_tot = defaultdict(lambda: [0, 0])
    target = [ls for f, ls in files if f.endswith(sys.argv[4].split("/")[-1].replace("engine_r01", "jsl_engine"))]
    for ls in target[:1]:
        for r in ls:
            if not r[I_EXEC].isdigit():
                continue
            ln = int(r[0])
            name = "?"
            for n, nm in starts:
                if n <= ln:
                    name = nm
                else:
                    break
            fn_tot[name][0] += int(r[I_EXEC]); fn_tot[name][1] += int(r[I_SAMP])
    print("\nper function (instr/unit, % of all, samples):")
    for nm, (a, b) in sorted(fn_tot.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"  {nm:<28} {a / units:10.1f} {100 * a / total:6.1f} %  {b:7d} smp")
