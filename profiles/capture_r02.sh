#!/bin/bash
# Round-2 evidence capture, run on the B200 through gpurun (one GPU):
#   gpurun --timeout 2400 -- 'bash profiles/capture_r02.sh'
# then, back in the build container:  python profiles/make_r02_artifacts.py
set -x
O=gpurun_out
# 1. the bench lines themselves (never under a profiler)
python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
python profiles/secondary_bench.py > $O/r02_secondary.jsonl 2> $O/r02_secondary.err
# 2. launch list of the bench command (shares of the step, cold-cache serialised times)
CMD="python bench.py --steps 2 --warmup 3 --envs-per-gpu 65536 --step-mode-steps 40 --no-cpu --no-secondary"
$CMD > $O/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $CMD > $O/r02_ncu_launches.log 2>&1
# 3. DRAM traffic and executed warp-instructions of every launch of the capture workload
W="python profiles/prof_workload.py 262144 620"
$W > $O/r02_plain_workload.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:'k_rollout|k_step' --csv --log-file $O/r02_traffic.csv $W > $O/r02_ncu_traffic.log 2>&1
python - <<PY
import json, re
t = open("$O/r02_plain_workload.log").read()
json.dump({"envs": 262144, "rollout_env_steps_per_launch": int(re.search(r"per launch: (\d+)", t).group(1))}, open("$O/r02_traffic_meta.json", "w"))
PY
# 4. full captures with source correlation
S="python profiles/prof_workload.py 65536 620"
$S > $O/r02_plain_s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 600 -c 1 -o $O/r02_full_step $S > $O/r02_ncu_s.log 2>&1
echo 65536 > $O/r02_full_step.units
R="python profiles/prof_workload.py 16384 1"
$R > $O/r02_plain_r.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 1 -c 1 -o $O/r02_full_rollout $R > $O/r02_ncu_r.log 2>&1
grep -o "per launch: [0-9]*" $O/r02_plain_r.log | grep -o "[0-9]*" > $O/r02_full_rollout.units
for w in fifo stoch; do
  G="python profiles/prof_general.py 16384 $w"
  $G > $O/r02_plain_$w.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 1 -c 1 -o $O/r02_full_$w $G > $O/r02_ncu_$w.log 2>&1
  grep -o "per launch: [0-9]*" $O/r02_plain_$w.log | grep -o "[0-9]*" > $O/r02_full_$w.units
done
ls -la $O | tail -30
