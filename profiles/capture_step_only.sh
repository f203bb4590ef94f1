set -x
O=gpurun_out
W="python profiles/prof_workload.py 262144 620"
$W > $O/r02_plain_workload.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
    -k regex:'k_rollout|k_step' --csv --log-file $O/r02_traffic.csv $W > $O/r02_ncu_traffic.log 2>&1
python - <<PY
import json, re
t = open("$O/r02_plain_workload.log").read()
json.dump({"envs": 262144, "rollout_env_steps_per_launch": int(re.search(r"per launch: (\d+)", t).group(1))}, open("$O/r02_traffic_meta.json", "w"))
PY
S="python profiles/prof_workload.py 65536 620"
$S > $O/r02_plain_s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step -s 600 -c 1 -o $O/r02_full_step $S > $O/r02_ncu_s.log 2>&1
echo 65536 > $O/r02_full_step.units
ls -la $O | tail
