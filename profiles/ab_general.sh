# A/B of every libjsl_b200*.so in csrc/ on the headline (262144 envs) and on the non-classic workloads
for lib in jobshoplab_b200/csrc/libjsl_b200*.so; do
  case $lib in *_log.so|*_user.so) continue;; esac
  echo "== $lib"
  JSL_LIB=$PWD/$lib python bench.py --envs-per-gpu 262144 --steps 3 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('rollout', round(d['value']/1e6,1), 'step_mode', round(d['step_mode']['value']/1e6,1))"
  JSL_LIB=$PWD/$lib JSL_SECONDARY=general python profiles/secondary_bench.py 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d['workload'][:58], round(d['env_steps_per_sec']/1e6,1))"
done
