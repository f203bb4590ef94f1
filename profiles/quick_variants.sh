# A/B helper: bench every libjsl_b200*.so found in csrc/ (not the logging build) on a 262144-env batch
for lib in jobshoplab_b200/csrc/libjsl_b200*.so; do
  case $lib in *_log.so) continue;; esac
  JSL_LIB=$PWD/$lib python bench.py --envs-per-gpu ${ENVS:-262144} --steps 3 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('$lib', 'rollout', round(d['value']/1e6,1), 'M/s  step_mode', round(d['step_mode']['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1))"
done
