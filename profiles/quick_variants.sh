# A/B helper: bench every libjsl_b200*.so found in csrc/ on a 65536-env batch
for lib in jobshoplab_b200/csrc/libjsl_b200*.so; do
  JSL_LIB=$PWD/$lib python bench.py --envs-per-gpu 65536 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('$lib', 'rollout', round(d['value']/1e6,1), 'M/s  step_mode', round(d['step_mode']['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1))"
done
