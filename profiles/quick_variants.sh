for lib in libjsl_b200.so libjsl_b200_c10.so libjsl_b200_c12.so; do
  JSL_LIB=/root/repo/jobshoplab_b200/csrc/$lib python bench.py --envs-per-gpu 65536 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('$lib', 'rollout', round(d['value']/1e6,1), 'M/s  step_mode', round(d['step_mode']['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1))"
done
