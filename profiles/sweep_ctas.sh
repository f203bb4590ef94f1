# resident rollout CTAs per SM vs throughput of the non-classic kernels (JSL_ROLLOUT_CTAS_PER_SM is a diagnostic knob)
for n in 32 28 24 20 16; do echo "== ctas_per_sm $n"; JSL_ROLLOUT_CTAS_PER_SM=$n JSL_SECONDARY=general python profiles/secondary_bench.py 2>&1 | grep -E "fifo5|stochastic|ta41-t" | cut -c1-150; done
