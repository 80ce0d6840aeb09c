# quick correctness + kernel-time check on one B200
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench "$@" > gpurun_out/quick.json 2> gpurun_out/quick.err
tail -3 gpurun_out/quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])
print("value %.2f G/s  ms/step %.1f  launches %d"%(d['value']/1e9,d['ms_per_step'],d['gpu_launches']))
for k,v in d['kernels'].items(): print("  %-28s n=%3d avg %.3f ms total/step %.2f"%(k,v['launches'],v['ms_avg'],v['ms_total']/d['steps']))
print(d['stage_sizes'])
PY
