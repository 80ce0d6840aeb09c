python -m pytest tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -5
for v in filter nofilter; do
  unset KDF_TABLE_FILTER
  if [ $v = nofilter ]; then export KDF_TABLE_FILTER=0; fi
  python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-random-bench "$@" > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -5 gpurun_out/ab_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open('gpurun_out/ab_%s.json'%v).read().strip().splitlines()[-1])
print("%-8s %.2f G/s %.1f ms | "%(v,d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
print(d['stage_sizes'])
PY
done
