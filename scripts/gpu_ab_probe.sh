# parity of the binned count --if route, then direct vs binned probing of a filter table larger than L2
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -4
for v in binned; do
  unset KDF_PROBE_DIRECT; export KDF_PROBE_DIRECT_MB=100
  if [ $v = direct ]; then export KDF_PROBE_DIRECT=1; fi
  python bench.py --steps 2 --warmup 1 --genome-mbp ${1:-160} --no-e2e --no-cpu-baseline --no-random-bench > gpurun_out/probe_$v.json 2> gpurun_out/probe_$v.err || tail -5 gpurun_out/probe_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open('gpurun_out/probe_%s.json'%v).read().strip().splitlines()[-1])
print("%-7s %.2f G/s %.1f ms | "%(v,d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:12]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
print(d['stage_sizes'])
PY
done
