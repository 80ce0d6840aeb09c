# correctness, then A/B of the packed count form against the plane form on one B200
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -15
for v in packed packed32 packed128; do
  unset KDF_COUNT_BINS_PLANES KDF_SLICE_MB
  if [ $v = planes ]; then export KDF_COUNT_BINS_PLANES=1; fi
  if [ $v = packed32 ]; then export KDF_SLICE_MB=32; fi
  if [ $v = packed128 ]; then export KDF_SLICE_MB=128; fi
  python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-random-bench "$@" > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -5 gpurun_out/ab_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open('gpurun_out/ab_%s.json'%v).read().strip().splitlines()[-1])
print("%-7s %.2f G/s %.1f ms | "%(v,d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:12]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
print(d['stage_sizes'])
PY
done
