for v in base "$@"; do
  if [ $v = base ]; then unset KDF_LIB; else export KDF_LIB=$PWD/build/libkdf_$v.so; fi
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err || tail -3 gpurun_out/var_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open('gpurun_out/var_%s.json'%v).read().strip().splitlines()[-1])
print("%-5s %.2f G/s %.1f ms | "%(v,d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:12]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>1))
PY
done
