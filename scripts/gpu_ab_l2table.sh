for mb in 100 64 48; do
  export KDF_L2_TABLE_MB=$mb
  python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-random-bench > gpurun_out/ab_l2t_$mb.json 2> gpurun_out/ab_l2t_$mb.err || tail -5 gpurun_out/ab_l2t_$mb.err
  python - $mb <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open('gpurun_out/ab_l2t_%s.json'%v).read().strip().splitlines()[-1])
print("L2_TABLE_MB=%-4s %.2f G/s %.1f ms | "%(v,d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
PY
done
