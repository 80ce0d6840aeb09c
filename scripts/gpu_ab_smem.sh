# parity, then shared-memory count vs L2 packed count on one B200
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -x -q 2>&1 | tail -15
for v in "$@"; do
  unset KDF_COUNT_SMEM KDF_SMEM_PARTS
  case $v in
    l2) export KDF_COUNT_SMEM=0;;
    smem128) export KDF_SMEM_PARTS=128;;
    smem64) export KDF_SMEM_PARTS=64;;
  esac
  python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-random-bench > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -5 gpurun_out/ab_$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open('gpurun_out/ab_%s.json'%v).read().strip().splitlines()[-1])
print("%-7s %.2f G/s %.1f ms | "%(v,d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
print(d['stage_sizes'])
PY
done
