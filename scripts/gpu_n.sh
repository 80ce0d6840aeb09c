# multi-GPU parity (NCCL) and the bench line at N GPUs: bash scripts/gpu_n.sh N [pytest]
N=$1
if [ "$2" = pytest ]; then
  python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -4
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
for line in open('gpurun_out/bench_n%s.json'%n):
    if line.startswith('{'):
        d=json.loads(line)
        print("N=%s %.1f G/s %.1f ms e2e %.1f G/s %.1f ms | "%(n,d['value']/1e9,d['ms_per_step'],d['e2e']['value']/1e9,d['e2e']['ms_per_step'])+" ".join("%s=%.1f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
        print(d['stage_sizes'])
PY
