# round 2n (4 GPUs): the default bench line at N=4 (weak scaling + the multi-GPU wall leg)
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2n_bench_n4.json 2> gpurun_out/r2n_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n_bench_n4.json') if l.startswith('{')][-1])
print("%.1f G/s %.1f ms e2e %s | "%(d['value']/1e9,d['ms_per_step'], d['e2e'] and "%.1f G/s"%(d['e2e']['value']/1e9))+" ".join("%s=%.1f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.3))
print(d['stage_sizes'], d['parity_checked'] and d['parity_checked']['ok'])
print(json.dumps(d['discovery_wall'])[:900])
PY
