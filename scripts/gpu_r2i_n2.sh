# round 2i (2 GPUs): BASELINE config 4 (3 Gbp x 30x trio) with the table hash-partitioned across 2 B200
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 1700 $T bench.py --gpus 2 --total-genome-mbp 3000 --steps 2 --warmup 1 --no-k-sweep --no-wall > gpurun_out/r2i_bench_config4_n2.json 2> gpurun_out/r2i_bench_config4_n2.err; echo "config4 rc=$?"
tail -5 gpurun_out/r2i_bench_config4_n2.err
python - <<'PY'
import json
for f in ("r2i_bench_config4_n2",):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "%.1f G/s %.1f ms e2e %s | "%(d['value']/1e9,d['ms_per_step'], d['e2e'] and "%.1f G/s"%(d['e2e']['value']/1e9))+" ".join("%s=%.1f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
        print(d['stage_sizes'], d['count_passes'], d['peak_hbm_bytes_rank0'], d['parity_checked'] and d['parity_checked']['ok'])
    except Exception as e: print(f, "ERR", e)
PY
