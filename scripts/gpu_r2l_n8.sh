# round 2l (8 GPUs): the default bench line at N=8 (weak scaling + the multi-GPU wall leg) and
# BASELINE config 4 (3 Gbp x 30x trio, hash-partitioned across 8 B200)
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
nproc
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
KDF_BAM_TIMING=1 timeout 900 $T bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err; echo "bench n8 rc=$?"
tail -3 gpurun_out/r2l_bench_n8.err
timeout 900 $T bench.py --gpus 8 --total-genome-mbp 3000 --steps 3 --warmup 1 --no-k-sweep --no-wall > gpurun_out/r2l_bench_config4_n8.json 2> gpurun_out/r2l_bench_config4_n8.err; echo "config4 rc=$?"
tail -3 gpurun_out/r2l_bench_config4_n8.err
python - <<'PY'
import json
for f in ("r2l_bench_n8", "r2l_bench_config4_n8"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "%.1f G/s %.1f ms e2e %s | "%(d['value']/1e9,d['ms_per_step'], d['e2e'] and "%.1f G/s"%(d['e2e']['value']/1e9))+" ".join("%s=%.1f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.3))
        print(d['stage_sizes'], d['count_passes'], d['peak_hbm_bytes_rank0'], d['parity_checked'] and d['parity_checked']['ok'])
        print(json.dumps(d['discovery_wall'])[:1200])
    except Exception as e: print(f, "ERR", e)
PY
