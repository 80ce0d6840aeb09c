# round 2l (2 GPUs): distributed tests, the default bench line at N=2 with the multi-GPU wall leg
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2l_pytest_gpu_dist_n2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_gpu_dist_n2.txt
tail -4 gpurun_out/r2l_pytest_gpu_dist_n2.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
KDF_BAM_TIMING=1 timeout 900 $T bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2l_bench_n2.err
python - <<'PY'
import json
for f in ("r2l_bench_n2",):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "%.1f G/s %.1f ms e2e %s | "%(d['value']/1e9,d['ms_per_step'], d['e2e'] and "%.1f G/s"%(d['e2e']['value']/1e9))+" ".join("%s=%.1f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.3))
        print(d['stage_sizes'], d['count_passes'], d['parity_checked'] and d['parity_checked']['ok'])
        print(json.dumps(d['discovery_wall'])[:1500])
    except Exception as e: print(f, "ERR", e)
PY
