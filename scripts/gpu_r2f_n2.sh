# round 2f (2 GPUs): distributed parity tests, bench line at N=2, and a scaled-down config 4
# (512 Mbp genome over 2 GPUs: samples as stream lists, 4 hash-range passes forced)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2f_pytest_gpu_n2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu_n2.txt
tail -5 gpurun_out/r2f_pytest_gpu_n2.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2f_bench_n2.err
timeout 1200 $T bench.py --gpus 2 --total-genome-mbp 512 --n-passes 4 --steps 2 --warmup 1 --no-k-sweep --no-wall > gpurun_out/r2f_bench_512mbp_n2.json 2> gpurun_out/r2f_bench_512mbp_n2.err; echo "512mbp rc=$?"
tail -5 gpurun_out/r2f_bench_512mbp_n2.err
python - <<'PY'
import json
for f in ("r2f_bench_n2", "r2f_bench_512mbp_n2"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "%.1f G/s %.1f ms e2e %s | "%(d['value']/1e9,d['ms_per_step'], d['e2e'] and "%.1f G/s"%(d['e2e']['value']/1e9))+" ".join("%s=%.1f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
        print(d['stage_sizes'], d['count_passes'], d['peak_hbm_bytes_rank0'], d['parity_checked'] and d['parity_checked']['ok'], d['synth_seconds'])
    except Exception as e: print(f, "ERR", e)
PY
