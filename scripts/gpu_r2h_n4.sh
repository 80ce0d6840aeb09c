# round 2h (4 GPUs): all distributed tests (chain vs oracle, pipeline vs golden files at world 2 and
# 4), the weak-scaling bench line at N=4 and BASELINE config 4 (3 Gbp x 30x) on 4 GPUs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2h_pytest_gpu_dist_n4.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_gpu_dist_n4.txt
tail -5 gpurun_out/r2h_pytest_gpu_dist_n4.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2h_bench_n4.err; echo "bench n4 rc=$?"
tail -3 gpurun_out/r2h_bench_n4.err
timeout 1200 $T bench.py --gpus 4 --total-genome-mbp 3000 --steps 2 --warmup 1 --no-k-sweep --no-wall > gpurun_out/r2h_bench_config4_n4.json 2> gpurun_out/r2h_bench_config4_n4.err; echo "config4 rc=$?"
tail -5 gpurun_out/r2h_bench_config4_n4.err
python - <<'PY'
import json
for f in ("r2h_bench_n4", "r2h_bench_config4_n4"):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, "%.1f G/s %.1f ms e2e %s | "%(d['value']/1e9,d['ms_per_step'], d['e2e'] and "%.1f G/s"%(d['e2e']['value']/1e9))+" ".join("%s=%.1f"%(k.split('/')[0][:14]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
        print(d['stage_sizes'], d['count_passes'], d['peak_hbm_bytes_rank0'], d['parity_checked'] and d['parity_checked']['ok'])
    except Exception as e: print(f, "ERR", e)
PY
