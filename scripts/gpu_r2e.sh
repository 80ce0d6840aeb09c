# round 2e: pipeline with one child decode + prefetch + multi-pass candidates (GPU tests), the
# default bench line with the discovery wall-time leg
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu.log
tail -5 gpurun_out/r2e_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2e_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9)
print(json.dumps(d['discovery_wall'], indent=1))
PY
KDF_PREFETCH_PARENTS=0 KDF_CHILD_CACHE_GB=0 timeout 900 python bench.py --steps 3 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep > gpurun_out/r2e_bench_wall_nocache.json 2> gpurun_out/r2e_bench_wall_nocache.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench_wall_nocache.json').read().strip().splitlines()[-1])
print("no cache / no prefetch:", json.dumps(d['discovery_wall'], indent=1))
PY
