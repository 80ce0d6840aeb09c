# round 2k: new BAM decoder (own inflate, chained walk) + one count launch per slice over all sources
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
(nproc; lscpu | grep -i "model name\|cache\|thread\|socket"; gcc -O2 -o /tmp/pf scripts/pagefault_cost.c && /tmp/pf) > gpurun_out/r2k_host.txt 2>&1
cat gpurun_out/r2k_host.txt
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep --no-wall"
timeout 600 $B > gpurun_out/r2k_bench_plain.json 2> gpurun_out/r2k_bench_plain.err; echo "plain rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench_plain.json').read().strip().splitlines()[-1])
print("plain", d['value']/1e9, d['ms_per_step'], d.get('kernels'))
PY
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_gpu.txt
tail -4 gpurun_out/r2k_pytest_gpu.txt
KDF_BAM_TIMING=1 timeout 1500 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench rc=$?"
grep "kdf_bam" gpurun_out/r2k_bench_n1.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['parity_checked']['ok'], d['roofline']['kernel'], d['roofline']['frac'])
print(json.dumps(d['discovery_wall'])[:1800])
PY
