# round 2m: the state the round ends with — GPU tests, smoke(), the default bench line (with the
# BAM -> BED wall legs), the reference arm, launch list with DRAM bytes, ncu of the hot kernels
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total,driver_version --format=csv > gpurun_out/r2n_box.txt; nproc >> gpurun_out/r2n_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_gpu.txt
tail -4 gpurun_out/r2n_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2n_smoke.txt
KDF_BAM_TIMING=1 timeout 1200 python bench.py > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err; echo "bench rc=$?"
grep "kdf_bam\|Step\|finished" gpurun_out/r2n_bench_n1.err | tail -30 > gpurun_out/r2n_wall_log.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; echo "ref rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep --no-wall"
$B > gpurun_out/r2n_bench_plain_under_profile_cmd.json 2> gpurun_out/r2n_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/r2n_launches.csv $B > gpurun_out/r2n_ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_packed_keys -s 70 -c 1 -f -o gpurun_out/r2n_k_packed_keys $B > gpurun_out/r2n_ncu_a.log 2>&1; echo "ncu a rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['parity_checked']['ok'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('l2_mixed_rate_fraction'), d['roofline']['traffic'])
print(json.dumps(d['discovery_wall'])[:1800])
print(d['k_sweep'])
r=json.loads(open('gpurun_out/r2n_bench_reference.json').read().strip().splitlines()[-1])
print("reference", r['value']/1e9, r['ran'], r['cpu_baseline']['cores'])
PY
cat gpurun_out/r2n_wall_log.txt | tail -16
